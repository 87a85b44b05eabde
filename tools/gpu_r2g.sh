#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -k "all_geometries and A2H64S50 and layerwg" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --phase-clocks > gpurun_out/r2g_phase.json 2> gpurun_out/r2g.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_phase.json')); print({k:round(v) for k,v in d['cycles_per_simulation_mean'].items()})"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/r2g_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_team -s 3 -c 1 -o gpurun_out/r2g_team \
    python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/r2g_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_plain.log').read().strip().splitlines()[-1]); print('kern', d['roofline']['kernel_ms'])
PY
