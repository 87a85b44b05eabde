#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -k "all_geometries and H64 and layerwg" 2>&1 | tail -15 > gpurun_out/r2e_t1.log; cat gpurun_out/r2e_t1.log
grep -q passed gpurun_out/r2e_t1.log || exit 1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8 > gpurun_out/r2e_test.log; cat gpurun_out/r2e_test.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value %.4g ms %.4f kern %.4f e2e_ms %.4f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']), d['plan'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; show gpurun_out/r2e_bench.json "default"
tail -3 gpurun_out/r2e_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --phase-clocks > gpurun_out/r2e_phase.json 2>> gpurun_out/r2e_bench.err; cat gpurun_out/r2e_phase.json
