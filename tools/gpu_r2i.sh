#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "tensor_core or config2" 2>&1 | tail -25 > gpurun_out/r2i_t1.log; tail -12 gpurun_out/r2i_t1.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --dtype bfloat16 --trees 4096 > gpurun_out/r2i_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 3 -c 1 -o gpurun_out/r2i_tc \
    python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --dtype bfloat16 --trees 4096 > gpurun_out/r2i_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_plain.log').read().strip().splitlines()[-1]); print('bf16 B=8192 kern', d['roofline']['kernel_ms'])
PY
for args in "--trees 32768" "--workload cartpole-h128" "--workload atari"; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --dtype bfloat16 $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 $args ms/step', d['ms_per_step'], 'kernel', d['roofline'].get('kernel_ms', d['roofline'].get('search_ms')))"
done
