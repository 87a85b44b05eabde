#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8 > gpurun_out/r2j_test.log; cat gpurun_out/r2j_test.log
for B in 8192 16384 32768 65536; do
for sch in auto tensor; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --sustain 0 --dtype bfloat16 --trees $B --schedule $sch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 B=$B sched=$sch kernel_ms %.4f'%d['roofline']['kernel_ms'], d['plan']['schedule'])"
done; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --dtype bfloat16 --workload cartpole-h128 --schedule tensor 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bf16 H128 tensor kernel_ms %.4f'%d['roofline']['kernel_ms'])"
