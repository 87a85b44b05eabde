#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "tensor_core or config2" 2>&1 | tail -25 > gpurun_out/r2h_t1.log; cat gpurun_out/r2h_t1.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value %.4g ms %.4f kern %.4f e2e_ms %.4f'%(d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms', d['roofline'].get('search_ms',0)), d['e2e']['ms_per_step']), d['plan'], d['roofline'].get('frac'))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --workload atari > gpurun_out/r2h_atari.json 2> gpurun_out/r2h.err; show gpurun_out/r2h_atari.json atari
tail -3 gpurun_out/r2h.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --trees 32768 > gpurun_out/r2h_b32k.json 2>> gpurun_out/r2h.err; show gpurun_out/r2h_b32k.json "fp32 B=32768"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --trees 32768 --dtype bfloat16 > gpurun_out/r2h_b32k_tc.json 2>> gpurun_out/r2h.err; show gpurun_out/r2h_b32k_tc.json "bf16 B=32768"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --workload cartpole-h128 > gpurun_out/r2h_h128.json 2>> gpurun_out/r2h.err; show gpurun_out/r2h_h128.json "fp32 H=128"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --workload cartpole-h128 --dtype bfloat16 > gpurun_out/r2h_h128_tc.json 2>> gpurun_out/r2h.err; show gpurun_out/r2h_h128_tc.json "bf16 H=128"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --dtype bfloat16 > gpurun_out/r2h_b4k_tc.json 2>> gpurun_out/r2h.err; show gpurun_out/r2h_b4k_tc.json "bf16 B=4096"
tail -3 gpurun_out/r2h.err
