#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "all_geometries and A2H64S50 and layerwg" 2>&1 | tail -15 > gpurun_out/r2b_t1.log; cat gpurun_out/r2b_t1.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -8 > gpurun_out/r2b_test.log; cat gpurun_out/r2b_test.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], 'value %.4g ms %.4f kern %.4f e2e_ms %.4f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']), d['plan'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
for nw in 0 3 5; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 --net-warps $nw > gpurun_out/r2b_bench_nw$nw.json 2> gpurun_out/r2b_bench.err; show gpurun_out/r2b_bench_nw$nw.json "netwarps=$nw"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 --teams 3 --net-warps 4 > gpurun_out/r2b_bench_t3.json 2> gpurun_out/r2b_bench.err; show gpurun_out/r2b_bench_t3.json "teams=3"
tail -3 gpurun_out/r2b_bench.err
CUMIND_TEAM_LW=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 > gpurun_out/r2b_bench_classic.json 2>/dev/null; show gpurun_out/r2b_bench_classic.json classic
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --phase-clocks > gpurun_out/r2b_phase.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_phase.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --phase-clocks --net-warps 3 > gpurun_out/r2b_phase3.json 2>> gpurun_out/r2b_bench.err; cat gpurun_out/r2b_phase3.json
