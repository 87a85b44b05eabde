#!/bin/bash
# Secondary artifacts of a round (no profiler in this call).  Usage: tools/gpu_round_extra.sh <tag>
tag=$1
mkdir -p gpurun_out
B="timeout 400 python bench.py --no-cpu --no-secondary --sustain 0 --steps 30 --warmup 3"
$B --workload atari > gpurun_out/${tag}_bench_atari.json 2>/dev/null
$B --workload atari-literal > gpurun_out/${tag}_bench_atari-literal.json 2>/dev/null
$B --workload cartpole-h128 > gpurun_out/${tag}_bench_h128.json 2>/dev/null
$B --workload cartpole-h128 --dtype bfloat16 --schedule tensor > gpurun_out/${tag}_bench_h128_tc.json 2>/dev/null
for n in 16384 32768 65536; do
  $B --trees $n > gpurun_out/${tag}_bench_B${n}.json 2>/dev/null
  $B --trees $n --dtype bfloat16 > gpurun_out/${tag}_bench_B${n}_tc.json 2>/dev/null
done
timeout 300 python bench.py --phase-clocks --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/${tag}_phase_clocks.json
PYTHONPATH=. timeout 300 python tools/bench_selfplay.py > gpurun_out/${tag}_selfplay.log 2>&1
PYTHONPATH=. timeout 300 python tools/bench_replay.py > gpurun_out/${tag}_replay.log 2>&1
PYTHONPATH=. timeout 300 python tools/bench_train.py > gpurun_out/${tag}_train.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${tag}_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'],4), 'value %.3e'%d['value'], d['plan']['schedule'], 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4))
    except Exception as e: print(f,'ERR',e)
PY
for f in selfplay replay train; do tail -n 3 gpurun_out/${tag}_$f.log; done
