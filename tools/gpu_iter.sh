#!/bin/bash
# One GPU iteration: parity tests, bench sweep, one ncu capture of the fused kernel.  Usage: tools/gpu_iter.sh <tag> [lanes...]
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6 > gpurun_out/test_$tag.log; tail -3 gpurun_out/test_$tag.log
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[2],"value %.3e e2e %.3e ms/step %.3f kern_ms %.3f frac %.4f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["kernel_ms"],d["roofline"]["frac"]),d["plan"])
PY
}
for l in "$@"; do python bench.py --steps 20 --warmup 3 --no-cpu --lanes $l 2>&1 | tail -1 > gpurun_out/bench_${tag}_l$l.json; show gpurun_out/bench_${tag}_l$l.json "B=4096 lanes $l"; done
for l in 4 8; do python bench.py --steps 10 --warmup 3 --no-cpu --trees 32768 --lanes $l 2>&1 | tail -1 > gpurun_out/bench_${tag}_big_l$l.json; show gpurun_out/bench_${tag}_big_l$l.json "B=32768 lanes $l"; done
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:search_fused -s 3 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
ls -la gpurun_out/prof_$tag.ncu-rep
