#!/bin/bash
# Scaling on one 8-GPU box: weak (4096 trees per GPU) and strong (BASELINE configs[3], 65 536 trees in total)
# at N = 1, 2, 4, 8.  Usage (under gpurun --gpus 8): tools/gpu_scale.sh <tag>
tag=$1
mkdir -p gpurun_out
run() { n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu "$@" > gpurun_out/${tag}_${name}_$n.json 2> gpurun_out/${tag}_${name}_$n.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 50 --warmup 5 --no-cpu "$@" > gpurun_out/${tag}_${name}_$n.json 2> gpurun_out/${tag}_${name}_$n.err; fi
  python - gpurun_out/${tag}_${name}_$n.json "$name N=$n" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-14s value %.3e e2e %.3e ms/step %.3f sched %s trees/gpu %s" % (sys.argv[2], d["value"], d["e2e"]["value"], d["ms_per_step"], d["plan"].get("schedule"), d["config"].get("trees_per_gpu")))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for n in 1 2 4 8; do run $n weak; done
for n in 1 2 4 8; do run $n strong --total-trees 65536; done
