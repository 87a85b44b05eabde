#!/bin/bash
# usage: tools/gpu_scale.sh <tag> <N>   — the driver's multi-GPU bench command for N ranks
tag=$1; N=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N > gpurun_out/${tag}_scale_${N}.json 2> gpurun_out/${tag}_scale_${N}.err
echo rc=$?; tail -c 1500 gpurun_out/${tag}_scale_${N}.json; tail -5 gpurun_out/${tag}_scale_${N}.err
