#!/bin/bash
# One --set full capture of one kernel inside a bench command, after the plain command exited 0.
# Usage: tools/gpu_ncu_kernel.sh <tag> <kernel-regex> <skip> -- <bench args...>
tag=$1; kern=$2; skip=$3; shift 4
mkdir -p gpurun_out
timeout 400 python bench.py "$@" > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -o gpurun_out/${tag} \
    python bench.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -c 400 gpurun_out/${tag}_plain.json; ls -la gpurun_out/${tag}.ncu-rep
