#!/bin/bash
mkdir -p gpurun_out
export CUMIND_CONV_CTAS=1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari > gpurun_out/r2n_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo_tma -s 6 -c 1 -o gpurun_out/r2n_conv \
    python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari > gpurun_out/r2n_ncu.log 2>&1
ls -la gpurun_out/r2n_conv.ncu-rep
