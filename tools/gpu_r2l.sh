#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --sustain 0 > gpurun_out/r2l_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_team -s 3 -c 1 -o gpurun_out/r2l_team \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary --sustain 0 > gpurun_out/r2l_ncu.log 2>&1
ls -la gpurun_out/r2l_team.ncu-rep
