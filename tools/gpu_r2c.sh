#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --phase-clocks > gpurun_out/r2c_phase.json 2> gpurun_out/r2c.err; cat gpurun_out/r2c_phase.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/r2c_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_team -s 3 -c 1 -o gpurun_out/r2c_team \
    python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out/r2c_team.ncu-rep; tail -3 gpurun_out/r2c_ncu.log
