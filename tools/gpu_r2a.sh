#!/bin/bash
# round 2, call A: microbenchmarks of the layer designs + the PDL / FFMA2 state of the product path
mkdir -p gpurun_out
tools/microbench_lds.bin > gpurun_out/r2a_lds.txt 2>&1; cat gpurun_out/r2a_lds.txt
tools/mb/layer_bench.bin > gpurun_out/r2a_layer.txt 2>&1; cat gpurun_out/r2a_layer.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -6 > gpurun_out/r2a_test.log; tail -3 gpurun_out/r2a_test.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --phase-clocks > gpurun_out/r2a_phase.json 2>> gpurun_out/r2a_bench.err; cat gpurun_out/r2a_phase.json
