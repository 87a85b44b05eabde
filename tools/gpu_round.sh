#!/bin/bash
# One GPU call that produces every judged artifact of a round.  Usage: tools/gpu_round.sh <tag> [kernel-regex]
#   full parity suite, smoke(), the default bench line (+ reference arm), then — only after the plain run
#   exited 0 — the ncu launch list and one --set full capture of the dominant kernel.
tag=$1; kern=${2:-search_team}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6 > gpurun_out/${tag}_test.log; tail -3 gpurun_out/${tag}_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench_reference.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_list.log 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_plain2.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 3 -c 1 -o gpurun_out/${tag}_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/${tag}_fused.ncu-rep gpurun_out/${tag}_launches.csv
