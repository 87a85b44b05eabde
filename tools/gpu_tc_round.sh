#!/bin/bash
# tensor-core search: bench at three batch sizes, then one --set full capture of search_tc_kernel at 32 768 trees
tag=$1
B="timeout 400 python bench.py --no-cpu --no-secondary --sustain 0 --steps 30 --warmup 3 --dtype bfloat16"
for n in 16384 32768 65536; do
  $B --trees $n > gpurun_out/${tag}_bench_B${n}_tc.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench_B${n}_tc.json').read().strip().splitlines()[-1]); print($n, round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), '%.3e'%d['value'])"
done
bash tools/gpu_ncu_kernel.sh ${tag}_tc search_tc 3 -- --no-cpu --no-secondary --sustain 0 --steps 2 --warmup 3 --dtype bfloat16 --trees 32768
