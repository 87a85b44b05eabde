#!/bin/bash
for sl in 1 2 3 4; do
CUMIND_U8_SLICES=$sl timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('slices $sl e2e ms', round(d['e2e']['ms_per_step'],4))
"
done
