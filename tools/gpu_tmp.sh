#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in 4096 7000 8192 9600; do
timeout 300 python bench.py --trees $n --steps 50 --warmup 5 --no-cpu --sustain 0 --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print($n,'auto', 'ms/step',round(d['ms_per_step'],4),'kernel',round(d['roofline']['kernel_ms'],4), d['plan']['schedule'], d['roofline']['kernel'])
"
done
