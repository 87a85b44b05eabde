#!/bin/bash
for ns in 0 20 50 100 200 400; do
CUMIND_TEAM_WAIT_NS=$ns timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --sustain 0 --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('wait_ns $ns', 'ms/step',round(d['ms_per_step'],4),'kernel',round(d['roofline']['kernel_ms'],4),'e2e',round(d['e2e']['ms_per_step'],4))
"
done
