#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv or config2 or uint8 or tensor_core" 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('atari ms/step',round(d['ms_per_step'],4),'enc',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['ms_per_step'],4))
"
