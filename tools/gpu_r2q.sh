#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "uint8 or host_path or config2" 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('atari ms/step',d['ms_per_step'],'e2e',d['e2e'])
"
