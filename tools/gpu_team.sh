#!/bin/bash
# team-kernel iteration: parity of every team variant, then the headline bench and the phase clocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "team or baseline_config or golden or leaves_the_bare or split or host_path" 2>&1 | tail -4
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --sustain 0 --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step',round(d['ms_per_step'],4),'kernel',round(d['roofline']['kernel_ms'],4),'e2e',round(d['e2e']['ms_per_step'],4), d['plan'])
"
timeout 300 python bench.py --phase-clocks --no-cpu --no-secondary 2>&1 | tail -1 > gpurun_out/team_phase.json; python -c "
import json; d=json.load(open('gpurun_out/team_phase.json')); print({k:round(v) for k,v in d['cycles_per_simulation_mean'].items()})"
