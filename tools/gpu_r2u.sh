#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "team or baseline_config1 or golden" 2>&1 | tail -6
for ks in 1 0; do
CUMIND_TEAM_KS=$ks timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --sustain 0 --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('KS $ks', 'ms/step',round(d['ms_per_step'],4),'kernel',round(d['roofline']['kernel_ms'],4),'e2e',round(d['e2e']['ms_per_step'],4), d['plan'])
"
done
timeout 300 python bench.py --phase-clocks --no-cpu --no-secondary 2>&1 | tail -1 > gpurun_out/r2u_phase.json; python -c "
import json; d=json.load(open('gpurun_out/r2u_phase.json')); print({k:round(v) for k,v in d['cycles_per_simulation_mean'].items()})"
for pdl in 1 0; do
if [ $pdl = 0 ]; then export CUMIND_NO_PDL=1; fi
timeout 300 python bench.py --trees 8192 --steps 50 --warmup 5 --no-cpu --sustain 0 --no-secondary 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('B8192 pdl $pdl', 'ms/step',round(d['ms_per_step'],4),'kernel',round(d['roofline']['kernel_ms'],4),'e2e',round(d['e2e']['ms_per_step'],4), d['plan']['schedule'])
"
done
unset CUMIND_NO_PDL
timeout 60 tools/mb/lds_rate.bin > gpurun_out/r2u_lds_rate.txt; cat gpurun_out/r2u_lds_rate.txt
