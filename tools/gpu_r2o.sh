#!/bin/bash
for ctas in 1 2 3 4; do
CUMIND_CONV_CTAS=$ctas timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('ctas=$ctas atari ms/step %.4f encoder_ms %.4f frac %.3f search_ms %.4f e2e_ms %.3f'%(d['ms_per_step'], r['kernel_ms'], r['frac'], r['search_ms'], d['e2e']['ms_per_step']))"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --sustain 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cartpole value %.4g ms %.4f kern %.4f e2e_ms %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step']))"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "conv or config2 or all_geometries" 2>&1 | tail -3
