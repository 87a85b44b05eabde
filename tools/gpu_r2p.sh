#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari > gpurun_out/r2p_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2p_atari_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --sustain 0 --no-secondary --workload atari > gpurun_out/r2p_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2p_atari_launches.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
for r in rows[hi+1:hi+40]:
    print(r[kn][:70], r[mv])
PY
