#!/bin/bash
CMD="python tools/kbench.py cfg3 --reps 3"
$CMD > gpurun_out/cfg3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cfg3_launches.csv $CMD > gpurun_out/cfg3_ncu.log 2>&1
echo rc=$?
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/cfg3_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
seq=[(r[kn][:50], float(r[mv].replace(',','')) * (1e-3 if r[mu]=='ns' else 1.0)) for r in rows[start:] if len(r)>mv]
# last 12 launches
for k,v in seq[-14:]: print("%-52s %9.1f us"%(k,v))
PY
