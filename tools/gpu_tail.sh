#!/bin/bash
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "rmat or scaled or edge or hybrid or deterministic or pipelined") > gpurun_out/pytest_tail.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tail.log
SPMVB200_VERBOSE=1 python tools/kbench.py cfg3 --reps 25 2>&1 | grep -E "csr_warp|adapt|tuning" | cut -c1-330
SPMVB200_NO_WARP_MID=1 python tools/kbench.py cfg3 --reps 25 2>&1 | grep -E "csr_warp|adapt" | cut -c1-160
CMD="python tools/kbench.py cfg3 --reps 3"
SPMVB200_SERIAL_TAIL=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cfg3_launches.csv $CMD > gpurun_out/cfg3_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/cfg3_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
seq=[(r[kn][:60], float(r[mv].replace(',','')) * (1e-3 if r[mu]=='ns' else 1.0)) for r in rows[start:] if len(r)>mv]
seen=set()
for k,v in seq:
    if any(t in k for t in ("csr_midrow","csr_longrow","csr_vector_kernel<2")) and k not in seen:
        seen.add(k); print("%-62s %9.1f us"%(k,v))
PY
