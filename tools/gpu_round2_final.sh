#!/bin/bash
# final visit of round 2: the whole GPU suite, smoke(), the contract bench line and its ncu launch list (after the clean run)
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02ac_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ac_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02ac_bench_n1.json 2> $O/r02ac_bench_n1.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/r02ac_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['cfg2']['ms_per_step'], d['cfg1']['ms_per_step'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02ac_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-side --e2e-steps 3 --e2e-blocks 1 > $O/r02ac_ncu.log 2>&1; echo "ncu rc=$?"
