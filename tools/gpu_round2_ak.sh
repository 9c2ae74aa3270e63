#!/bin/bash
# visit AK (last of round 2): the per-lane count words packed for one load per tile: whole GPU suite on two GPUs, smoke, bench at N = 1 and 2
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02ak_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ak_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02ak_bench_n1.json 2> $O/r02ak_bench_n1.err; echo "bench n1 rc=$?"
python -c "import json; d=json.load(open('$O/r02ak_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['cfg2']['ms_per_step'], d['cfg1']['ms_per_step'])"
BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02ak_bench_n2.json 2> $O/r02ak_bench_n2.err; echo "bench n2 rc=$?"
python -c "import json; d=json.load(open('$O/r02ak_bench_n2.json')); print(d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['parity']['ranks_checked'])"
