#!/bin/bash
# final visit of round 2, two GPUs: the two-GPU tests the driver's one-GPU box skips, and the contract bench line at N = 2
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02ac_pytest_multi_n2.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/r02ac_pytest_multi_n2.log
BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02ac_bench_n2.json 2> $O/r02ac_bench_n2.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/r02ac_bench_n2.json')); print(d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step'], d['e2e']['ms_per_step'], d['parity'])"
