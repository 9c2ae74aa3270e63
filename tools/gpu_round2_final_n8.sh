#!/bin/bash
# final visit of round 2, eight GPUs: the contract bench line at N = 8 (strong scaling of cfg4 with the exchange inside the timed step)
O=gpurun_out
mkdir -p $O
BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02ae_bench_n8.json 2> $O/r02ae_bench_n8.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/r02ae_bench_n8.json')); print(d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['parity']['ranks_checked'])"
