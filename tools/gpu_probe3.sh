#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
P=29550
for sk in "warmup,kloop" "warmup,kloop,allreduce" "kloop"; do
P=$((P+1))
BENCH_SKIP=$sk BENCH_NO_NVML=1 $TR --master-port $P bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 60 > gpurun_out/b3.out 2>gpurun_out/b3.err
python -c "import json; d=json.loads(open('gpurun_out/b3.out').read()); print('skip=$sk', d['value'], d['e2e']['ms_per_step'])"
done
