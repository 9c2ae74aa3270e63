#!/bin/bash
# One 8-GPU visit: fused iterated SpMV vs NCCL, strong-scaled cfg4 and weak-scaled cfg2 bench lines.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29701 tools/multi_gpu_iterate.py > gpurun_out/iterate_n8.log 2>&1; echo "iterate rc=$?"
grep -E "parity|ITERATE" gpurun_out/iterate_n8.log | tail -12
timeout 400 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --workload cfg4 --steps 30 --warmup 5 2>gpurun_out/scale2_cfg4_n8.err > gpurun_out/scale2_cfg4_n8.json; echo "cfg4 N=8 rc=$?"; cat gpurun_out/scale2_cfg4_n8.json | cut -c1-600
timeout 300 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 300 --warmup 10 2>gpurun_out/scale2_cfg2_n8.err > gpurun_out/scale2_cfg2_n8.json; echo "cfg2 N=8 rc=$?"; cat gpurun_out/scale2_cfg2_n8.json | cut -c1-600
