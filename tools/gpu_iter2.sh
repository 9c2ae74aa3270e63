#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_iterate.py > gpurun_out/iterate_n$N.log 2>&1; echo "iterate rc=$?"
grep -E "parity|ITERATE|Error|error|Traceback" gpurun_out/iterate_n$N.log | head -40; tail -5 gpurun_out/iterate_n$N.log
