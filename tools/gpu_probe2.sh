#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export PROBE_ONLY=whole_step_synced,allgather_synced,whole_step
PROBE_REPS=50 $TR --master-port 29541 tools/e2e_multi_probe.py 2>&1 | grep E2E_PROBE
PROBE_REPS=200 $TR --master-port 29542 tools/e2e_multi_probe.py 2>&1 | grep E2E_PROBE
PROBE_REPS=1000 $TR --master-port 29543 tools/e2e_multi_probe.py 2>&1 | grep E2E_PROBE
