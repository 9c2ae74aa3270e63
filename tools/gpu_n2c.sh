#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29711 bench.py --gpus 2 --steps 300 --warmup 10 2>gpurun_out/n2c.err > gpurun_out/n2c_cfg2.json; echo "cfg2 rc=$?"
timeout 300 $TR --master-port 29712 tools/multi_gpu_iterate.py > gpurun_out/iterate_n2c.log 2>&1; echo "iterate rc=$?"
grep -E "parity|ITERATE" gpurun_out/iterate_n2c.log | tail -12 | cut -c1-400
timeout 300 $TR --master-port 29713 tools/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU" 
python - <<'PY'
import json
d=json.load(open("gpurun_out/n2c_cfg2.json")); print("value %.0f ms %.4f frac %.3f kernel %s | e2e %.1f GFLOP/s %.3f ms %s | parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["path"][:60], d["parity"]))
PY
