#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 300 --warmup 10 2>gpurun_out/scale3_cfg2_n8.err > gpurun_out/scale3_cfg2_n8.json; echo "cfg2 N=8 rc=$?"
timeout 400 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --workload cfg4 --steps 30 --warmup 5 2>gpurun_out/scale3_cfg4_n8.err > gpurun_out/scale3_cfg4_n8.json; echo "cfg4 N=8 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 300 --warmup 10 2>gpurun_out/scale3_cfg2_n4.err > gpurun_out/scale3_cfg2_n4.json; echo "cfg2 N=4 rc=$?"
python - <<'PY'
import json
for f in ("scale3_cfg2_n8","scale3_cfg4_n8","scale3_cfg2_n4"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.0f ms %.4f frac %.3f kernel %s | e2e %.1f GFLOP/s %.3f ms | parity %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"], d["clocks"]))
    except Exception as e: print(f, "missing", e)
PY
tail -3 gpurun_out/scale3_cfg2_n8.err
