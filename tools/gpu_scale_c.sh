#!/bin/bash
# strong-scaled cfg4 and weak-scaled cfg2 at N GPUs (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29721 bench.py --gpus $N --workload cfg4 --steps 30 --warmup 5 2>gpurun_out/scale4_cfg4_n$N.err > gpurun_out/scale4_cfg4_n$N.json; echo "cfg4 N=$N rc=$?"
timeout 300 $TR --master-port 29722 bench.py --gpus $N --steps 300 --warmup 10 2>gpurun_out/scale4_cfg2_n$N.err > gpurun_out/scale4_cfg2_n$N.json; echo "cfg2 N=$N rc=$?"
python - <<PY
import json
for f in ("scale4_cfg4_n$N","scale4_cfg2_n$N"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, "value %.0f ms %.4f frac %.3f kernel %s | e2e %.1f GFLOP/s %.3f ms | parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["parity"]))
    except Exception as e: print(f, "missing", e)
PY
