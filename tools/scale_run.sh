#!/bin/bash
# Multi-GPU scaling on one 8-GPU box (run under `gpurun --gpus 8`): weak-scaled cfg2 and strong-scaled cfg4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
P=29600
for N in 8 4 2; do
  P=$((P+1)); timeout 300 $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --steps 300 --warmup 10 2>gpurun_out/scale_cfg2_n$N.err > gpurun_out/scale_cfg2_n$N.json; echo "cfg2 N=$N rc=$?"
  P=$((P+1)); timeout 400 $TR --nproc-per-node $N --master-port $P bench.py --gpus $N --workload cfg4 --steps 30 --warmup 5 2>gpurun_out/scale_cfg4_n$N.err > gpurun_out/scale_cfg4_n$N.json; echo "cfg4 N=$N rc=$?"
done
P=$((P+1)); timeout 300 $TR --nproc-per-node 4 --master-port $P tools/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU|world" | tail -8
python bench.py --steps 300 --warmup 10 --no-cpu > gpurun_out/scale_cfg2_n1.json 2>/dev/null
python bench.py --workload cfg4 --steps 30 --warmup 5 --no-cpu > gpurun_out/scale_cfg4_n1.json 2>/dev/null
python - <<'PY'
import json, glob
for wl in ("cfg2", "cfg4"):
    for n in (1, 2, 4, 8):
        try:
            d = json.load(open("gpurun_out/scale_%s_n%d.json" % (wl, n)))
            print("%s N=%d value %.1f GFLOP/s  ms/step %.4f  hbm %.0f GB/s  frac/GPU %.3f  e2e %.1f GFLOP/s (%.3f ms)  kernel %s clocks %s" % (
                wl, n, d["value"], d["ms_per_step"], d["hbm_gbs"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel"], d["clocks"]))
        except Exception as e:
            print(wl, n, "missing", e)
PY
