#!/bin/bash
# round-2 visit L: N = $1 GPUs, final bench line (+ multi-GPU parity tests at N = 2, NCCL-fallback variant at N = 2)
N=${1:-1}
O=gpurun_out
mkdir -p $O
if [ "$N" = "1" ]; then
  BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02l_bench_n1.json 2> $O/r02l_bench_n1.err
else
  if [ "$N" = "2" ]; then
    timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02l_pytest_multi.log 2>&1; tail -3 $O/r02l_pytest_multi.log
  fi
  BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02l_bench_n$N.json 2> $O/r02l_bench_n$N.err
fi
tail -c 300 $O/r02l_bench_n$N.err
python - <<PY
import json
d = json.load(open("$O/r02l_bench_n$N.json"))
e = d["e2e"]
print("N=$N value %.1f GFLOP/s  %.4f ms  kernel_only %.4f ms  frac %.3f | e2e %.3f ms = %.3f of the link ceiling %.3f (registered %.3f of %.3f, unregistered %.1f)  parity %s %s" % (
    d["value"], d["ms_per_step"], d["kernel_only"]["ms_per_step"], d["roofline"]["frac"], e["ms_per_step"], e["link_ceiling"]["frac_achieved"], e["link_ceiling"]["duplex_ms"],
    e["registered"]["ms_per_step"], e["registered"]["link_ceiling_duplex_ms"], e["unregistered"]["ms_per_step"], d["parity"]["bit_identical"], d["parity"]["e2e_bit_identical"]))
PY
if [ "$N" = "2" ]; then
  BENCH_FORCE_NCCL_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 5 --e2e-blocks 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NCCL-fallback exchange: value %.1f  %.4f ms (kernel only %.4f) e2e %.3f ms parity %s | %s' % (d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['exchange'][:60]))" | tee $O/r02l_bench_n${N}_nccl_fallback.log
fi
