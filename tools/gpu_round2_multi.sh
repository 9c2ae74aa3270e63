#!/bin/bash
# round-2 multi-GPU visit: N = $1 GPUs.  2-GPU parity tests (peer stores, shard), then bench.py at N.
N=${1:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo_n$N.log 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02_pytest_multi.log 2>&1; tail -3 $O/r02_pytest_multi.log
fi
BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
tail -c 400 $O/r02_bench_n$N.err
# A/B: the synchronisation fused into the boundary CTAs of the SpMV kernel instead of the separate barrier kernel
SPMVB200_SHARD_FUSED_SYNC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 5 --e2e-blocks 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fused-sync variant: value %.1f  %.4f ms (kernel only %.4f)' % (d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step']))" | tee $O/r02_bench_n${N}_fused_sync_variant.log
python - <<PY
import json
d = json.load(open("$O/r02_bench_n$N.json"))
print("N=$N value %.1f GFLOP/s  %.4f ms  kernel_only %.4f ms  e2e %.3f ms (pinned %.3f, ceiling %.3f)  parity %s" % (d["value"], d["ms_per_step"], d["kernel_only"]["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pinned"]["ms_per_step"], d["e2e"]["link_ceiling"]["duplex_ms"], d["parity"]))
PY

if [ "$N" = "2" ]; then
  BENCH_FORCE_NCCL_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 5 --e2e-blocks 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NCCL-fallback exchange: value %.1f  %.4f ms (kernel only %.4f) e2e %.3f ms parity %s | %s' % (d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['exchange'][:60]))" | tee $O/r02_bench_n${N}_nccl_fallback.log
fi
