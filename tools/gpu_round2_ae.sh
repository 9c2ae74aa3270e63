#!/bin/bash
# visit AE: the lean form of the x-window kernel as the default of plain launches: whole GPU suite on two GPUs, slice timing, bench lines at N = 1 and 2
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02ae_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ae_pytest_gpu.log
for lean in 0 5; do
  SPMVB200_XW_LEAN=$lean timeout 300 python tools/xwbench.py cfg4s 2048:8192 --reps 20 2>&1 | grep -v "^#" | sed "s/^/lean=$lean /;s/nw=- u=- nbuf=-  *//" | tee -a $O/r02ae_xw_lean.log
  SPMVB200_XW_LEAN=$lean timeout 300 python tools/xwbench.py cfg4 2048:8192 --reps 10 2>&1 | grep -v "^#" | sed "s/^/lean=$lean /;s/nw=- u=- nbuf=-  *//" | tee -a $O/r02ae_xw_lean.log
done
BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02ae_bench_n1.json 2> $O/r02ae_bench_n1.err; echo "bench n1 rc=$?"
python -c "import json; d=json.load(open('$O/r02ae_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['parity']['bit_identical'])"
BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02ae_bench_n2.json 2> $O/r02ae_bench_n2.err; echo "bench n2 rc=$?"
python -c "import json; d=json.load(open('$O/r02ae_bench_n2.json')); print(d['value'], d['ms_per_step'], d['kernel_only']['ms_per_step'], d['e2e']['ms_per_step'], d['parity']['bit_identical'], d['parity']['ranks_checked'])"
