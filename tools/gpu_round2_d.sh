#!/bin/bash
# round-2 GPU visit D (1 GPU): full GPU test-suite, drop-in register-budget variants, L2 persistence of x A/B, bench + launch list
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02d_pytest.log 2>&1; tail -3 $O/r02d_pytest.log
B=tests/integration/_build
for w in "lap2d 1024" "stencil27 128"; do
  echo "=== drop-in, fast tier, 64-register budget"; $B/dropin_bench_b200 $w
  echo "=== drop-in, fast tier, 32-register budget (launch_bounds(1024,2))"; $B/dropin_bench_b200_r32 $w
done 2>&1 | tee $O/r02d_dropin_regs.log | grep -E "===|CSR 0|ELL 0"
for xp in 0 1; do
  for t in "cfg5 csr_rows 32 0.15" "cfg5 ell_rows 32 0.15" "cfg3 csr_warp" "cfg2 sell_rows" "cfg2 csr_warp"; do
    echo -n "X_PERSIST=$xp  "; SPMVB200_X_PERSIST=$xp NCU_TARGET_REPS=20 python tools/ncu_target.py $t
  done
done 2>&1 | tee $O/r02d_x_persist_ab.log
SPMVB200_X_PERSIST=1 timeout 600 ncu --set full --clock-control none -k regex:sell_kernel -s 8 -c 1 -f -o $O/r02d_sell_cfg5_xpersist python tools/ncu_target.py cfg5 csr_rows 32 0.15 > /dev/null 2>&1
BENCH_VERBOSE=1 timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02d_bench.json 2> $O/r02d_bench.err; tail -c 300 $O/r02d_bench.err
python -c "
import json; d=json.load(open('$O/r02d_bench.json')); print('N=1 value %.1f  %.4f ms  frac %.3f  e2e %.3f ms (pinned %.3f, ceiling %.3f)  cfg1 %.2f us graph %.2f us  cfg2 %.2f us  parity %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['pinned']['ms_per_step'], d['e2e']['link_ceiling']['duplex_ms'], d['cfg1']['ms_per_step']*1e3, d['cfg1']['graph_replay_ms_per_step']*1e3, d['cfg2']['ms_per_step']*1e3, d['parity']['bit_identical']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02d_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-side --e2e-steps 3 --e2e-blocks 1 > $O/r02d_ncu_bench.log 2>&1
