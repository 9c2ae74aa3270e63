#!/bin/bash
# round-2 GPU visit K (1 GPU): full GPU test-suite and the bench line after the host-path changes (no time-stamped chunk events,
# ramp schedule, explicit page-locking)
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02k_pytest.log 2>&1; tail -3 $O/r02k_pytest.log
BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02k_bench.json 2> $O/r02k_bench.err; tail -c 300 $O/r02k_bench.err
python -c "
import json; d=json.load(open('$O/r02k_bench.json')); print('N=1 value %.1f  %.4f ms  frac %.3f  e2e %.3f ms (unregistered %.3f, registered malloc %.3f, ceiling %.3f)  cfg1 %.2f us graph %.2f us  cfg2 %.2f us  parity %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['unregistered']['ms_per_step'], d['e2e']['registered']['ms_per_step'], d['e2e']['link_ceiling']['duplex_ms'], d['cfg1']['ms_per_step']*1e3, d['cfg1']['graph_replay_ms_per_step']*1e3, d['cfg2']['ms_per_step']*1e3, d['parity']))"
