#!/bin/bash
# A/B runs of bench.py with several builds of the library: scripts/ab_bench.sh <lib suffix> ...   (prints step / fluid / solid ms)
# BENCH_ARGS: extra bench.py flags (e.g. --no-pair)
for v in "$@"; do
  L=$PWD/turtlefsi_b200/libturtlefsi_b200$v.so
  TFSI_LIB=$L python bench.py --no-cpu-baseline --newton 0 --steps 5 $BENCH_ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); b=d['breakdown_ms']; print('$v', 'step %.2f fluid %.2f solid %.2f init %.2f parity %.1e pair %s e2e %.2f' % (d['ms_per_step'], b['jac_fluid_kernel'], b['jac_solid_kernel'], b['init_A_pre_rows_if_foreground'], d['parity']['max_rel_A'], d['problem_size']['matrix_pair'], d['e2e']['ms_per_step']))
"
done
