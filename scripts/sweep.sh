#!/bin/bash
# BASELINE config 5: assembly sweep over mesh sizes (per-GPU tube of 6*n^3 tets)
for n in "$@"; do
  timeout 600 python bench.py --ncube $n --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']['workload'].split(',')[1]
print('ncube $n |', c.strip(), '| %.2f ms | %.1f M cells/s | fluid frac %.3f | setup %.1f s' % (d['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['setup_s']))"
done
