#!/bin/bash
# development aid: residual kernel variants (cells per CTA) on the bench workload
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('residual ms', d['breakdown_ms']['residual'], 'e2e ms', d['e2e']['ms_per_step'])"; }
for nc in 16 32; do
  echo "== TFSI_RES_NC=$nc"
  TFSI_RES_NC=$nc timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | show
done
