#!/bin/bash
# usage: sweep_multi.sh NGPU ncube...
N=$1; shift
for n in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --ncube $n --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['config']['workload'].split(',')[1]
print('gpus $N ncube $n |', c.strip(), '| %.2f ms | %.1f M cells/s total | e2e %.1f M cells/s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['value']/1e6))"
done
