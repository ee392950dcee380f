#!/bin/bash
# development aid: residual-kernel timing of library variants turtlefsi_b200/lib_<name>.so against the current build
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('residual ms', d['breakdown_ms']['residual'], 'e2e ms', d['e2e']['ms_per_step'])"; }
cp turtlefsi_b200/libturtlefsi_b200.so /tmp/orig.so
echo "== current"; timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | show
for v in "$@"; do
  cp turtlefsi_b200/lib_$v.so turtlefsi_b200/libturtlefsi_b200.so
  echo "== $v"; timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | show
done
cp /tmp/orig.so turtlefsi_b200/libturtlefsi_b200.so
