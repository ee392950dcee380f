#!/bin/bash
# A/B of library variants turtlefsi_b200/lib_<name>.so against the current build (development aid)
cp turtlefsi_b200/libturtlefsi_b200.so /tmp/orig.so
echo "== current"
timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['breakdown_ms']['jac_fluid_kernel'], d['breakdown_ms']['jac_solid_kernel'], d['roofline']['frac'])"
for v in "$@"; do
  cp turtlefsi_b200/lib_$v.so turtlefsi_b200/libturtlefsi_b200.so
  echo "== $v"
  timeout 300 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['breakdown_ms']['jac_fluid_kernel'], d['breakdown_ms']['jac_solid_kernel'], d['roofline']['frac'])"
done
cp /tmp/orig.so turtlefsi_b200/libturtlefsi_b200.so
