set -x
python bench.py > gpurun_out/r4_bench_1gpu.json 2> gpurun_out/r4_bench_1gpu.err
python bench.py --no-pair --no-cpu-baseline --newton 0 > gpurun_out/r4_bench_1gpu_single_matrix.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4_bench_reference_arm.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --newton 0"
$B > gpurun_out/r4_pre.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4_bench_launches.csv $B > gpurun_out/r4_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jac_fluid_kernel --launch-skip 5 -c 1 -o gpurun_out/r4_fluid -f $B > gpurun_out/r4_ncu_fluid.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jac_solid_kernel --launch-skip 5 -c 1 -o gpurun_out/r4_solid -f $B > gpurun_out/r4_ncu_solid.log 2>&1
tail -c 300 gpurun_out/r4_bench_1gpu.json | head -c 300
