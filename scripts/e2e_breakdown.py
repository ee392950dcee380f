"""where the end-to-end call (tfsi_newton_assemble_host) spends its time: wall time per call against the kernel times the
handle records inside it (development aid).  python scripts/e2e_breakdown.py [ncube]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 69
dev = torch.device("cuda:0")
w = bench.workload_single(n, dev)
bench.prepare_matrices(w)
asm = w.asm
N = w.space.num_dofs
xh, x1h = torch.as_tensor(w.x).pin_memory(), torch.as_tensor(w.x1).pin_memory()
gh = torch.zeros(len(w.bc), dtype=torch.float64).pin_memory()
bh = torch.empty(N, dtype=torch.float64).pin_memory()
for _ in range(3):
    asm.newton_assemble_host(xh, x1h, gh, "compressed", bench.next_matrix(w), bh, recompute=True)
torch.cuda.synchronize()
for label, prof in (("plain", False), ("kernel timers on", True)):
    asm.set_profiling(prof)
    base = [asm.kernel_time(k) for k in range(3)]
    steps = 8
    t0 = time.perf_counter()
    for it in range(steps):
        asm.newton_assemble_host(xh, x1h if it % 4 == 0 else None, gh, "compressed", bench.next_matrix(w), bh, recompute=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    k = [asm.kernel_time(i) for i in range(3)]
    parts = [(k[i][0] - base[i][0]) / steps for i in range(3)]
    print("%-18s call %.2f ms | fluid %.2f solid %.2f residual %.2f (%d launches per call) -> kernels %.2f, rest %.2f"
          % (label, ms, parts[0], parts[1], parts[2], (k[2][1] - base[2][1]) // steps, sum(parts), ms - sum(parts)))
# PCIe rates of the buffers the call moves
xd = torch.empty(N, dtype=torch.float64, device=dev)
for name, f in (("H2D x", lambda: xd.copy_(xh, non_blocking=True)), ("D2H b", lambda: bh.copy_(xd, non_blocking=True))):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%s: %.2f ms for %.1f MB = %.1f GB/s" % (name, ms, N * 8 / 1e6, N * 8 / ms / 1e6))
