"""Where the fluid Jacobian kernel spends its cycles: run with a library built with -DTFSI_PHASE_CLOCKS
(TFSI_LIB=turtlefsi_b200/libturtlefsi_b200_clk.so python scripts/phase_clocks.py [ncube]). Cycle sums over warps."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 69
w = bench.workload_single(n, torch.device("cuda", 0))
bench.prepare_matrices(w)
step = bench.make_step(w)
from turtlefsi_b200 import lib as _lib  # noqa: E402
L = _lib.load()
out = (C.c_ulonglong * 16)()
for _ in range(2):
    step()
L.tfsi_debug_phase_clocks(out)
w.asm.set_profiling(True)
for _ in range(3):
    step()
L.tfsi_debug_phase_clocks(out)
print("fluid kernel ms per launch (instrumented build):", w.asm.kernel_time(0)[0] / 3)
v = list(out)
names = ["stage+gather", "phase0 work", "phase0 barrier wait", "big tasks", "  loop (in big)", "  mbar wait (in big)",
         "  scatter (in big)", "small tasks", "warp lifetime"]
tot = v[8]
print("CTAs", v[9], "warps", v[9] * 12, "mean warp lifetime (cycles)", tot / (v[9] * 12))
for i, nm in enumerate(names):
    print("%-24s %6.2f %%   %10.0f cycles per warp" % (nm, 100.0 * v[i] / tot, v[i] / (v[9] * 12)))
