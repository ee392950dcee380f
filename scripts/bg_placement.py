"""GPU experiment: on which SMs do the blocks of the persistent background copy land (one per SM wanted)?"""
import sys, os
os.environ["TFSI_BG_STATS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
w = bench.workload_single(int(sys.argv[1]) if len(sys.argv) > 1 else 40, torch.device("cuda"))
bench.prepare_matrices(w)
step = bench.make_step(w)
for _ in range(3):
    step()
torch.cuda.synchronize()
w.asm.bg_placement()
n = 6
for _ in range(n):
    step()
torch.cuda.synchronize()
c = w.asm.bg_placement()[:160]
print("launches", n, "hist of blocks per SM per launch:", np.bincount((c[:148] // 1)), "sum", c.sum())
print(c[:148])
