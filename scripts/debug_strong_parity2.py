"""GPU debug: bench flow (pair matrices, timed loop, e2e, parity check) for every emulated rank of an 8-way partition, no NCCL."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from turtlefsi_b200.mesh import synthetic_state, tube_mesh
from turtlefsi_b200.partition_rcb import RankView, rcb_partition
from turtlefsi_b200.space import MixedSpace
n, world = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
gmesh = tube_mesh(n, seed=1234)
gspace = MixedSpace(gmesh)
part = rcb_partition(gmesh, world, metric="cells")
xg, x1g = synthetic_state(gmesh, gspace, 2024, h=2.0 / n), synthetic_state(gmesh, gspace, 2025, h=2.0 / n)
args = types.SimpleNamespace()
for r in range(world):
    view = RankView(gmesh, r, world, cell_part=part, gspace=gspace, balanced=True)
    asm = bench.make_assembler(view.mesh, view.space, dev, 2.0 / n)
    x, x1 = view.scatter_global(xg), view.scatter_global(x1g)
    bc = bench.tube_bc(view.space, view.mesh, 4.0)
    asm.set_bc_dofs(bc)
    ghost = np.nonzero(~view.dof_owned)[0]
    w = bench.Workload(view.mesh, view.space, asm, x, x1, bc, None, view.cell_owned.sum(), view.node_owned, ghost, 0.0, "strong")
    for full in (False, True):
        m = bench.measure(w, args, dev, 1, None, 0, 5, 3, full=full)
        p = bench.parity_check(w, dev, 1, None)
        print("rank", r, "full" if full else "short", "step %.2f kf %.2f" % (m["ms_step"], m["kf_ms"]), "errA %.2e errb %.2e" % (p["max_rel_A"], p["max_rel_b"]), flush=True)
    del asm, w
    torch.cuda.empty_cache()
