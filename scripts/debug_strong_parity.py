"""GPU debug: every rank of an 8-way strong partition emulated on one GPU; spot rows vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle.spotcheck import check_nodes
from turtlefsi_b200.mesh import synthetic_state, tube_mesh
from turtlefsi_b200.partition_rcb import RankView, rcb_partition
from turtlefsi_b200.space import MixedSpace
n, world = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda")
gmesh = tube_mesh(n, seed=1234)
gspace = MixedSpace(gmesh)
part = rcb_partition(gmesh, world, metric="cells")
xg, x1g = synthetic_state(gmesh, gspace, 2024, h=2.0 / n), synthetic_state(gmesh, gspace, 2025, h=2.0 / n)
for r in range(world):
    view = RankView(gmesh, r, world, cell_part=part, gspace=gspace, balanced=True)
    asm = bench.make_assembler(view.mesh, view.space, dev, 2.0 / n)
    x, x1 = view.scatter_global(xg), view.scatter_global(x1g)
    xd, x1d = torch.as_tensor(x).to(dev), torch.as_tensor(x1).to(dev)
    A_pre = asm.assemble_A_pre(xd, x1d)
    for mode in ("dense", "compressed"):
        if mode == "compressed":
            asm.compress_A_pre(A_pre)
        A = asm.assemble_jacobian(xd, x1d, A_pre)
        b = asm.assemble_residual(xd, x1d)
        torch.cuda.synchronize()
        rng = np.random.default_rng(100 + r)
        owned = np.nonzero(view.node_owned)[0]
        picks = list(rng.choice(owned, size=3, replace=False))
        rowptr, colidx = asm.rowptr, asm.colidx
        def get_row(rw):
            lo, hi = int(rowptr[rw].item()), int(rowptr[rw + 1].item())
            return colidx[lo:hi].cpu().numpy(), A[lo:hi].cpu().numpy()
        eA, eb, sb = check_nodes(view.mesh, view.space, bench.REGIONS, bench.COMMON, x, x1, picks, get_row, lambda rw: b[rw].item())
        print("rank", r, mode, "cells", view.mesh.num_cells, "nnz", asm.nnz, "errA %.2e errb %.2e" % (eA, eb / float(b.abs().max())), flush=True)
    del asm, A, b, A_pre
    torch.cuda.empty_cache()
