"""ORACLE (test infrastructure only): spot check of an assembled matrix / residual at sizes the oracle cannot
assemble in full — the CSR rows of a few mesh nodes are recomputed by the numpy complex-step oracle from the cells
touching those nodes. Used by tests/test_gpu_fullsize.py and, OUTSIDE the timed region, by bench.py's in-run parity
check (every rank checks rows it owns after a real NCCL halo exchange)."""
import numpy as np

from .assembly import region_params
from .forms import CellForms


def check_nodes(mesh, space, regions, common, x, x1, nodes, get_row, get_b, deg_j=4, deg_r=6, part="all"):
    """nodes: P2-node ids of `space` (interleaved numbering). get_row(r) -> (cols int[], vals float[]) of CSR row r of
    the matrix under test; get_b(r) -> float. x, x1: host state vectors the oracle evaluates.
    Returns (max over rows of |A - ref| / max|ref_row|, max |b - ref|, max |ref b| over the checked rows)."""
    g = mesh.gdim
    cfJ, cfR = CellForms(g, space.biharmonic, deg_j), CellForms(g, space.biharmonic, deg_r)
    worst_A = worst_b = scale_b = 0.0
    for node in nodes:
        cells = np.nonzero((space.cell_nodes == node).any(axis=1))[0]
        cd = space.cell_dofs[cells].astype(np.int64)
        Xc = mesh.coords[mesh.cells[cells]]
        rows = np.arange(space.node_base[node], space.node_base[node] + space.node_ncomp[node])
        ref_rows = {int(rw): {} for rw in rows}
        ref_b = {int(rw): 0.0 for rw in rows}
        for reg in regions:
            sel = mesh.markers[cells] == reg["marker"]
            if not sel.any():
                continue
            prm = region_params(common, reg)
            Je = cfJ.jacobian(reg["kind"], Xc[sel], x[cd[sel]], x1[cd[sel]], prm, part)
            Rl, Rn = cfR.residual(reg["kind"], Xc[sel], x[cd[sel]], x1[cd[sel]], prm)
            for ci, dofs in enumerate(cd[sel]):
                for li, rw in enumerate(dofs):
                    if int(rw) in ref_rows:
                        ref_b[int(rw)] -= (Rl + Rn)[ci, li]
                        acc = ref_rows[int(rw)]
                        for lj, cl in enumerate(dofs):
                            acc[int(cl)] = acc.get(int(cl), 0.0) + Je[ci, li, lj]
        for rw in rows:
            cols, vals = get_row(int(rw))
            ref = np.array([ref_rows[int(rw)].get(int(c), 0.0) for c in cols])
            if not set(ref_rows[int(rw)]) <= set(int(c) for c in cols):
                return float("inf"), float("inf"), 1.0
            scale = max(np.abs(ref).max(), 1e-300)
            worst_A = max(worst_A, float(np.abs(np.asarray(vals) - ref).max() / scale))
            worst_b = max(worst_b, abs(float(get_b(int(rw))) - ref_b[int(rw)]))
            scale_b = max(scale_b, abs(ref_b[int(rw)]))
    return worst_A, worst_b, scale_b
