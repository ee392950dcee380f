"""P1/P2 Lagrange tabulation on the UFC reference simplex (host side, numpy) — used by the
post-processing functionals and point evaluation; the kernels carry their own tabulation.
Conventions as DOLFIN/FIAT [dep]: P2 dofs = vertices then edge midpoints in UFC edge order."""
import numpy as np

from .space import EDGES


def tabulate(gdim, pts):
    """pts [n, g] -> (phi2 [n, n2], dphi2 [n, n2, g], phi1 [n, g+1])."""
    pts = np.atleast_2d(np.asarray(pts, dtype=float))
    g = gdim
    lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)
    D = np.zeros((g + 1, g))
    D[0] = -1.0
    D[1:] = np.eye(g)
    n2 = g + 1 + len(EDGES[g])
    phi = np.zeros((len(pts), n2))
    dphi = np.zeros((len(pts), n2, g))
    for i in range(g + 1):
        phi[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
        dphi[:, i] = (4 * lam[:, i] - 1)[:, None] * D[i]
    for e, (a, b) in enumerate(EDGES[g]):
        phi[:, g + 1 + e] = 4 * lam[:, a] * lam[:, b]
        dphi[:, g + 1 + e] = 4 * (lam[:, a, None] * D[b] + lam[:, b, None] * D[a])
    return phi, dphi, lam
