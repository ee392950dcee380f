"""ORACLE (test infrastructure only): exterior-facet terms of the biharmonic lifting
(turtleFSI/modules/biharmonic.py:38-41), written independently of turtlefsi_b200.boundary:
the P2 x P1 facet integrand is integrated with Simpson's rule on edges (exact to degree 3) and with the
degree-6 triangle rule of oracle/fem_tables.py on faces; basis functions from oracle/fem_tables.py."""
import numpy as np

from . import fem_tables as ft


def biharmonic_facet_matrix(mesh, space, facet_vertices, facet_cell, facet_local, alpha_u=0.01):
    """dense-free COO (rows, cols, vals) for the given exterior facets (arrays: facet -> vertices [nf, g],
    incident cell, local index of the vertex opposite to the facet)."""
    g = mesh.gdim
    n2 = ft.n_p2(g)
    ow = 2 * g * n2 + g + 1
    if g == 2:
        bary = np.array([[1.0, 0.0], [0.5, 0.5], [0.0, 1.0]])
        wts = np.array([1.0, 4.0, 1.0]) / 6.0
    else:
        P, W = ft.quadrature(2, 6)
        bary = np.concatenate([1.0 - P.sum(axis=1, keepdims=True), P], axis=1)
        wts = W / W.sum()
    refv = np.vstack([np.zeros(g), np.eye(g)])
    rows, cols, vals = [], [], []
    for f in range(len(facet_cell)):
        c, k = facet_cell[f], facet_local[f]
        loc = [v for v in range(g + 1) if v != k]
        X = mesh.coords[mesh.cells[c]]
        Jg = (X[1:] - X[0]).T
        K = np.linalg.inv(Jg)
        Xf = X[loc]
        if g == 2:
            e = Xf[1] - Xf[0]
            meas = np.hypot(*e)
            n = np.array([e[1], -e[0]]) / meas
        else:
            cr = np.cross(Xf[1] - Xf[0], Xf[2] - Xf[0])
            meas = 0.5 * np.linalg.norm(cr)
            n = cr / np.linalg.norm(cr)
        if np.dot(n, X[k] - Xf[0]) > 0:
            n = -n
        pts = bary @ refv[loc]
        phi, dphi = ft.tabulate_p2(g, pts)
        dn = (dphi @ K) @ n
        M = alpha_u * meas * (phi * wts[:, None]).T @ dn
        dofs = space.cell_dofs[c]
        for a in range(g):
            wd, dd = dofs[ow + a * n2: ow + (a + 1) * n2], dofs[a * n2:(a + 1) * n2]
            for i in range(n2):
                for j in range(n2):
                    rows += [wd[i], dd[i]]
                    cols += [dd[j], wd[j]]
                    vals += [M[i, j], -M[i, j]]
    return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64), np.array(vals)
