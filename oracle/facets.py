"""ORACLE (test infrastructure only): facet terms, written independently of turtlefsi_b200.boundary.
Robin condition (turtleFSI/modules/solid.py:83-84) and interior-facet normal load (problem_aneu.py:214):
CLOSED-FORM P2 facet mass matrices / integrals (no quadrature). Exterior-facet terms of the biharmonic lifting
(turtleFSI/modules/biharmonic.py:38-41):
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


# closed forms on a facet, local P2 order (vertices, then the facet's own edges):
#   edge (2-D), nodes (a, b, mid):              int phi_i phi_j = L/30 * _M_EDGE,  int phi_i = L * (1/6, 1/6, 2/3)
#   triangle (3-D), nodes (v0,v1,v2,m12,m02,m01): int phi_i phi_j = A/180 * _M_TRI, int phi_i = A * (0,0,0,1/3,1/3,1/3)
_M_EDGE = np.array([[4.0, -1.0, 2.0], [-1.0, 4.0, 2.0], [2.0, 2.0, 16.0]])
_M_TRI = np.array([[6.0, -1, -1, -4, 0, 0], [-1, 6, -1, 0, -4, 0], [-1, -1, 6, 0, 0, -4],
                   [-4, 0, 0, 32, 16, 16], [0, -4, 0, 16, 32, 16], [0, 0, -4, 16, 16, 32]])


def _facet_nodes(g, k):
    """cell-local P2 node ids living on the facet opposite to local vertex k, in the closed-form order."""
    loc = [v for v in range(g + 1) if v != k]
    edge_id = {tuple(e): g + 1 + i for i, e in enumerate(ft.EDGES[g])}
    if g == 2:
        return loc, loc + [edge_id[(loc[0], loc[1])]]
    return loc, loc + [edge_id[(loc[1], loc[2])], edge_id[(loc[0], loc[2])], edge_id[(loc[0], loc[1])]]


def _measure_normal(mesh, c, k, loc):
    g = mesh.gdim
    X = mesh.coords[mesh.cells[c]]
    Xf = X[loc]
    if g == 2:
        e = Xf[1] - Xf[0]
        meas = np.hypot(*e)
        n = np.array([e[1], -e[0]]) / meas
    else:
        cr = np.cross(Xf[1] - Xf[0], Xf[2] - Xf[0])
        meas = 0.5 * np.linalg.norm(cr)
        n = cr / (2.0 * meas)
    return meas, (-n if np.dot(n, X[k] - Xf[0]) > 0 else n)


def robin_matrices(mesh, space, facet_cell, facet_local, k_s, c_s, theta):
    """scipy CSR pair (now, prev) of solid.py:83-84 over the given exterior facets:
    rows psi; now = theta (k_s M | c_s M) on (d, v) columns, prev = (1 - theta) (k_s M | c_s M)."""
    import scipy.sparse as sp
    g = mesh.gdim
    n2 = ft.n_p2(g)
    N = space.num_dofs
    rows, cols, vals = [], [], []
    for c, k in zip(facet_cell, facet_local):
        loc, nodes = _facet_nodes(g, k)
        meas, _ = _measure_normal(mesh, c, k, loc)
        M = meas / 30.0 * _M_EDGE if g == 2 else meas / 180.0 * _M_TRI
        dofs = space.cell_dofs[c]
        for a in range(g):
            for i, ni in enumerate(nodes):
                for j, nj in enumerate(nodes):
                    rows += [dofs[(g + a) * n2 + ni]] * 2
                    cols += [dofs[a * n2 + nj], dofs[(g + a) * n2 + nj]]
                    vals += [k_s * M[i, j], c_s * M[i, j]]
    base = sp.coo_matrix((vals, (rows, cols)), shape=(N, N)).tocsr()
    return theta * base, (1.0 - theta) * base


def normal_load(mesh, space, facet_cell, facet_local):
    """dense vector l_i = int n . psi_i dS over the given (cell, local facet) pairs, n = outward normal of that cell."""
    g = mesh.gdim
    n2 = ft.n_p2(g)
    l = np.zeros(space.num_dofs)
    w = np.array([1 / 6, 1 / 6, 2 / 3]) if g == 2 else np.array([0, 0, 0, 1 / 3, 1 / 3, 1 / 3])
    for c, k in zip(facet_cell, facet_local):
        loc, nodes = _facet_nodes(g, k)
        meas, n = _measure_normal(mesh, c, k, loc)
        dofs = space.cell_dofs[c]
        for a in range(g):
            for i, ni in enumerate(nodes):
                l[dofs[(g + a) * n2 + ni]] += n[a] * meas * w[i]
    return l
