"""ORACLE (test infrastructure only — never imported by the product path).

Reference-element tables for the turtleFSI mixed space DVP = [P2^g, P2^g, P1 (, P2^g)]
(reference: turtleFSI/monolithic.py:77-92) and the FIAT "default" quadrature rules that
DOLFIN/FFC use when a quadrature degree is given (reference call sites:
turtleFSI/problems/__init__.py:16-17 degree 4, problems/TF_fsi.py:61 degree 5,
problems/Test_Cylinder/problem_aneu.py:12 degree 6).

Third-party origin: the arithmetic lives in FIAT/FFC/DOLFIN 2019.1 (un-vendored, absent from
/root/reference and from this image) — the conventions below restate their published
algorithms (UFC reference simplex, P2 dofs = vertices then edge midpoints in UFC edge order,
FIAT quadrature_schemes.py literal tables for degree <= 6, collapsed Gauss–Jacobi above).
PARITY STATUS: entry-level parity with DOLFIN is UNPINNED (no DOLFIN here); the end-to-end
scalars of tests/test_turtleFSI.py:50-53 pin the whole chain (see tests/test_turek_pin.py).
"""
import itertools

import numpy as np

# UFC edge numbering: edge k is opposite to ... (pairs of local vertices), FIAT/UFC order.
EDGES = {
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


def n_p2(g):
    return (g + 1) + len(EDGES[g])


def n_p1(g):
    return g + 1


def _bary(pts):
    """barycentric coordinates lam[nq, g+1] of reference points pts[nq, g] (lam0 = 1 - sum x)."""
    pts = np.asarray(pts)
    return np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)


def _dbary(g):
    """d lam_i / d xi_c, shape [g+1, g] (constant)."""
    D = np.zeros((g + 1, g))
    D[0, :] = -1.0
    for c in range(g):
        D[c + 1, c] = 1.0
    return D


def tabulate_p1(g, pts):
    """values [nq, g+1] and reference gradients [nq, g+1, g] of P1 Lagrange."""
    lam = _bary(pts)
    D = _dbary(g)
    return lam, np.broadcast_to(D, (lam.shape[0],) + D.shape).copy()


def tabulate_p2(g, pts):
    """values [nq, n2] and reference gradients [nq, n2, g] of P2 Lagrange.

    vertex i: lam_i (2 lam_i - 1); edge (i, j): 4 lam_i lam_j   (SURVEY Appendix B)."""
    lam = _bary(pts)
    D = _dbary(g)
    nq = lam.shape[0]
    n = n_p2(g)
    phi = np.zeros((nq, n))
    dphi = np.zeros((nq, n, g))
    for i in range(g + 1):
        phi[:, i] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
        dphi[:, i, :] = (4.0 * lam[:, i] - 1.0)[:, None] * D[i][None, :]
    for e, (i, j) in enumerate(EDGES[g]):
        phi[:, g + 1 + e] = 4.0 * lam[:, i] * lam[:, j]
        dphi[:, g + 1 + e, :] = 4.0 * (lam[:, i, None] * D[j][None, :] + lam[:, j, None] * D[i][None, :])
    return phi, dphi


def _orbit(bary):
    """all distinct permutations of a barycentric tuple, deterministic order."""
    return sorted(set(itertools.permutations(bary)), reverse=True)


def _from_orbits(g, orbits):
    vol = {2: 0.5, 3: 1.0 / 6.0}[g]
    P, W = [], []
    for w, bary in orbits:
        for b in _orbit(tuple(bary)):
            P.append(b[1:])       # cartesian = last g barycentrics (lam0 = 1 - sum)
            W.append(w * vol)
    return np.array(P, dtype=float), np.array(W, dtype=float)


def _gauss_jacobi_collapsed(g, m):
    """Collapsed (Stroud conical) Gauss–Jacobi rule with m points per axis on the UFC simplex
    (FIAT make_quadrature / CollapsedQuadrature{Triangle,Tetrahedron}Rule)."""
    from scipy.special import roots_jacobi
    if g == 2:
        x0, w0 = roots_jacobi(m, 0.0, 0.0)
        x1, w1 = roots_jacobi(m, 1.0, 0.0)
        P, W = [], []
        for a, wa in zip(x0, w0):
            for b, wb in zip(x1, w1):
                x = 0.5 * (1.0 + a) * (1.0 - b) - 1.0
                y = b
                P.append((0.5 * (x + 1.0), 0.5 * (y + 1.0)))
                W.append(wa * wb * 0.5 * 0.25)      # (1-b)/2 absorbed in Jacobi(1,0) weight: factor 1/2; map [-1,1]^2 tri (area 2) -> UFC (area 1/2): 1/4
        return np.array(P), np.array(W)
    x0, w0 = roots_jacobi(m, 0.0, 0.0)
    x1, w1 = roots_jacobi(m, 1.0, 0.0)
    x2, w2 = roots_jacobi(m, 2.0, 0.0)
    P, W = [], []
    for a, wa in zip(x0, w0):
        for b, wb in zip(x1, w1):
            for c, wc in zip(x2, w2):
                x = 0.25 * (1.0 + a) * (1.0 - b) * (1.0 - c) - 1.0
                y = 0.5 * (1.0 + b) * (1.0 - c) - 1.0
                z = c
                P.append((0.5 * (x + 1.0), 0.5 * (y + 1.0), 0.5 * (z + 1.0)))
                W.append(wa * wb * wc * 0.125 * 0.125)
        # weights: (1-b)/2 -> 1/2, ((1-c)/2)^2 -> 1/4 ; [-1,1]^3 tet (vol 4/3) -> UFC (1/6): 1/8
    return np.array(P), np.array(W)


_TRI = {
    1: [(1.0, (1 / 3, 1 / 3, 1 / 3))],
    2: [(1 / 3, (2 / 3, 1 / 6, 1 / 6))],
    3: [(1 / 6, (0.659027622374092, 0.231933368553031, 0.109039009072877))],
    4: [(0.109951743655322, (0.816847572980459, 0.091576213509771, 0.091576213509771)),
        (0.223381589678011, (0.108103018168070, 0.445948490915965, 0.445948490915965))],
    5: [(0.225, (1 / 3, 1 / 3, 1 / 3)),
        (0.12593918054482717, (0.79742698535308720, 0.10128650732345633, 0.10128650732345633)),
        (0.13239415278850616, (0.05971587178976981, 0.47014206410511505, 0.47014206410511505))],
    6: [(0.050844906370207, (0.873821971016996, 0.063089014491502, 0.063089014491502)),
        (0.116786275726379, (0.501426509658179, 0.249286745170910, 0.249286745170910)),
        (0.082851075618374, (0.636502499121399, 0.310352451033785, 0.053145049844816))],
}
_TET = {
    1: [(1.0, (0.25, 0.25, 0.25, 0.25))],
    2: [(0.25, (0.585410196624969, 0.138196601125011, 0.138196601125011, 0.138196601125011))],
    3: [(-0.8, (0.25, 0.25, 0.25, 0.25)), (0.45, (0.5, 1 / 6, 1 / 6, 1 / 6))],
    4: [(0.0190476190476190, (0.0, 0.0, 0.5, 0.5)),
        (0.0885898247429807, (0.6984197043243866, 0.1005267652252045, 0.1005267652252045, 0.1005267652252045)),
        (0.1328387466855907, (0.0568813795204234, 0.3143728734931922, 0.3143728734931922, 0.3143728734931922))],
    5: [(0.1817020685825351, (0.25, 0.25, 0.25, 0.25)),
        (0.0361607142857143, (0.0, 1 / 3, 1 / 3, 1 / 3)),
        (0.0698714945161738, (0.7272727272727273, 0.0909090909090909, 0.0909090909090909, 0.0909090909090909)),
        (0.0656948493683187, (0.0665501535736643, 0.0665501535736643, 0.4334498464263357, 0.4334498464263357))],
    6: [(0.0399227502581679, (0.3561913862225449, 0.2146028712591517, 0.2146028712591517, 0.2146028712591517)),
        (0.0100772110553207, (0.8779781243961660, 0.0406739585346113, 0.0406739585346113, 0.0406739585346113)),
        (0.0553571815436544, (0.0329863295731731, 0.3223378901422757, 0.3223378901422757, 0.3223378901422757)),
        (0.0482142857142857, (0.0636610018750175, 0.0636610018750175, 0.2696723314583159, 0.6030056647916491))],
}


def quadrature(g, degree):
    """FIAT create_quadrature(cell, degree, "default"): points [nq, g] on the UFC simplex and
    weights [nq] (sum = reference volume). SURVEY Appendix A."""
    degree = max(int(degree), 1)
    table = _TRI if g == 2 else _TET
    if degree in table:
        return _from_orbits(g, table[degree])
    m = (degree + 2) // 2
    return _gauss_jacobi_collapsed(g, m)
