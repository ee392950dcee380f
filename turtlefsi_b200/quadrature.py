"""Quadrature rules on the UFC reference simplex passed to the C-ABI (`tfsi_problem.qpts_*`).

The reference never states a rule: it passes a *degree* to FFC
(turtleFSI/problems/__init__.py:16-17 `quadrature_degree: 4`, problems/TF_fsi.py:61 degree 5,
problems/Test_Cylinder/problem_aneu.py:12 degree 6) and FFC asks FIAT
`create_quadrature(cell, degree, "default")` [dep, un-vendored]. `default_rule` reproduces
that selection: tabulated symmetric rules up to degree 6, collapsed Gauss-Jacobi with
(degree + 2) // 2 points per axis above. A caller that has FIAT can pass
FIAT's own points/weights through the ABI instead.
"""
import itertools

import numpy as np

_T3 = 1.0 / 3.0
_RULES = {
    (2, 1): [(1.0, (_T3, _T3, _T3))],
    (2, 2): [(_T3, (2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0))],
    (2, 3): [(1.0 / 6.0, (0.659027622374092, 0.231933368553031, 0.109039009072877))],
    (2, 4): [(0.109951743655322, (0.816847572980459, 0.091576213509771, 0.091576213509771)),
             (0.223381589678011, (0.108103018168070, 0.445948490915965, 0.445948490915965))],
    (2, 5): [(0.225, (_T3, _T3, _T3)),
             (0.12593918054482717, (0.79742698535308720, 0.10128650732345633, 0.10128650732345633)),
             (0.13239415278850616, (0.05971587178976981, 0.47014206410511505, 0.47014206410511505))],
    (2, 6): [(0.050844906370207, (0.873821971016996, 0.063089014491502, 0.063089014491502)),
             (0.116786275726379, (0.501426509658179, 0.249286745170910, 0.249286745170910)),
             (0.082851075618374, (0.636502499121399, 0.310352451033785, 0.053145049844816))],
    (3, 1): [(1.0, (0.25, 0.25, 0.25, 0.25))],
    (3, 2): [(0.25, (0.585410196624969,) + (0.138196601125011,) * 3)],
    (3, 3): [(-0.8, (0.25,) * 4), (0.45, (0.5,) + (1.0 / 6.0,) * 3)],
    (3, 4): [(0.0190476190476190, (0.0, 0.0, 0.5, 0.5)),
             (0.0885898247429807, (0.6984197043243866,) + (0.1005267652252045,) * 3),
             (0.1328387466855907, (0.0568813795204234,) + (0.3143728734931922,) * 3)],
    (3, 5): [(0.1817020685825351, (0.25,) * 4),
             (0.0361607142857143, (0.0,) + (_T3,) * 3),
             (0.0698714945161738, (0.7272727272727273,) + (0.0909090909090909,) * 3),
             (0.0656948493683187, (0.0665501535736643,) * 2 + (0.4334498464263357,) * 2)],
    (3, 6): [(0.0399227502581679, (0.3561913862225449,) + (0.2146028712591517,) * 3),
             (0.0100772110553207, (0.8779781243961660,) + (0.0406739585346113,) * 3),
             (0.0553571815436544, (0.0329863295731731,) + (0.3223378901422757,) * 3),
             (0.0482142857142857, (0.0636610018750175, 0.0636610018750175, 0.2696723314583159, 0.6030056647916491))],
}


def default_rule(gdim, degree):
    """(points [nq, gdim], weights [nq]); weights sum to the reference volume (1/2, 1/6)."""
    degree = max(int(degree), 1)
    vol = 0.5 if gdim == 2 else 1.0 / 6.0
    key = (gdim, degree)
    if key in _RULES:
        pts, wts = [], []
        for w, bary in _RULES[key]:
            for perm in sorted(set(itertools.permutations(bary))):
                pts.append(perm[1:])
                wts.append(w * vol)
        return np.ascontiguousarray(pts, dtype=np.float64), np.ascontiguousarray(wts, dtype=np.float64)
    from scipy.special import roots_jacobi
    m = (degree + 2) // 2
    axes = [roots_jacobi(m, float(a), 0.0) for a in range(gdim)]
    pts, wts = [], []
    for idx in itertools.product(range(m), repeat=gdim):
        r = [axes[a][0][idx[a]] for a in range(gdim)]
        w = np.prod([axes[a][1][idx[a]] for a in range(gdim)])
        if gdim == 2:
            x = 0.5 * (1 + r[0]) * (1 - r[1]) - 1
            pts.append((0.5 * (x + 1), 0.5 * (r[1] + 1)))
            wts.append(w / 8.0)
        else:
            x = 0.25 * (1 + r[0]) * (1 - r[1]) * (1 - r[2]) - 1
            y = 0.5 * (1 + r[1]) * (1 - r[2]) - 1
            pts.append((0.5 * (x + 1), 0.5 * (y + 1), 0.5 * (r[2] + 1)))
            wts.append(w / 64.0)
    return np.ascontiguousarray(pts, dtype=np.float64), np.ascontiguousarray(wts, dtype=np.float64)
