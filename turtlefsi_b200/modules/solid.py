"""turtleFSI/modules/solid.py counterpart: selects the solid terms assembled by the CUDA kernels
(solid.py:41-43, 62-68 St.Venant-Kirchhoff; zero forms of no_solid.py:14-15)."""


def solid_setup(**namespace):
    term = dict(solid="solid")
    return dict(F_solid_linear=term, F_solid_nonlinear=term)


def facet_terms(mesh, DVP, boundaries, theta, robin_bc=False, ds_s_ext_id_list=None, k_s=0.0, c_s=0.0, **namespace):
    """solid.py:76-86: Robin condition theta0 (k_s d + c_s v).psi ds_s + theta1 (k_s d^{n-1} + c_s v^{n-1}).psi ds_s.
    Linear in the state -> two fixed COO operators: `now` (A_pre += , b -= now x) and `prev` (b -= prev x1).
    Same argument checks as the reference (:79-81)."""
    if not robin_bc:
        return None
    from collections.abc import Iterable
    assert isinstance(k_s, Iterable), "k_s should be an iterable (e.g., list, tuple, etc.)."
    assert isinstance(c_s, Iterable), "c_s should be an iterable (e.g., list, tuple, etc.)."
    assert len(k_s) == len(c_s) == len(ds_s_ext_id_list), "k_s, c_s, and ds_s_ext_id_list should have the same length."
    from ..boundary import robin_facet_operators
    now, prev = robin_facet_operators(mesh, DVP, boundaries, ds_s_ext_id_list, k_s, c_s, theta)
    return dict(now=now, prev=prev)
