"""turtleFSI/modules/common.py counterpart for the HOST side: the kinematic and stress helpers problem files use in
their post-processing functionals (`sigma(v, p, d, mu_f)` in TF_fsi.py:229-233, TF_cfd.py:129), written in the form
language of turtlefsi_b200.ufl_lite. The kernels carry their own per-point versions of the same tensors
(csrc/tfsi_math.cuh); these definitions are the reference's (common.py:31-109), restated."""
from ..ufl_lite import Identity, as_expr, grad, inv, det, tr

__all__ = ["get_dimension", "F_", "L_", "D_", "J_", "eps", "sigma_f_u", "sigma_f_p", "sigma", "E", "S", "Piola1",
           "S_linear"]


def get_dimension(u):
    u = as_expr(u)
    return getattr(u, "gdim", None) or (u.shape[0] if u.shape else 3)


def F_(d):                                                   # deformation gradient, common.py:31-35
    return Identity(get_dimension(d)) + grad(d)


def L_(v):
    return grad(v)


def D_(v):
    return 0.5 * (grad(v) + grad(v).T)


def J_(d):                                                   # common.py:49-53
    return det(F_(d))


def eps(d):
    return 0.5 * (grad(d) * inv(F_(d)) + inv(F_(d)).T * grad(d).T)


def sigma_f_u(u, d, mu_f):                                   # common.py:63-67
    return mu_f * (grad(u) * inv(F_(d)) + inv(F_(d)).T * grad(u).T)


def sigma_f_p(p, u):                                         # common.py:70-74
    return -p * Identity(get_dimension(u))


def sigma(u, p, d, mu_f):                                    # common.py:77-81
    return sigma_f_u(u, d, mu_f) + sigma_f_p(p, u)


def E(d):                                                    # Green-Lagrange strain, common.py:84-88
    return 0.5 * (F_(d).T * F_(d) - Identity(get_dimension(d)))


def Piola1(d, solid_properties):                             # common.py:104-109 (St.Venant-Kirchhoff only on this path)
    if solid_properties["material_model"] != "StVenantKirchoff":
        raise RuntimeError("material model '%s' is not available on the B200 path" % solid_properties["material_model"])
    I = Identity(get_dimension(d))
    S_svk = 2 * solid_properties["mu_s"] * E(d) + solid_properties["lambda_s"] * tr(E(d)) * I
    return F_(d) * S_svk


def S(d, solid_properties):                                  # second Piola-Kirchhoff, common.py:91-98
    return inv(F_(d)) * Piola1(d, solid_properties)


def S_linear(d, alfa_mu, alfa_lam):                          # common.py:141-145
    return alfa_lam * tr(eps(d)) * Identity(get_dimension(d)) + 2.0 * alfa_mu * eps(d)
