"""turtleFSI/modules/common.py counterpart for the HOST side: the kinematic and stress helpers problem files use in
their post-processing functionals (`sigma(v, p, d, mu_f)` in TF_fsi.py:229-233, TF_cfd.py:129), written in the form
language of turtlefsi_b200.ufl_lite. The kernels carry their own per-point versions of the same tensors
(csrc/tfsi_math.cuh); these definitions are the reference's (common.py:31-109), restated."""
from ..ufl_lite import Identity, as_expr, grad, inv, det, tr

__all__ = ["get_dimension", "F_", "L_", "D_", "J_", "eps", "sigma_f_u", "sigma_f_p", "sigma", "E", "S", "Piola1",
           "S_linear", "Svisc_D"]


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


def Piola1(d, solid_properties):                             # common.py:100-127
    """First Piola-Kirchhoff stress. StVenantKirchoff in the reference's explicit form (:104-109); the energy-based
    models (:110-125, P = diff(W, F) of the densities :214-273) in the closed form of that derivative — this small form
    language has no `variable` / `diff`:
      dI1/dF = 2 F,  dI2/dF = 2 (I1 F - F C),  dJ/dF = J F^-T,  I1 = tr C, I2 = (I1^2 - tr(C C)) / 2, C = F^T F."""
    model = solid_properties["material_model"]
    g = get_dimension(d)
    I = Identity(g)
    F = F_(d)
    lambda_s, mu_s = solid_properties["lambda_s"], solid_properties["mu_s"]
    if model in ("StVenantKirchoff", "StVenantKirchoffEnergy"):
        S_svk = 2 * mu_s * E(d) + lambda_s * tr(E(d)) * I
        return F * S_svk
    J = det(F)
    FinvT = inv(F).T
    C = F.T * F
    I1 = tr(C)
    if model == "NeoHookean":                                # W = C1 (I1 - dim - 2 ln J) + D1 (J - 1)^2, :228-240
        return mu_s * (F - FinvT) + lambda_s * (J - 1) * J * FinvT
    if model == "Gent":                                      # W = -mu/2 (Jm ln(1 - (I1 - 3)/Jm) + 2 ln J), :242-256
        Jm = solid_properties["Jm"]
        return (mu_s / (1 - (I1 - 3) / Jm)) * F - mu_s * FinvT
    if model == "MooneyRivlin":                              # :258-273
        C01, C10, C11 = solid_properties["C01"], solid_properties["C10"], solid_properties["C11"]
        K = lambda_s + 2 * mu_s / 3
        I2 = 0.5 * (I1 ** 2 - tr(C * C))
        Ibar1, Ibar2 = J ** (-2.0 / 3.0) * I1, J ** (-4.0 / 3.0) * I2
        dIbar1 = J ** (-2.0 / 3.0) * (2 * F - (2.0 / 3.0) * I1 * FinvT)
        dIbar2 = J ** (-4.0 / 3.0) * (2 * I1 * F - 2 * F * C - (4.0 / 3.0) * I2 * FinvT)
        return (C01 + C11 * (Ibar1 - 3)) * dIbar2 + (C10 + C11 * (Ibar2 - 3)) * dIbar1 + K * (J - 1) * J * FinvT
    raise RuntimeError('Invalid entry for material_model, choose from ["StVenantKirchoff",'
                       '"StVenantKirchoffEnergy","NeoHookean","MooneyRivlin","Gent"]')


def Svisc_D(v, solid_properties):                            # common.py:129-139
    I = Identity(get_dimension(v))
    return solid_properties["mu_visc_s"] * tr(D_(v)) * I + 2 * solid_properties["lambda_visc_s"] * D_(v)


def S(d, solid_properties):                                  # second Piola-Kirchhoff, common.py:91-98
    return inv(F_(d)) * Piola1(d, solid_properties)


def S_linear(d, alfa_mu, alfa_lam):                          # common.py:141-145
    return alfa_lam * tr(eps(d)) * Identity(get_dimension(d)) + 2.0 * alfa_mu * eps(d)
