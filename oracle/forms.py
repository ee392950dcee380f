"""ORACLE (test infrastructure only — never imported by the product path).

numpy float64/complex128 restatement of turtleFSI's per-cell weak forms, written directly
from the UFL expressions of the reference:

  fluid      turtleFSI/modules/fluid.py:32-56      (N1..N6 non-linear, L1, L2, L4 "linear")
  solid      turtleFSI/modules/solid.py:29-68      (S1-S3, S5 linear; S4 non-linear; gravity; viscoelasticity :46-59;
             material models common.py:100-127, 214-273)
  laplace    turtleFSI/modules/laplace.py:33-46    (4 alfa sub-types)
  biharmonic turtleFSI/modules/biharmonic.py:31-36 (cell terms)
  elastic    turtleFSI/modules/elastic.py:18-27
  helpers    turtleFSI/modules/common.py:31-35 (F_), 49-53 (J_), 63-74 (sigma_f_u/p),
             84-88 (E), 104-109 (Piola1, SVK), 56-60 (eps), 141-145 (S_linear)
  split      turtleFSI/modules/newtonsolver.py:18-24 (F_lin / F_nonlin, derivative wrt dvp_["n"])

The Jacobian is obtained by COMPLEX-STEP differentiation of these residuals (an independent
route from the hand-derived Gateaux derivative used by the CUDA kernels and by oracle/c).
Element/quadrature conventions: oracle/fem_tables.py ([dep] FIAT/FFC/DOLFIN 2019.1).
PARITY STATUS: entry-level parity with DOLFIN unpinned; see fem_tables.py header.

Local dof layout of one cell: [d_0(n2) .. d_{g-1}, v_0 .. v_{g-1}, p(n1) (, w_0 ..)].
"""
import numpy as np

from . import fem_tables as ft


class Space:
    """local layout of DVP on one cell."""

    def __init__(self, gdim, biharmonic=False):
        self.g = g = gdim
        self.n2 = ft.n_p2(g)
        self.n1 = ft.n_p1(g)
        self.biharmonic = biharmonic
        self.off_d = 0
        self.off_v = g * self.n2
        self.off_p = 2 * g * self.n2
        self.off_w = self.off_p + self.n1
        self.ndof = self.off_w + (g * self.n2 if biharmonic else 0)


def geometry(X):
    """X[nc, g+1, g] vertex coordinates (vertices sorted by global index, DOLFIN ordered mesh).
    returns K = J_geo^-1 [nc,g,g] and |det J_geo| [nc]."""
    Jg = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)        # Jg[c, b, k] = d x_b / d xi_k
    det = np.linalg.det(Jg)
    return np.linalg.inv(Jg), np.abs(det)


def _inv_det(F):
    return np.linalg.inv(F), np.linalg.det(F)


def strain_energy(Fm, model, prm):
    """W(F) of turtleFSI/modules/common.py:214-273, transcribed literally (incl. the hard-coded 3s of Gent and
    MooneyRivlin and get_dimension(F) of NeoHookean). Fm[..., g, g]."""
    g = Fm.shape[-1]
    lam, mu = prm['lambda_s'], prm['mu_s']
    FT = np.swapaxes(Fm, -1, -2)
    tr = lambda A: np.trace(A, axis1=-2, axis2=-1)
    J = np.linalg.det(Fm)
    if model == 'StVenantKirchoffEnergy':                                  # :214-225
        E_ = 0.5 * (FT @ Fm - np.eye(g))
        return lam / 2 * tr(E_) ** 2 + mu * tr(E_ @ E_)
    if model == 'NeoHookean':                                              # :228-240
        C1, D1 = mu / 2, lam / 2
        return C1 * (tr(FT @ Fm) - g - 2 * np.log(J)) + D1 * (J - 1) ** 2
    if model == 'Gent':                                                    # :242-256
        Jm = prm['Jm']
        return -(mu / 2) * (Jm * np.log(1 - (tr(Fm @ FT) - 3) / Jm) + 2 * np.log(J))
    if model == 'MooneyRivlin':                                            # :258-273
        Kb = lam + 2 * mu / 3
        D1 = 2 / Kb
        B = Fm @ FT
        I1 = tr(B)
        I2 = 0.5 * (tr(B) ** 2 - tr(B @ B))
        Ibar1, Ibar2 = J ** (-2 / 3) * I1, J ** (-4 / 3) * I2
        return (prm['C01'] * (Ibar2 - 3) + prm['C10'] * (Ibar1 - 3) + prm['C11'] * (Ibar2 - 3) * (Ibar1 - 3)
                + (1 / D1) * (J - 1) ** 2)
    raise RuntimeError('Invalid entry for material_model')


def piola1(Fm, model, prm):
    """First Piola-Kirchhoff stress, common.py:100-127. StVenantKirchoff: the explicit form :104-109. The other models
    are P = diff(W, F) (:110-125); here the derivative is written out by hand in terms of F
    (tests/test_oracle_forms.py checks it against a complex-step derivative of `strain_energy`). Works on complex F."""
    g = Fm.shape[-1]
    I = np.eye(g)
    lam, mu = prm['lambda_s'], prm['mu_s']
    FT = np.swapaxes(Fm, -1, -2)
    tr = lambda A: np.trace(A, axis1=-2, axis2=-1)[..., None, None]
    if model in ('StVenantKirchoff', 'StVenantKirchoffEnergy'):            # common.py:84-88,104-109
        E = 0.5 * (FT @ Fm - I)
        return Fm @ (2.0 * mu * E + lam * tr(E) * I)
    FiT = np.swapaxes(np.linalg.inv(Fm), -1, -2)
    J = np.linalg.det(Fm)[..., None, None]
    C = FT @ Fm
    I1 = tr(C)
    if model == 'NeoHookean':
        return mu * (Fm - FiT) + lam * (J - 1) * J * FiT
    if model == 'Gent':
        return mu / (1 - (I1 - 3) / prm['Jm']) * Fm - mu * FiT
    if model == 'MooneyRivlin':
        Kb = lam + 2 * mu / 3
        I2 = 0.5 * (I1 ** 2 - tr(C @ C))
        Ib1, Ib2 = J ** (-2 / 3) * I1, J ** (-4 / 3) * I2
        dIb1 = J ** (-2 / 3) * (2 * Fm - (2 / 3) * I1 * FiT)
        dIb2 = J ** (-4 / 3) * (2 * I1 * Fm - 2 * Fm @ C - (4 / 3) * I2 * FiT)
        return ((prm['C01'] + prm['C11'] * (Ib1 - 3)) * dIb2 + (prm['C10'] + prm['C11'] * (Ib2 - 3)) * dIb1
                + Kb * (J - 1) * J * FiT)
    raise RuntimeError('Invalid entry for material_model, choose from ["StVenantKirchoff",'
                       '"StVenantKirchoffEnergy","NeoHookean","MooneyRivlin","Gent"]')


class CellForms:
    """Evaluates F_linear / F_nonlinear cell vectors for a batch of cells of one region."""

    def __init__(self, gdim, biharmonic, degree):
        self.sp = Space(gdim, biharmonic)
        g = gdim
        self.pts, self.wts = ft.quadrature(g, degree)
        self.phi2, self.dphi2 = ft.tabulate_p2(g, self.pts)
        self.phi1, self.dphi1 = ft.tabulate_p1(g, self.pts)

    # ---- field evaluation -------------------------------------------------------------
    def _vec(self, u, off, K):
        sp = self.sp
        g, n2 = sp.g, sp.n2
        c = u[:, off:off + g * n2].reshape(-1, g, n2)                      # [nc, a, i]
        val = np.einsum('cai,qi->cqa', c, self.phi2)
        gref = np.einsum('cai,qik->cqak', c, self.dphi2)                   # d u_a / d xi_k
        grad = np.einsum('cqak,ckb->cqab', gref, K)                        # d u_a / d x_b
        return val, grad

    def _p(self, u):
        sp = self.sp
        return np.einsum('ci,qi->cq', u[:, sp.off_p:sp.off_p + sp.n1], self.phi1)

    # ---- testing ----------------------------------------------------------------------
    def _test_vec(self, R, off, K, wdet, fval=None, fgrad=None):
        """R[:, off + a*n2 + i] += sum_q wdet * (fval_a phi_i + fgrad_ab d_b phi_i)."""
        sp = self.sp
        g, n2 = sp.g, sp.n2
        out = np.zeros((R.shape[0], g, n2), dtype=R.dtype)
        if fval is not None:
            out = out + np.einsum('cqa,qi->cai', wdet[:, :, None] * fval, self.phi2)
        if fgrad is not None:
            # d phi_i / d x_b = K^T ghat_i ; contract the flux with K first (cheap), then with ghat
            fref = np.einsum('cqab,ckb->cqak', wdet[:, :, None, None] * fgrad, K)
            out = out + np.einsum('cqak,qik->cai', fref, self.dphi2)
        R[:, off:off + g * n2] += out.reshape(R.shape[0], g * n2)

    def _test_p(self, R, wdet, f):
        sp = self.sp
        R[:, sp.off_p:sp.off_p + sp.n1] += np.einsum('cq,qi->ci', wdet * f, self.phi1)

    # ---- the forms ---------------------------------------------------------------------
    def residual(self, kind, X, u, u1, prm):
        """returns (F_lin[nc, ndof], F_nonlin[nc, ndof]) for cells of `kind` in
        {"fluid", "solid"}; u, u1 = local dofs at time n, n-1; prm = dict of parameters:
        dt, theta, rho_f, mu_f, rho_s, mu_s, lambda_s, extrapolation, extrapolation_sub_type,
        hmin, gravity, fluid ("fluid"|"no_fluid"), solid ("solid"|"no_solid")."""
        sp = self.sp
        g = sp.g
        K, adet = geometry(X)
        dtype = np.result_type(u.dtype, u1.dtype, np.float64)
        nc = u.shape[0]
        Rl = np.zeros((nc, sp.ndof), dtype=dtype)
        Rn = np.zeros((nc, sp.ndof), dtype=dtype)
        wdet = adet[:, None] * self.wts[None, :]
        I = np.eye(g)
        k = prm['dt']
        th0, th1 = prm['theta'], 1.0 - prm['theta']

        d, Gd = self._vec(u, sp.off_d, K)
        v, Gv = self._vec(u, sp.off_v, K)
        d1, Gd1 = self._vec(u1, sp.off_d, K)
        v1, Gv1 = self._vec(u1, sp.off_v, K)
        p = self._p(u)
        F = I + Gd
        F1 = I + Gd1

        if kind == 'fluid' and prm.get('fluid', 'fluid') == 'fluid':
            rho, mu = prm['rho_f'], prm['mu_f']
            Fi, J = _inv_det(F)
            Fi1, J1 = _inv_det(F1)
            FiT = np.swapaxes(Fi, -1, -2)
            Fi1T = np.swapaxes(Fi1, -1, -2)
            # fluid.py:32  temporal derivative, non-linear part
            self._test_vec(Rn, sp.off_v, K, wdet, fval=rho / k * th0 * J[..., None] * (v - v1))
            # fluid.py:33  temporal derivative, "linear" part (J of d^{n-1})
            self._test_vec(Rl, sp.off_v, K, wdet, fval=rho / k * th1 * J1[..., None] * (v - v1))
            # fluid.py:36-37 convection
            H = Gv @ Fi
            self._test_vec(Rn, sp.off_v, K, wdet,
                           fval=th0 * rho * J[..., None] * np.einsum('cqab,cqb->cqa', H, v))
            # fluid.py:38-39
            H1 = Gv1 @ Fi1
            self._test_vec(Rl, sp.off_v, K, wdet,
                           fval=th1 * rho * J1[..., None] * np.einsum('cqab,cqb->cqa', H1, v1))
            # fluid.py:42-43 pressure stress  J (-p I) F^-T : grad psi
            self._test_vec(Rn, sp.off_v, K, wdet, fgrad=(-p * J)[..., None, None] * FiT)
            # fluid.py:46-47 viscous stress (common.py:63-67)
            sig = mu * (H + np.swapaxes(H, -1, -2))
            self._test_vec(Rn, sp.off_v, K, wdet, fgrad=th0 * J[..., None, None] * (sig @ FiT))
            # fluid.py:48-49
            sig1 = mu * (H1 + np.swapaxes(H1, -1, -2))
            self._test_vec(Rl, sp.off_v, K, wdet, fgrad=th1 * J1[..., None, None] * (sig1 @ Fi1T))
            # fluid.py:52 div(J F^-1 v) gamma ; Piola identity: div(adj F) = 0 cell-wise, so
            # div(J F^-1 v) = tr(J F^-1 grad v)   (SURVEY 8a note on N5; checked in
            # tests/test_oracle_forms.py::test_piola_identity against the literal expression)
            self._test_p(Rn, wdet, J * np.einsum('cqab,cqba->cq', Fi, Gv))
            # fluid.py:55-56 ALE convection
            self._test_vec(Rn, sp.off_v, K, wdet,
                           fval=-rho / k * J[..., None] * np.einsum('cqab,cqb->cqa', H, d - d1))

        if kind == 'fluid':
            ext = prm.get('extrapolation', 'laplace')
            if ext == 'laplace':
                sub = prm.get('extrapolation_sub_type', 'constant')
                if sub == 'volume_change':
                    alfa = 1.0 / np.linalg.det(F)                          # laplace.py:33-34
                elif sub == 'volume':
                    vol = adet / {2: 2.0, 3: 6.0}[g]
                    alfa = (1.0 / vol)[:, None] * np.ones(len(self.wts))   # laplace.py:35-36
                elif sub == 'small_constant':
                    alfa = 0.01 * prm['hmin'] ** 2 * np.ones((nc, len(self.wts)))  # :37-38
                elif sub == 'constant':
                    alfa = np.ones((nc, len(self.wts)))                    # :39-40
                else:
                    raise RuntimeError("Could not find extrapolation method {}".format(sub))
                self._test_vec(Rl, sp.off_d, K, wdet, fgrad=alfa[..., None, None] * Gd)   # :45
            elif ext == 'biharmonic':
                au = 0.01                                                  # biharmonic.py:31
                w, Gw = self._vec(u, sp.off_w, K)
                self._test_vec(Rl, sp.off_w, K, wdet, fval=au * w, fgrad=-au * Gd)      # :35
                self._test_vec(Rl, sp.off_d, K, wdet, fgrad=au * Gw)                    # :36
            elif ext == 'elastic':
                vol = adet / {2: 2.0, 3: 6.0}[g]
                E_y = (1.0 / vol)[:, None, None, None]                     # elastic.py:18
                nu = 0.25
                a_lam = nu * E_y / ((1.0 + nu) * (1.0 - 2.0 * nu))
                a_mu = E_y / (2.0 * (1.0 + nu))
                Fi, J = _inv_det(F)
                FiT = np.swapaxes(Fi, -1, -2)
                eps = 0.5 * (Gd @ Fi + FiT @ np.swapaxes(Gd, -1, -2))      # common.py:56-60
                Sl = a_lam * np.trace(eps, axis1=-2, axis2=-1)[..., None, None] * I + 2.0 * a_mu * eps
                self._test_vec(Rl, sp.off_d, K, wdet, fgrad=J[..., None, None] * (Sl @ FiT))  # :24-25
            elif ext == 'no_extrapolation':
                pass
            else:
                raise RuntimeError(ext)

        if kind == 'solid' and prm.get('solid', 'solid') == 'solid':
            rho, mu, lam = prm['rho_s'], prm['mu_s'], prm['lambda_s']
            delta = 1.0e7                                                  # solid.py:29
            # solid.py:41-43
            self._test_vec(Rl, sp.off_v, K, wdet, fval=rho / k * (v - v1))
            self._test_vec(Rl, sp.off_d, K, wdet,
                           fval=delta * rho / k * (d - d1) - delta * rho * (th0 * v + th1 * v1))

            model = prm.get('material_model', 'StVenantKirchoff')

            def piola(Fm):                                                 # common.py:100-127
                return piola1(Fm, model, prm)
            self._test_vec(Rn, sp.off_v, K, wdet, fgrad=th0 * piola(F))    # solid.py:62
            self._test_vec(Rl, sp.off_v, K, wdet, fgrad=th1 * piola(F1))   # solid.py:63
            visco = prm.get('viscoelasticity')
            if visco in ('Form1', 'Form2'):                                # solid.py:46-55

                def svisc(Gw):                                             # common.py:129-139, D_ :43-47
                    D = 0.5 * (Gw + np.swapaxes(Gw, -1, -2))
                    return (prm['mu_visc_s'] * np.trace(D, axis1=-2, axis2=-1)[..., None, None] * I
                            + 2.0 * prm['lambda_visc_s'] * D)
                if visco == 'Form1':                                       # :48-50
                    self._test_vec(Rn, sp.off_v, K, wdet, fgrad=th0 * (F @ svisc(Gv)))
                    self._test_vec(Rl, sp.off_v, K, wdet, fgrad=th1 * (F1 @ svisc(Gv1)))
                else:                                                      # :51-55
                    self._test_vec(Rn, sp.off_v, K, wdet, fgrad=(1.0 / k) * (F @ (th0 * svisc(Gd - Gd1))))
                    self._test_vec(Rl, sp.off_v, K, wdet, fgrad=(1.0 / k) * (F1 @ (th1 * svisc(Gd - Gd1))))
            if prm.get('gravity') is not None:                            # solid.py:65-68
                gv = np.zeros(g)
                gv[1] = -prm['gravity'] * rho
                self._test_vec(Rl, sp.off_v, K, wdet,
                               fval=-np.broadcast_to(gv, v.shape).astype(dtype))
        return Rl, Rn

    def jacobian(self, kind, X, u, u1, prm, part='nl', h=1e-30):
        """complex-step derivative of F_lin ('lin') / F_nonlin ('nl') / both ('all') wrt the
        local dofs at time n (newtonsolver.py:22-24). returns [nc, ndof, ndof]."""
        sp = self.sp
        nc = u.shape[0]
        Jm = np.zeros((nc, sp.ndof, sp.ndof))
        uc = u.astype(np.complex128)
        for j in range(sp.ndof):
            uc[:, j] += 1j * h
            Rl, Rn = self.residual(kind, X, uc, u1, prm)
            R = {'lin': Rl, 'nl': Rn, 'all': Rl + Rn}[part]
            Jm[:, :, j] = R.imag / h
            uc[:, j] = u[:, j]
        return Jm
