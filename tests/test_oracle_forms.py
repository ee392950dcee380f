"""CPU tests of the oracle itself and of the kernels' per-point mathematics (host-compiled).

The reference holds NO golden vector at matrix/residual-entry level; what can be
pinned on the CPU is internal consistency between three independent derivations:
  (1) numpy residual forms written from the UFL + complex-step Jacobian   (oracle/forms.py)
  (2) plain-C FFC-style restatement with hand-derived Jacobian           (oracle/c)
  (3) the CUDA kernels' C'-tensor / sum-factorised formulation, host-compiled (tests/emul)
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import assembly as oa
from oracle import cport
from oracle import fem_tables as ft
from oracle.forms import CellForms
from tests.conftest import TUBE_PARAMS, TUBE_REGIONS
from tests.helpers import square_mesh
from turtlefsi_b200.mesh import synthetic_state, tube_mesh
from turtlefsi_b200.space import MixedSpace

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libtfsi_emul.so")
    src = os.path.join(HERE, "emul", "emul.cpp")
    hdr = os.path.join(HERE, "..", "turtlefsi_b200", "csrc", "tfsi_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return C.CDLL(so)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _case(g, bih, amp):
    mesh = tube_mesh(3) if g == 3 else square_mesh(5)
    sp = MixedSpace(mesh, biharmonic=bih)
    x = synthetic_state(mesh, sp, 2024, h=amp)
    x1 = synthetic_state(mesh, sp, 2025, h=amp)
    return mesh, sp, x, x1


def test_quadrature_rules_exact():
    import itertools
    import math
    for g in (2, 3):
        for deg in range(1, 11):
            P, W = ft.quadrature(g, deg)
            assert abs(W.sum() - (0.5 if g == 2 else 1 / 6)) < 1e-14
            for e in itertools.product(range(deg + 1), repeat=g):
                if sum(e) > deg:
                    continue
                exact = np.prod([math.factorial(k) for k in e]) / math.factorial(sum(e) + g)
                assert abs((W * np.prod(P ** np.array(e), axis=1)).sum() - exact) < 2e-14 * max(exact, 1e-3)
    assert len(ft.quadrature(3, 4)[1]) == 14 and len(ft.quadrature(3, 6)[1]) == 24
    assert len(ft.quadrature(2, 5)[1]) == 7 and len(ft.quadrature(2, 10)[1]) == 36


def test_product_rule_equals_oracle_rule():
    from turtlefsi_b200.quadrature import default_rule
    for g in (2, 3):
        for d in range(1, 12):
            p, w = default_rule(g, d)
            P, W = ft.quadrature(g, d)
            a = np.lexsort(np.c_[p, w].T)
            b = np.lexsort(np.c_[P, W].T)
            assert np.abs(np.c_[p, w][a] - np.c_[P, W][b]).max() < 1e-15


def test_p2_basis_is_nodal_and_partition_of_unity():
    for g in (2, 3):
        verts = np.vstack([np.zeros(g), np.eye(g)])
        nodes = [verts[i] for i in range(g + 1)] + [0.5 * (verts[i] + verts[j]) for i, j in ft.EDGES[g]]
        phi, dphi = ft.tabulate_p2(g, np.array(nodes))
        assert np.allclose(phi, np.eye(len(nodes)), atol=1e-15)
        pts = np.random.default_rng(0).dirichlet(np.ones(g + 1), size=7)[:, 1:]
        phi, dphi = ft.tabulate_p2(g, pts)
        assert np.allclose(phi.sum(axis=1), 1.0) and np.allclose(dphi.sum(axis=1), 0.0, atol=1e-14)


def test_complex_step_matches_finite_difference():
    mesh, sp, x, x1 = _case(3, False, 0.5)
    cf = CellForms(3, False, 4)
    prm = oa.region_params(TUBE_PARAMS, TUBE_REGIONS[0])
    sel = np.nonzero(mesh.markers == 1)[0][:6]
    X, u, u1 = mesh.coords[mesh.cells[sel]], x[sp.cell_dofs[sel]], x1[sp.cell_dofs[sel]]
    J = cf.jacobian("fluid", X, u, u1, prm, "all")
    eps = 1e-6
    for j in (3, 17, 35, 44, 61):
        up, um = u.copy(), u.copy()
        up[:, j] += eps
        um[:, j] -= eps
        fd = (sum(cf.residual("fluid", X, up, u1, prm)) - sum(cf.residual("fluid", X, um, u1, prm))) / (2 * eps)
        assert np.abs(fd - J[:, :, j]).max() < 1e-7 * np.abs(J).max()      # FD round-off scales with |R|


def test_piola_identity():
    """fluid.py:52 div(J F^-1 v): the literal expression (with second derivatives of d) equals
    tr(J F^-1 grad v) because div(adj F) = 0 for a gradient field — checked by finite differences
    of adj(F) on one P2 cell."""
    rng = np.random.default_rng(5)
    g = 3
    dco = 0.05 * rng.uniform(-1, 1, size=(g, ft.n_p2(g)))

    def adjF(pt):
        _, dphi = ft.tabulate_p2(g, pt[None, :])
        F = np.eye(g) + dco @ dphi[0]
        return np.linalg.det(F) * np.linalg.inv(F)
    pt = np.array([0.2, 0.3, 0.1])
    h = 1e-5
    div = np.zeros(g)
    for a in range(g):
        e = np.zeros(g)
        e[a] = h
        div += (adjF(pt + e)[a, :] - adjF(pt - e)[a, :]) / (2 * h)      # d_a A_ab
    assert np.abs(div).max() < 1e-8


@pytest.mark.parametrize("g,bih,deg,amp", [(3, False, 4, 0.5), (3, True, 4, 3.0), (2, False, 5, 0.2), (2, True, 4, 2.0)])
def test_kernel_math_matches_oracle(emul, g, bih, deg, amp):
    """host-compiled kernel mathematics (C' tensors, sum factorisation) vs complex-step oracle,
    including a large-deformation state (amp=3: |grad d| ~ 0.2)."""
    mesh, sp, x, x1 = _case(g, bih, amp)
    common = dict(TUBE_PARAMS, extrapolation="biharmonic" if bih else "laplace")
    cf = CellForms(g, bih, deg)
    pts, wts = map(np.ascontiguousarray, ft.quadrature(g, deg))
    X = np.ascontiguousarray(mesh.coords[mesh.cells])
    u, u1 = np.ascontiguousarray(x[sp.cell_dofs]), np.ascontiguousarray(x1[sp.cell_dofs])
    for reg in TUBE_REGIONS:
        kind = 0 if reg["kind"] == "fluid" else 1
        prm = oa.region_params(common, reg)
        sel = mesh.markers == reg["marker"]
        Xs, us, u1s = map(np.ascontiguousarray, (X[sel], u[sel], u1[sel]))
        Jn = cf.jacobian(reg["kind"], Xs, us, u1s, prm, "nl")
        Ae = np.zeros_like(Jn)
        rc = emul.emul_jacobian(g, len(wts), _dp(pts), _dp(wts), kind, C.c_double(reg["rho"]), C.c_double(reg["mu"]),
                                C.c_double(reg.get("lambda_", 0.0)), C.c_double(common["dt"]), C.c_double(common["theta"]),
                                C.c_int64(len(Xs)), _dp(Xs), _dp(us), _dp(u1s), sp.ndof_cell, _dp(Ae))
        assert rc == 0
        scale = np.abs(Jn).max(axis=2, keepdims=True)
        scale[scale == 0] = 1
        assert (np.abs(Ae - Jn) / scale).max() < 1e-12
        Rl, Rn = cf.residual(reg["kind"], Xs, us, u1s, prm)
        Fe = np.zeros_like(Rl)
        emul.emul_residual(g, len(wts), _dp(pts), _dp(wts), kind, C.c_double(reg["rho"]), C.c_double(reg["mu"]),
                           C.c_double(reg.get("lambda_", 0.0)), C.c_double(common["dt"]), C.c_double(common["theta"]),
                           1, 1, 2 if bih else 1, 0, C.c_double(0.0), C.c_double(0.0), C.c_int64(len(Xs)),
                           _dp(Xs), _dp(us), _dp(u1s), sp.ndof_cell, _dp(Fe))
        assert np.abs(Fe - (Rl + Rn)).max() < 1e-13 * np.abs(Rl + Rn).max()
        if not bih and kind == 0:
            # every laplace sub-type of the residual (laplace.py:33-40), incl. "volume_change" alfa = 1/J(d)
            for code, sub in ((1, "small_constant"), (2, "volume"), (3, "volume_change")):
                prm2 = oa.region_params(dict(common, extrapolation_sub_type=sub, hmin=0.37), reg)
                Rl2, Rn2 = cf.residual(reg["kind"], Xs, us, u1s, prm2)
                Fe2 = np.zeros_like(Rl2)
                emul.emul_residual(g, len(wts), _dp(pts), _dp(wts), kind, C.c_double(reg["rho"]), C.c_double(reg["mu"]),
                                   C.c_double(reg.get("lambda_", 0.0)), C.c_double(common["dt"]), C.c_double(common["theta"]),
                                   1, 1, 1, code, C.c_double(0.37), C.c_double(0.0), C.c_int64(len(Xs)),
                                   _dp(Xs), _dp(us), _dp(u1s), sp.ndof_cell, _dp(Fe2))
                assert np.abs(Fe2 - (Rl2 + Rn2)).max() < 1e-13 * np.abs(Rl2 + Rn2).max(), sub
                nd = g * (6 if g == 2 else 10)                          # the d rows carry the lifting term
                assert np.abs(Fe2[:, :nd] - Fe[:, :nd]).max() > 1e-3 * np.abs(Fe[:, :nd]).max()


@pytest.mark.parametrize("g,deg,amp", [(3, 4, 0.5), (3, 4, 3.0), (2, 5, 2.0)])
def test_nonlinear_lifting_math_matches_oracle(emul, g, deg, amp):
    """elastic lifting (elastic.py:18-27) and laplace "volume_change" (laplace.py:33-34): residual fluxes and the
    phi x delta d block of the FROZEN A_pre (newtonsolver.py:26-27) at a NON-ZERO set-up state, host-compiled kernel
    mathematics vs the complex-step oracle."""
    mesh, sp, x, x1 = _case(g, False, amp)
    pts, wts = map(np.ascontiguousarray, ft.quadrature(g, deg))
    cf = CellForms(g, False, deg)
    X = np.ascontiguousarray(mesh.coords[mesh.cells])
    u, u1 = np.ascontiguousarray(x[sp.cell_dofs]), np.ascontiguousarray(x1[sp.cell_dofs])
    reg = TUBE_REGIONS[0]
    sel = mesh.markers == reg["marker"]
    Xs, us, u1s = map(np.ascontiguousarray, (X[sel], u[sel], u1[sel]))
    nd = g * (6 if g == 2 else 10)
    for ext, sub, ext_code, sub_code, gen_kind in (("elastic", "constant", 3, 0, 2), ("laplace", "volume_change", 1, 3, 1)):
        common = dict(TUBE_PARAMS, extrapolation=ext, extrapolation_sub_type=sub)
        prm = oa.region_params(common, reg)
        Rl, Rn = cf.residual("fluid", Xs, us, u1s, prm)
        Fe = np.zeros_like(Rl)
        emul.emul_residual(g, len(wts), _dp(pts), _dp(wts), 0, C.c_double(reg["rho"]), C.c_double(reg["mu"]),
                           C.c_double(0.0), C.c_double(common["dt"]), C.c_double(common["theta"]),
                           1, 1, ext_code, sub_code, C.c_double(0.0), C.c_double(0.0), C.c_int64(len(Xs)),
                           _dp(Xs), _dp(us), _dp(u1s), sp.ndof_cell, _dp(Fe))
        assert np.abs(Fe - (Rl + Rn)).max() < 1e-13 * np.abs(Rl + Rn).max(), ext
        Jl = cf.jacobian("fluid", Xs, us, u1s, prm, "lin")[:, :nd, :nd]
        Ae = np.zeros((len(Xs), sp.ndof_cell, sp.ndof_cell))
        rc = emul.emul_lifting_jacobian(g, len(wts), _dp(pts), _dp(wts), gen_kind, C.c_int64(len(Xs)), _dp(Xs), _dp(us),
                                        sp.ndof_cell, _dp(Ae))
        assert rc == 0
        scale = np.abs(Jl).max(axis=2, keepdims=True)
        assert (np.abs(Ae[:, :nd, :nd] - Jl) / scale).max() < 1e-11, ext
        # the state matters: the block at d = 0 differs
        Ae0 = np.zeros_like(Ae)
        emul.emul_lifting_jacobian(g, len(wts), _dp(pts), _dp(wts), gen_kind, C.c_int64(len(Xs)), _dp(Xs), _dp(0 * us),
                                   sp.ndof_cell, _dp(Ae0))
        assert np.abs(Ae0 - Ae).max() > 1e-4 * np.abs(Ae).max()


MATERIALS = [dict(material_model="StVenantKirchoffEnergy"), dict(material_model="NeoHookean"),
             dict(material_model="Gent", Jm=2.5), dict(material_model="MooneyRivlin", C01=3.0e4, C10=1.2e5, C11=7.0e3)]
MAT_CODE = {"StVenantKirchoff": 0, "StVenantKirchoffEnergy": 1, "NeoHookean": 2, "MooneyRivlin": 3, "Gent": 4}


@pytest.mark.parametrize("g", [2, 3])
@pytest.mark.parametrize("mat", MATERIALS, ids=lambda m: m["material_model"])
def test_piola_is_the_derivative_of_the_energy(g, mat):
    """common.py:110-125: P = diff(W, F). The oracle's hand-written P against a complex-step derivative of the
    literally transcribed energy densities (common.py:214-273), at random deformation gradients."""
    from oracle.forms import piola1, strain_energy
    rng = np.random.default_rng(5)
    prm = dict(mat, lambda_s=3.1e6, mu_s=3.4e5)
    F = np.eye(g) + 0.15 * rng.uniform(-1, 1, (40, g, g))
    P = piola1(F, mat["material_model"], prm)
    h = 1e-30
    for a in range(g):
        for b in range(g):
            Fc = F.astype(np.complex128)
            Fc[:, a, b] += 1j * h
            dW = strain_energy(Fc, mat["material_model"], prm).imag / h
            assert np.abs(dW - P[:, a, b]).max() < 1e-12 * np.abs(P).max()
    if mat["material_model"] == "StVenantKirchoffEnergy":      # the two St.Venant-Kirchhoff routes of the reference agree
        assert np.abs(P - piola1(F, "StVenantKirchoff", prm)).max() < 1e-12 * np.abs(P).max()


def _general_solid(emul, g, deg, reg, common, Xs, us, u1s, ndof, part):
    ptsJ, wtsJ = map(np.ascontiguousarray, ft.quadrature(g, deg))
    ptsR, wtsR = map(np.ascontiguousarray, ft.quadrature(g, deg))
    out = np.zeros((len(Xs), ndof) if part == 2 else (len(Xs), ndof, ndof))
    d = C.c_double
    rc = emul.emul_solid_general(g, len(wtsJ), _dp(ptsJ), _dp(wtsJ), len(wtsR), _dp(ptsR), _dp(wtsR), d(reg["rho"]), d(reg["mu"]),
                                 d(reg["lambda_"]), d(common["dt"]), d(common["theta"]), MAT_CODE[reg.get("material_model", "StVenantKirchoff")],
                                 d(reg.get("C01", 0.0)), d(reg.get("C10", 0.0)), d(reg.get("C11", 0.0)), d(reg.get("Jm", 0.0)),
                                 {None: 0, "Form1": 1, "Form2": 2}[reg.get("viscoelasticity")], d(reg.get("mu_visc_s", 0.0)),
                                 d(reg.get("lambda_visc_s", 0.0)), part, C.c_int64(len(Xs)), _dp(Xs), _dp(us), _dp(u1s), ndof, _dp(out))
    assert rc == 0
    return out


@pytest.mark.parametrize("g,deg,amp", [(3, 4, 3.0), (2, 5, 2.0)])
@pytest.mark.parametrize("extra", MATERIALS + [dict(viscoelasticity="Form1", mu_visc_s=1.0e4, lambda_visc_s=2.0e4),
                                               dict(viscoelasticity="Form2", mu_visc_s=1.0e2, lambda_visc_s=3.0e2),
                                               dict(material_model="NeoHookean", viscoelasticity="Form1", mu_visc_s=1.0e4, lambda_visc_s=2.0e4)],
                         ids=lambda m: m.get("material_model", "SVK") + "-" + str(m.get("viscoelasticity")))
def test_material_and_viscoelastic_math_matches_oracle(emul, g, deg, amp, extra):
    """energy-based material models (common.py:110-125: tangent from hyper-dual derivatives of W(I1, I2, J)) and
    viscoelasticity (solid.py:46-59): host-compiled kernel mathematics — residual fluxes, J_nonlinear (psi x delta d,
    psi x delta v) and the "Form2" block of J_linear — against the complex-step oracle at a large-deformation state."""
    mesh, sp, x, x1 = _case(g, False, amp)
    common = dict(TUBE_PARAMS)
    reg = dict(TUBE_REGIONS[1], **extra)
    cf = CellForms(g, False, deg)
    X = np.ascontiguousarray(mesh.coords[mesh.cells])
    u, u1 = np.ascontiguousarray(x[sp.cell_dofs]), np.ascontiguousarray(x1[sp.cell_dofs])
    sel = mesh.markers == reg["marker"]
    Xs, us, u1s = map(np.ascontiguousarray, (X[sel], u[sel], u1[sel]))
    prm = oa.region_params(common, reg)
    Rl, Rn = cf.residual("solid", Xs, us, u1s, prm)
    Fe = _general_solid(emul, g, deg, reg, common, Xs, us, u1s, sp.ndof_cell, 2)
    assert np.abs(Fe - (Rl + Rn)).max() < 1e-12 * np.abs(Rl + Rn).max()
    for part, name in ((0, "nl"), (1, "lin")):
        Jo = cf.jacobian("solid", Xs, us, u1s, prm, name)
        Ae = _general_solid(emul, g, deg, reg, common, Xs, us, u1s, sp.ndof_cell, part)
        if part == 1:                      # the emulation returns only the state-dependent viscoelastic block of J_linear
            base = cf.jacobian("solid", Xs, us, u1s, oa.region_params(common, TUBE_REGIONS[1]), "lin")
            Jo = Jo - base
        scale = np.abs(Jo).max(axis=(1, 2), keepdims=True)
        scale[scale == 0] = 1
        assert (np.abs(Ae - Jo) / scale).max() < 1e-11, name
    if "viscoelasticity" in extra:         # the rate term matters
        Rl0, Rn0 = cf.residual("solid", Xs, us, u1s, oa.region_params(common, dict(reg, viscoelasticity=None)))
        assert np.abs(Rn - Rn0).max() > 1e-6 * np.abs(Rn).max()


@pytest.mark.parametrize("g,n", [(3, 3), (2, 6)])
def test_c_port_matches_numpy_oracle(g, n):
    mesh = tube_mesh(n) if g == 3 else square_mesh(n)
    sp = MixedSpace(mesh)
    pat = sp.pattern()
    x = synthetic_state(mesh, sp, 2024, h=1.0 / n)
    x1 = synthetic_state(mesh, sp, 2025, h=1.0 / n)
    degJ, degR = (4, 6) if g == 3 else (5, 10)
    ref = oa.assemble_matrix(mesh, sp, TUBE_REGIONS, TUBE_PARAMS, x, x1, "nl", degJ, pat)
    A = np.zeros(len(pat[1]))
    cport.assemble_jacobian(mesh, sp, TUBE_REGIONS, 1e-3, 0.501, x, x1, ft.quadrature(g, degJ), pat, A)
    rp = pat[0]
    scale = np.maximum.reduceat(np.abs(ref), rp[:-1])
    scale = np.repeat(np.where(scale > 0, scale, 1), np.diff(rp))
    assert (np.abs(A - ref) / scale).max() < 1e-12
    refb = oa.assemble_vector(mesh, sp, TUBE_REGIONS, TUBE_PARAMS, x, x1, degR)
    b = np.zeros(sp.num_dofs)
    cport.assemble_residual(mesh, sp, TUBE_REGIONS, 1e-3, 0.501, x, x1, ft.quadrature(g, degR), b)
    assert np.abs(b - refb).max() < 1e-13 * np.abs(refb).max()


def test_jacobian_block_structure():
    """structurally non-zero blocks of J_nonlinear / J_linear."""
    mesh, sp, x, x1 = _case(3, False, 0.5)
    cf = CellForms(3, False, 4)
    names = ["d"] * 30 + ["v"] * 30 + ["p"] * 4
    X, u, u1 = mesh.coords[mesh.cells], x[sp.cell_dofs], x1[sp.cell_dofs]

    def blocks(kind, marker, part):
        sel = np.nonzero(mesh.markers == marker)[0][:4]
        prm = oa.region_params(TUBE_PARAMS, TUBE_REGIONS[marker - 1])
        nz = np.abs(cf.jacobian(kind, X[sel], u[sel], u1[sel], prm, part)).max(axis=0) > 0
        return {(names[i], names[j]) for i in range(64) for j in range(64) if nz[i, j]}
    assert blocks("fluid", 1, "nl") == {("v", "d"), ("v", "v"), ("v", "p"), ("p", "d"), ("p", "v")}
    assert blocks("fluid", 1, "lin") == {("d", "d"), ("v", "v")}
    assert blocks("solid", 2, "nl") == {("v", "d")}
    assert blocks("solid", 2, "lin") == {("d", "d"), ("d", "v"), ("v", "v")}
