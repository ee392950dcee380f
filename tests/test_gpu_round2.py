"""GPU: parity of the round-2 features through the C-ABI — persistent (bound) matrix with the compressed-A_pre
sentinel, quadrature rules integrated in chunks (3-D degree 5 / 6), colour-ordered deterministic scatter, several
fluid / solid regions (modules/domain.py:56-63,80-86), elastic lifting and general-state volume_change in the
frozen A_pre (elastic.py:18-27, laplace.py:33-34, newtonsolver.py:26-27), released column array."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _row_rel_err(a, ref, rowptr):
    scale = np.maximum.reduceat(np.abs(ref), rowptr[:-1])
    scale = np.repeat(np.where(scale > 0, scale, 1.0), np.diff(rowptr))
    return float((np.abs(a - ref) / scale).max())


def _setup(n=3, biharmonic=False):
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = tube_mesh(n)
    space = MixedSpace(mesh, biharmonic=biharmonic)
    x = synthetic_state(mesh, space, 2024, h=2.0 / n)
    x1 = synthetic_state(mesh, space, 2025, h=2.0 / n)
    return mesh, space, x, x1


def test_bound_matrix_partial_init_and_sentinel(tube_params):
    """A = Matrix(A_pre) (newtonsolver.py:28): after tfsi_bind_matrix an assembly rewrites only the rows J_nl touches.
    Across several Newton-like iterations (new state, ident_zeros, BC rows) the bound matrix must equal what the full
    path gives; the dense A_pre may be freed (TFSI_APRE_COMPRESSED)."""
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(4)
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"])
    rng = np.random.default_rng(5)
    bc = rng.choice(space.num_dofs, size=300, replace=False).astype(np.int32)
    asm.set_bc_dofs(bc)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    A_pre_keep = A_pre.clone()
    asm.compress_A_pre(A_pre)
    del A_pre                                   # the dense array is gone: only the sentinel works from here on
    A = asm.new_matrix()
    touched = asm.bind_matrix(A)
    assert 0 < touched < asm.nnz                # about half of the entries are (re)written per assembly
    assert torch.equal(A, A_pre_keep)           # bind = A := A_pre
    for it in range(3):
        xi = xd + 1e-3 * it * torch.sin(torch.arange(space.num_dofs, device="cuda", dtype=torch.float64))
        asm.assemble_jacobian(xi, x1d, "compressed", A)
        asm.ident_zeros(A)
        asm.apply_bc_matrix(A)
        ref = asm.assemble_jacobian(xi, x1d, A_pre_keep)          # unbound buffer, dense A_pre: full path
        asm.ident_zeros(ref)
        asm.apply_bc_matrix(ref)
        rp = asm.rowptr.cpu().numpy()
        assert _row_rel_err(A.cpu().numpy(), ref.cpu().numpy(), rp) < 1e-14, it
    # a changed Dirichlet set makes the next assembly a full one (unit rows of the old set must disappear)
    bc2 = rng.choice(space.num_dofs, size=100, replace=False).astype(np.int32)
    asm.set_bc_dofs(bc2)
    asm.assemble_jacobian(xd, x1d, "compressed", A)
    asm.ident_zeros(A)
    asm.apply_bc_matrix(A)
    ref = asm.assemble_jacobian(xd, x1d, A_pre_keep)
    asm.ident_zeros(ref)
    asm.apply_bc_matrix(ref)
    assert _row_rel_err(A.cpu().numpy(), ref.cpu().numpy(), asm.rowptr.cpu().numpy()) < 1e-14
    # host entry into the bound matrix, previous time level kept on the device (x1_host = None on the 2nd call)
    xh, x1h = torch.as_tensor(x).pin_memory(), torch.as_tensor(x1).pin_memory()
    gh = torch.zeros(len(bc2), dtype=torch.float64).pin_memory()
    bh = torch.empty(space.num_dofs, dtype=torch.float64).pin_memory()
    asm.newton_assemble_host(xh, x1h, gh, "compressed", A, bh, recompute=True)
    b1 = bh.clone()
    asm.newton_assemble_host(xh, None, gh, "compressed", A, bh, recompute=True)
    assert np.abs((bh - b1).numpy()).max() <= 1e-13 * np.abs(b1.numpy()).max()
    assert _row_rel_err(A.cpu().numpy(), ref.cpu().numpy(), asm.rowptr.cpu().numpy()) < 1e-14


def test_matrix_pair_fused_reinitialisation(tube_params):
    """tfsi_bind_matrix_pair: two persistent matrices used alternately, the idle one re-initialised by the fluid kernel
    of the running assembly (TMA zero-fill between its tasks; the A_pre non-zeros added back by REDs at the next assembly).
    Every assembly (also out-of-order use of the pair, and with ident_zeros / BC rows in between) must give the matrix
    of the plain path."""
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(6)
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"])
    rng = np.random.default_rng(9)
    asm.set_bc_dofs(rng.choice(space.num_dofs, size=500, replace=False).astype(np.int32))
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    asm.compress_A_pre(A_pre)
    A0, A1 = asm.new_matrix(), asm.new_matrix()
    asm.bind_matrix_pair(A0, A1, "compressed")
    rp = asm.rowptr.cpu().numpy()
    order = [A0, A1, A0, A1, A1, A0, A0, A1, A0]              # regular alternation, then repeats (foreground fallback)
    fused_before = asm.fused_launches()
    for it, A in enumerate(order):
        xi = xd + 1e-3 * (it + 1) * torch.cos(torch.arange(space.num_dofs, device="cuda", dtype=torch.float64))
        asm.assemble_jacobian(xi, x1d, "compressed", A)
        asm.ident_zeros(A)
        asm.apply_bc_matrix(A)
        ref = asm.assemble_jacobian(xi, x1d, A_pre)               # unbound buffer, full path
        asm.ident_zeros(ref)
        asm.apply_bc_matrix(ref)
        torch.cuda.synchronize()
        assert _row_rel_err(A.cpu().numpy(), ref.cpu().numpy(), rp) < 1e-14, it
    assert asm.fused_launches() - fused_before >= 5          # the alternating part did take the fused path
    # split call (halo overlap): phase 1 then phase 2
    A = A1 if order[-1] is A0 else A0
    asm.assemble_jacobian(None, None, "compressed", A, phase=1)
    asm.assemble_jacobian(xd, x1d, "compressed", A, phase=2)
    asm.ident_zeros(A)
    asm.apply_bc_matrix(A)                                    # bound matrices keep their unit rows: compare after both
    ref = asm.assemble_jacobian(xd, x1d, A_pre)
    asm.ident_zeros(ref)
    asm.apply_bc_matrix(ref)
    assert _row_rel_err(A.cpu().numpy(), ref.cpu().numpy(), rp) < 1e-14


@pytest.mark.parametrize("case", ["degree6_chunks", "materials", "form2_fallback", "facet_operator", "2d_biharmonic", "no_fluid_fallback"])
def test_matrix_pair_variants(case, tube_params):
    """the fused re-initialisation of the pair next to everything that changes what A_pre holds in the rewritten rows or how
    the cell kernels are launched: a quadrature rule run in chunks, several regions with energy-based / viscoelastic solids
    ("Form2" has a block in A_pre: foreground fallback), a facet operator on velocity rows (Robin term), the 2-D
    biharmonic space, a run without fluid cells (fallback). Alternating assemblies must equal the plain path."""
    import torch
    from tests.helpers import square_mesh
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    common, regions = tube_params
    regions = [dict(r) for r in regions]
    kw, gdim, bih, expect_fused = {}, 3, False, True
    if case == "degree6_chunks":
        kw = dict(quad_degree_jacobian=6)
    elif case == "materials":
        regions += [dict(regions[1], marker=3, material_model="NeoHookean", viscoelasticity="Form1", mu_visc_s=1.0e4, lambda_visc_s=2.0e4),
                    dict(regions[0], marker=1001, rho=1100.0, mu=7.0e-3)]
    elif case == "form2_fallback":
        regions[1].update(viscoelasticity="Form2", mu_visc_s=1.0e2, lambda_visc_s=3.0e2)
        expect_fused = False
    elif case == "2d_biharmonic":
        gdim, bih, kw = 2, True, dict(extrapolation="biharmonic", quad_degree_jacobian=5, quad_degree_residual=10)
    elif case == "no_fluid_fallback":
        kw, expect_fused = dict(fluid="no_fluid"), False
    mesh = tube_mesh(5) if gdim == 3 else square_mesh(8)
    if case == "materials":
        zc = mesh.coords[mesh.cells].mean(axis=1)[:, 2]
        mesh.markers = np.where(mesh.markers == 1, np.where(zc > 2.0, 1001, 1), np.where(zc > 2.0, 3, 2)).astype(np.int32)
    space = MixedSpace(mesh, biharmonic=bih)
    h = 2.0 / 5 if gdim == 3 else 0.5
    x, x1 = synthetic_state(mesh, space, 2024, h=h), synthetic_state(mesh, space, 2025, h=h)
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], **kw)
    rng = np.random.default_rng(3)
    asm.set_bc_dofs(rng.choice(space.num_dofs, size=200, replace=False).astype(np.int32))
    if case == "facet_operator":                     # entries on velocity rows (rewritten) and displacement rows (kept)
        rp, ci = asm.rowptr.cpu().numpy(), asm.colidx.cpu().numpy()
        rows = rng.choice(space.num_dofs, size=400, replace=False).astype(np.int32)
        cols = np.array([ci[rp[r] + rng.integers(0, rp[r + 1] - rp[r])] for r in rows], dtype=np.int32)
        asm.set_facet_operator(rows, cols, rng.uniform(1.0, 2.0, len(rows)) * 1e3)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    asm.compress_A_pre(A_pre)
    A0, A1 = asm.new_matrix(), asm.new_matrix()
    asm.bind_matrix_pair(A0, A1, "compressed")
    rp = asm.rowptr.cpu().numpy()
    f0 = asm.fused_launches()
    for it, A in enumerate([A0, A1, A0, A1, A0]):
        xi = xd + 1e-3 * (it + 1) * torch.cos(torch.arange(space.num_dofs, device="cuda", dtype=torch.float64))
        asm.assemble_jacobian(xi, x1d, "compressed", A)
        asm.ident_zeros(A)
        asm.apply_bc_matrix(A)
        ref = asm.assemble_jacobian(xi, x1d, A_pre)               # unbound buffer, full path
        asm.ident_zeros(ref)
        asm.apply_bc_matrix(ref)
        assert _row_rel_err(A.cpu().numpy(), ref.cpu().numpy(), rp) < 1e-13, (case, it)
    assert (asm.fused_launches() - f0 >= 3) == expect_fused, case


@pytest.mark.parametrize("degJ,npts", [(5, 15), (6, 24)])
def test_jacobian_rule_in_chunks(degJ, npts, tube_params):
    """compiler_parameters["quadrature_degree"] is user-settable (problems/__init__.py:16-17, argpar.py:224-229):
    3-D rules above 14 points are integrated in several launches whose REDs add up."""
    import torch
    from oracle import assembly as oa
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(3)
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], quad_degree_jacobian=degJ, pattern=pattern)
    assert len(asm.qJ[1]) == npts
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    l0 = asm.launch_count()
    Jnl = asm.assemble_jacobian(xd, x1d, None)
    assert asm.launch_count() - l0 >= 1 + 2 + 1          # init + two chunks over the fluid cells (their per-point table
                                                         # holds 14 points) + one launch over the solid cells
    A_pre = asm.assemble_A_pre(xd, x1d)
    ref_nl = oa.assemble_matrix(mesh, space, regions, common, x, x1, "nl", degJ, pattern)
    ref_pre = oa.assemble_matrix(mesh, space, regions, common, x, x1, "lin", degJ, pattern)
    assert _row_rel_err(Jnl.cpu().numpy(), ref_nl, pattern[0]) < TOL
    assert _row_rel_err(A_pre.cpu().numpy(), ref_pre, pattern[0]) < TOL


@pytest.mark.parametrize("gdim", [3, 2])
def test_deterministic_scatter_is_bitwise_reproducible(gdim, tube_params):
    """north_star design point 3: colour-ordered scatter. Two runs give bit-identical matrices and residuals, and the
    same values as the RED path / the oracle."""
    import torch
    from oracle import assembly as oa
    from tests.helpers import square_mesh
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    common, regions = tube_params
    mesh = tube_mesh(6) if gdim == 3 else square_mesh(24)
    space = MixedSpace(mesh)
    h = 2.0 / 6 if gdim == 3 else 1.0 / 24
    x, x1 = synthetic_state(mesh, space, 2024, h=h), synthetic_state(mesh, space, 2025, h=h)
    degJ, degR = (4, 6) if gdim == 3 else (5, 10)
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], quad_degree_jacobian=degJ,
                       quad_degree_residual=degR, deterministic=True)
    assert asm.num_colours >= (8 if gdim == 3 else 4)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    runs = []
    for _ in range(3):
        A = asm.assemble_jacobian(xd, x1d, A_pre)
        b = asm.assemble_residual(xd, x1d)
        runs.append((A.clone(), b.clone()))
    for A, b in runs[1:]:
        assert torch.equal(A, runs[0][0]) and torch.equal(b, runs[0][1])
    # a second handle (fresh colouring, fresh maps) reproduces the same bits
    asm2 = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], quad_degree_jacobian=degJ,
                        quad_degree_residual=degR, deterministic=True)
    A2 = asm2.assemble_jacobian(xd, x1d, asm2.assemble_A_pre(xd, x1d))
    assert torch.equal(A2, runs[0][0])
    if gdim == 3 and mesh.num_cells <= 2000:
        pattern = space.pattern()
        ref = oa.assemble_matrix(mesh, space, regions, common, x, x1, "all", degJ, pattern)
        assert _row_rel_err(runs[0][0].cpu().numpy(), ref, pattern[0]) < TOL
    red = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], quad_degree_jacobian=degJ,
                       quad_degree_residual=degR)
    A3 = red.assemble_jacobian(xd, x1d, red.assemble_A_pre(xd, x1d))
    b3 = red.assemble_residual(xd, x1d)
    rp = red.rowptr.cpu().numpy()
    assert _row_rel_err(A3.cpu().numpy(), runs[0][0].cpu().numpy(), rp) < 1e-14
    assert np.abs((b3 - runs[0][1]).cpu().numpy()).max() < 1e-14 * np.abs(b3.cpu().numpy()).max()


def test_several_fluid_and_solid_regions():
    """dx_f_id = [1, 1001], dx_s_id = [2, 3] with different constants per region
    (problems/Test_Cylinder/problem_aneu_2fluid.py:49-50, problem_aneu_2solid.py:54-55, modules/domain.py:56-63,80-86)."""
    import torch
    from oracle import assembly as oa
    from turtlefsi_b200.assembler import FSIAssembler
    mesh, space, x, x1 = _setup(4)
    zc = mesh.coords[mesh.cells].mean(axis=1)[:, 2]
    mesh.markers = np.where(mesh.markers == 1, np.where(zc > 2.0, 1001, 1), np.where(zc > 2.0, 3, 2)).astype(np.int32)
    regions = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3),
               dict(marker=1001, kind="fluid", rho=1100.0, mu=7.0e-3),
               dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703),
               dict(marker=3, kind="solid", rho=1200.0, mu=5.0e5, lambda_=2.0e6)]
    for m in (1, 1001, 2, 3):
        assert (mesh.markers == m).sum() > 20
    common = dict(dt=1e-3, theta=0.501, extrapolation="laplace", extrapolation_sub_type="constant")
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], pattern=pattern)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    A = asm.assemble_jacobian(xd, x1d, A_pre)
    b = asm.assemble_residual(xd, x1d)
    ref = oa.assemble_matrix(mesh, space, regions, common, x, x1, "all", 4, pattern)
    ref_b = oa.assemble_vector(mesh, space, regions, common, x, x1, 6)
    assert _row_rel_err(A.cpu().numpy(), ref, pattern[0]) < TOL
    assert np.abs(b.cpu().numpy() - ref_b).max() / np.abs(ref_b).max() < TOL
    # the constants matter: one region table for all cells gives another matrix
    same = [dict(regions[0]), dict(regions[0], marker=1001), dict(regions[2]), dict(regions[2], marker=3)]
    ref_same = oa.assemble_matrix(mesh, space, same, common, x, x1, "all", 4, pattern)
    assert _row_rel_err(ref_same, ref, pattern[0]) > 1e-3


@pytest.mark.parametrize("gdim,ext,sub", [(3, "elastic", "constant"), (3, "laplace", "volume_change"),
                                          (2, "elastic", "constant"), (2, "laplace", "volume_change")])
def test_nonlinear_lifting_frozen_at_nonzero_state(gdim, ext, sub, tube_params):
    """elastic lifting / laplace volume_change: the residual uses the current state, the Jacobian block is part of
    A_pre, frozen at the state handed to tfsi_assemble_A_pre — here a NON-ZERO one (an `initiate` hook that sets d)."""
    import torch
    from oracle import assembly as oa
    from tests.helpers import square_mesh
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    common, regions = tube_params
    common.update(extrapolation=ext, extrapolation_sub_type=sub)
    mesh = tube_mesh(3) if gdim == 3 else square_mesh(6)
    space = MixedSpace(mesh)
    h = 3.0 if gdim == 3 else 2.0                     # large deformation gradients: |grad d| ~ 0.2
    x, x1 = synthetic_state(mesh, space, 2024, h=h), synthetic_state(mesh, space, 2025, h=h)
    degJ, degR = (4, 6) if gdim == 3 else (5, 10)
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], extrapolation=ext,
                       extrapolation_sub_type=sub, quad_degree_jacobian=degJ, quad_degree_residual=degR, pattern=pattern)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    b = asm.assemble_residual(xd, x1d)
    ref_pre = oa.assemble_matrix(mesh, space, regions, common, x, x1, "lin", degJ, pattern)
    ref_b = oa.assemble_vector(mesh, space, regions, common, x, x1, degR)
    assert _row_rel_err(A_pre.cpu().numpy(), ref_pre, pattern[0]) < 1e-11
    assert np.abs(b.cpu().numpy() - ref_b).max() / np.abs(ref_b).max() < TOL


SOLID_MODELS = [dict(material_model="StVenantKirchoffEnergy"), dict(material_model="NeoHookean"),
                dict(material_model="Gent", Jm=2.5), dict(material_model="MooneyRivlin", C01=3.0e4, C10=1.2e5, C11=7.0e3),
                dict(viscoelasticity="Form1", mu_visc_s=1.0e4, lambda_visc_s=2.0e4),
                dict(viscoelasticity="Form2", mu_visc_s=1.0e2, lambda_visc_s=3.0e2),
                dict(material_model="NeoHookean", viscoelasticity="Form1", mu_visc_s=1.0e4, lambda_visc_s=2.0e4)]


@pytest.mark.parametrize("gdim", [3, 2])
@pytest.mark.parametrize("extra", SOLID_MODELS, ids=lambda m: m.get("material_model", "SVK") + "-" + str(m.get("viscoelasticity")))
def test_material_models_and_viscoelasticity(gdim, extra, tube_params):
    """solid_properties["material_model"] in {StVenantKirchoffEnergy, NeoHookean, MooneyRivlin, Gent} (common.py:110-125,
    214-273) and ["viscoelasticity"] in {Form1, Form2} (solid.py:46-59): A_pre, A = A_pre + J_nl and b of a fluid region
    + a DEFAULT solid region + a solid region with the model, at a large-deformation state, against the oracle."""
    import torch
    from oracle import assembly as oa
    from tests.helpers import square_mesh
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    common, regions = tube_params
    mesh = tube_mesh(3) if gdim == 3 else square_mesh(6)
    space = MixedSpace(mesh)
    ax = mesh.coords[mesh.cells].mean(axis=1)[:, gdim - 1]
    mesh.markers = np.where((mesh.markers == 2) & (ax > np.median(ax)), 3, mesh.markers).astype(np.int32)
    regions = list(regions) + [dict(regions[1], marker=3, **extra)]
    assert (mesh.markers == 3).sum() > 5 and (mesh.markers == 2).sum() > 5
    h = 3.0 if gdim == 3 else 2.0
    x, x1 = synthetic_state(mesh, space, 2024, h=h), synthetic_state(mesh, space, 2025, h=h)
    degJ, degR = (4, 6) if gdim == 3 else (5, 10)
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], quad_degree_jacobian=degJ,
                       quad_degree_residual=degR, pattern=pattern)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    A = asm.assemble_jacobian(xd, x1d, A_pre)
    b = asm.assemble_residual(xd, x1d)
    ref_pre = oa.assemble_matrix(mesh, space, regions, common, x, x1, "lin", degJ, pattern)
    ref = oa.assemble_matrix(mesh, space, regions, common, x, x1, "all", degJ, pattern)
    ref_b = oa.assemble_vector(mesh, space, regions, common, x, x1, degR)
    assert _row_rel_err(A_pre.cpu().numpy(), ref_pre, pattern[0]) < 1e-11
    assert _row_rel_err(A.cpu().numpy(), ref, pattern[0]) < 1e-11
    assert np.abs(b.cpu().numpy() - ref_b).max() / np.abs(ref_b).max() < TOL
    # the model matters: the default St.Venant-Kirchhoff table for region 3 gives another matrix
    if extra.get("material_model", "") != "StVenantKirchoffEnergy":
        plain = list(regions[:-1]) + [dict(regions[1], marker=3)]
        assert _row_rel_err(oa.assemble_matrix(mesh, space, plain, common, x, x1, "all", degJ, pattern), ref, pattern[0]) > 1e-4


def test_release_pattern(tube_params):
    """the column array (4 B per entry) is only needed during set-up: after tfsi_release_pattern the hot path still
    runs, the calls that need columns fail loudly."""
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(3)
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"])
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    A0 = asm.assemble_jacobian(xd, x1d, A_pre)
    b0 = asm.assemble_residual(xd, x1d)
    asm.compress_A_pre(A_pre)
    asm.release_pattern()
    A1 = asm.assemble_jacobian(xd, x1d, "compressed")
    b1 = asm.assemble_residual(xd, x1d)
    rp = asm.rowptr.cpu().numpy()
    assert _row_rel_err(A1.cpu().numpy(), A0.cpu().numpy(), rp) < 1e-14
    assert np.abs((b1 - b0).cpu().numpy()).max() < 1e-14 * np.abs(b0.cpu().numpy()).max()
    with pytest.raises(RuntimeError, match="released"):
        asm.assemble_A_pre(xd, x1d)
    with pytest.raises(RuntimeError, match="released"):
        asm.matvec(A1, xd)
