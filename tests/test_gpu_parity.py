"""GPU parity: the CUDA path (through the C-ABI) against the numpy oracle on seeded inputs.
Tolerance: 1e-12 relative to the row inf-norm (matrix) / to ||b||_inf (vector) — BASELINE.json
north_star "assembled matrix and residual entries within 1e-12 relative in FP64"."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _row_rel_err(vals, ref, rowptr):
    scale = np.maximum.reduceat(np.abs(ref), rowptr[:-1])
    # rows holding only round-off (e.g. facet integrals of basis functions that vanish on the facet) are
    # measured against the global scale instead of their own ~1e-19 entries
    scale = np.maximum(scale, 1e-10 * np.abs(ref).max())
    scale = np.repeat(np.where(scale > 0, scale, 1.0), np.diff(rowptr))
    return float((np.abs(vals - ref) / scale).max())


def _setup(gdim, biharmonic, n, numbering="interleaved"):
    import torch
    from turtlefsi_b200.mesh import tube_mesh, synthetic_state
    from turtlefsi_b200.space import MixedSpace
    from tests.helpers import square_mesh
    mesh = tube_mesh(n) if gdim == 3 else square_mesh(n)
    space = MixedSpace(mesh, biharmonic=biharmonic, numbering=numbering)
    h = 2.0 / n if gdim == 3 else 1.0 / n
    x = synthetic_state(mesh, space, 2024, h=h)
    x1 = synthetic_state(mesh, space, 2025, h=h)
    return mesh, space, x, x1


@pytest.mark.parametrize("gdim,biharmonic,n,numbering", [
    (3, False, 4, "interleaved"), (3, True, 3, "interleaved"), (3, False, 3, "blocked"),
    (2, False, 8, "interleaved"), (2, True, 8, "interleaved"),
])
def test_assembly_matches_oracle(gdim, biharmonic, n, numbering, tube_params):
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    from oracle import assembly as oa
    common, regions = tube_params
    if biharmonic:
        common["extrapolation"] = "biharmonic"
    mesh, space, x, x1 = _setup(gdim, biharmonic, n, numbering)
    degJ, degR = (4, 6) if gdim == 3 else (5, 10)
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"],
                       extrapolation=common["extrapolation"],
                       quad_degree_jacobian=degJ, quad_degree_residual=degR, pattern=pattern)
    xd = torch.as_tensor(x).cuda()
    x1d = torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    A = asm.assemble_jacobian(xd, x1d, A_pre)
    Jnl = asm.assemble_jacobian(xd, x1d, None)
    b = asm.assemble_residual(xd, x1d)
    torch.cuda.synchronize()
    ref_pre = oa.assemble_matrix(mesh, space, regions, common, x, x1, "lin", degJ, pattern)
    ref_nl = oa.assemble_matrix(mesh, space, regions, common, x, x1, "nl", degJ, pattern)
    ref_b = oa.assemble_vector(mesh, space, regions, common, x, x1, degR)
    rp = pattern[0]
    assert _row_rel_err(A_pre.cpu().numpy(), ref_pre, rp) < TOL
    assert _row_rel_err(Jnl.cpu().numpy(), ref_nl, rp) < TOL
    assert _row_rel_err(A.cpu().numpy(), ref_pre + ref_nl, rp) < TOL
    assert np.abs(b.cpu().numpy() - ref_b).max() / np.abs(ref_b).max() < TOL
    # compressed A_pre path (zero-fill + sparse overwrite) gives the same matrix
    nz = asm.compress_A_pre(A_pre)
    assert nz == int((ref_pre != 0).sum()) or nz == int((A_pre != 0).sum().item())
    A2 = asm.assemble_jacobian(xd, x1d, A_pre)
    torch.cuda.synchronize()
    assert _row_rel_err(A2.cpu().numpy(), ref_pre + ref_nl, rp) < TOL
    # CSR pattern / dof map are inputs and must come back untouched (bit-exact)
    assert np.array_equal(asm.rowptr.cpu().numpy(), pattern[0])
    assert np.array_equal(asm.colidx.cpu().numpy(), pattern[1])
    assert asm.launch_count() > 0


def test_ident_zeros_and_bcs(tube_params):
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    from oracle import assembly as oa
    common, regions = tube_params
    mesh, space, x, x1 = _setup(3, False, 3)
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], pattern=pattern)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A = asm.assemble_jacobian(xd, x1d, asm.assemble_A_pre(xd, x1d))
    ref = A.cpu().numpy().copy()
    asm.ident_zeros(A)
    oa.ident_zeros(ref, pattern)
    assert np.array_equal(A.cpu().numpy(), ref)
    # pressure rows strictly inside the solid are empty -> unit diagonal (newtonsolver.py:80-81)
    assert (ref == 1.0).sum() > 0
    rng = np.random.default_rng(7)
    bc = rng.choice(space.num_dofs, size=200, replace=False).astype(np.int32)
    g = rng.uniform(-1, 1, size=200)
    asm.set_bc_dofs(bc)
    gd = torch.as_tensor(g).cuda()
    asm.apply_bc_matrix(A)
    oa.apply_bc_matrix(ref, pattern, bc)
    assert np.array_equal(A.cpu().numpy(), ref)
    b = asm.assemble_residual(xd, x1d)
    bref = b.cpu().numpy().copy()
    asm.apply_bc_rhs(b, xd, gd)
    oa.apply_bc_rhs(bref, x, bc, g)
    assert np.array_equal(b.cpu().numpy(), bref)
    nrm = asm.norm(b)
    assert abs(nrm - np.linalg.norm(bref)) <= 1e-13 * np.linalg.norm(bref)
    asm.apply_bc_vec(xd, gd)
    assert np.array_equal(xd.cpu().numpy()[bc], g)


def test_device_pattern_builder_matches_host():
    import torch
    from turtlefsi_b200.mesh import tube_mesh
    from turtlefsi_b200.space import MixedSpace, pattern_device
    from tests.helpers import square_mesh
    for mesh, bih in ((tube_mesh(5), False), (tube_mesh(3), True), (square_mesh(7), False)):
        sp = MixedSpace(mesh, biharmonic=bih)
        rp, ci = sp.pattern()
        rpd, cid = pattern_device(sp, torch.device("cuda"))
        assert np.array_equal(rpd.cpu().numpy(), rp)
        assert np.array_equal(cid.cpu().numpy(), ci)


def test_host_entry_matches_device_calls(tube_params):
    """tfsi_newton_assemble_host (host buffers, the e2e call) == the sequence of device calls."""
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(3, False, 3)
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"])
    rng = np.random.default_rng(3)
    bc = rng.choice(space.num_dofs, size=100, replace=False).astype(np.int32)
    g = rng.uniform(-1, 1, size=100)
    asm.set_bc_dofs(bc)
    xd, x1d, gd = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda(), torch.as_tensor(g).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d)
    A = asm.assemble_jacobian(xd, x1d, A_pre)
    asm.ident_zeros(A)
    asm.apply_bc_matrix(A)
    b = asm.assemble_residual(xd, x1d)
    asm.apply_bc_rhs(b, xd, gd)
    A2 = asm.new_matrix()
    bh = np.zeros(space.num_dofs)
    ss = asm.newton_assemble_host(x, x1, g, A_pre, A2, bh, recompute=True)
    scale = np.abs(A.cpu().numpy()).max()
    assert np.abs(A2.cpu().numpy() - A.cpu().numpy()).max() < 1e-12 * scale     # atomic order only
    assert np.abs(bh - b.cpu().numpy()).max() < 1e-12 * np.abs(bh).max()
    assert abs(ss - float((bh ** 2).sum())) < 1e-10 * ss
    # the same through the bitmap-compressed A_pre stream, and with the Jacobian re-used (A untouched)
    asm.compress_A_pre(A_pre)
    A3 = asm.new_matrix()
    bh3 = np.zeros(space.num_dofs)
    asm.newton_assemble_host(x, x1, g, A_pre, A3, bh3, recompute=True)
    assert np.abs(A3.cpu().numpy() - A.cpu().numpy()).max() < 1e-12 * scale
    A3.fill_(7.0)
    bh4 = np.zeros(space.num_dofs)
    ss4 = asm.newton_assemble_host(x, x1, g, A_pre, A3, bh4, recompute=False)
    assert float(A3.min()) == 7.0 and float(A3.max()) == 7.0
    assert np.abs(bh4 - bh).max() < 1e-12 * np.abs(bh).max() and abs(ss4 / ss - 1) < 1e-12


def test_partitioned_assembly_owned_rows_equal_global(tube_params):
    """multi-GPU path emulated on one GPU: each rank's local assembly (own cells + ghost layer) gives,
    on the rows it owns, exactly the rows of the single-domain assembly (no reduction needed)."""
    import scipy.sparse as sp
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.partition import LocalSpaceGlobalKeys, SlabMesh, synthetic_state_global
    from turtlefsi_b200.space import MixedSpace
    common, regions = tube_params
    n, world = 3, 2

    def local(r, w, nz=None):
        slab = SlabMesh(n, r, w, nz=nz)
        space = MixedSpace(slab.mesh)
        gk = LocalSpaceGlobalKeys(slab, space)
        asm = FSIAssembler(slab.mesh, space, regions, common["dt"], common["theta"])
        x = synthetic_state_global(slab, space, gk, 2024, 0.3)
        x1 = synthetic_state_global(slab, space, gk, 2025, 0.3)
        xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
        A = asm.assemble_jacobian(xd, x1d, asm.assemble_A_pre(xd, x1d))
        b = asm.assemble_residual(xd, x1d)
        torch.cuda.synchronize()
        M = sp.csr_matrix((A.cpu().numpy(), asm.colidx.cpu().numpy(), asm.rowptr.cpu().numpy()),
                          shape=(space.num_dofs, space.num_dofs))
        # (global node key, component) of every local dof
        node = np.repeat(np.arange(space.num_nodes), space.node_ncomp)
        comp = np.arange(space.num_dofs) - space.node_base[node]
        return slab, space, gk, M, b.cpu().numpy(), gk.key[node], comp

    gslab, gspace, ggk, GM, gb, gkey, gcomp = local(0, 1, nz=n * world)
    lookup = {(int(k), int(c)): i for i, (k, c) in enumerate(zip(gkey, gcomp))}
    covered = 0
    for r in range(world):
        slab, space, gk, M, b, key, comp = local(r, world)
        to_global = np.array([lookup[(int(k), int(c))] for k, c in zip(key, comp)])
        owned = np.repeat(gk.owned, space.node_ncomp)
        Mo = M[owned].tocoo()
        rows_g = to_global[np.nonzero(owned)[0][Mo.row]]
        cols_g = to_global[Mo.col]
        L = sp.csr_matrix((Mo.data, (rows_g, cols_g)), shape=GM.shape)
        sel = np.zeros(GM.shape[0], dtype=bool)
        sel[to_global[owned]] = True
        D = (L - sp.diags(sel.astype(float)) @ GM).tocoo()
        scale = abs(GM).max()
        assert abs(D.data).max() < 1e-12 * scale
        assert np.abs(b[owned] - gb[to_global[owned]]).max() < 1e-12 * np.abs(gb).max()
        covered += int(sel.sum())
    assert covered == GM.shape[0]


def test_rcb_partitioned_assembly_owned_rows_equal_global(tube_params):
    """general-mesh partition (recursive coordinate bisection, partition_rcb.py), ranks emulated on one GPU:
    owned rows of each rank's local CUDA assembly == rows of the single-domain CUDA assembly, and the halo
    pack/unpack kernels move the ghost values."""
    import scipy.sparse as sp
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.partition_rcb import RankView, rcb_partition
    from turtlefsi_b200.space import MixedSpace
    common, regions = tube_params
    mesh = tube_mesh(4)
    world = 3
    gs = MixedSpace(mesh)
    part = rcb_partition(mesh, world)
    xg, x1g = synthetic_state(mesh, gs, 2024, h=0.5), synthetic_state(mesh, gs, 2025, h=0.5)

    def assemble(m, space, x, x1):
        asm = FSIAssembler(m, space, regions, common["dt"], common["theta"], hmin=mesh.hmin())
        xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
        A = asm.assemble_jacobian(xd, x1d, asm.assemble_A_pre(xd, x1d))
        b = asm.assemble_residual(xd, x1d)
        M = sp.csr_matrix((A.cpu().numpy(), asm.colidx.cpu().numpy(), asm.rowptr.cpu().numpy()), shape=(space.num_dofs,) * 2)
        return asm, M, b.cpu().numpy()
    _, GA, gb = assemble(mesh, gs, xg, x1g)
    scale, covered = abs(GA).max(), 0
    views = [RankView(mesh, r, world, cell_part=part, gspace=gs) for r in range(world)]
    for v in views:
        asm, A, b = assemble(v.mesh, v.space, v.scatter_global(xg), v.scatter_global(x1g))
        own = np.nonzero(v.dof_owned)[0]
        Ao = A[own].tocoo()
        lifted = sp.csr_matrix((Ao.data, (v.dof_global[own][Ao.row], v.dof_global[Ao.col])), shape=GA.shape)
        rows = np.zeros(gs.num_dofs)
        rows[v.dof_global[own]] = 1.0
        assert abs(lifted - sp.diags(rows) @ GA).max() < 1e-12 * scale
        assert np.abs(b[own] - gb[v.dof_global[own]]).max() < 1e-12 * np.abs(gb).max()
        covered += len(own)
        # halo kernels: pack what the peers need / unpack into the ghosts (the transport itself is NCCL)
        send, recv = v.halo_lists(views)
        xl = torch.as_tensor(v.scatter_global(xg)).cuda()
        for p_, idx in send.items():
            buf = torch.empty(len(idx), dtype=torch.float64, device="cuda")
            it = torch.as_tensor(idx).cuda()
            asm._check(asm.lib.tfsi_halo_pack(asm.h, xl.data_ptr(), it.data_ptr(), len(idx), buf.data_ptr(), None))
            assert np.array_equal(buf.cpu().numpy(), xg[v.dof_global[idx]])
        for p_, idx in recv.items():
            buf = torch.as_tensor(xg[v.dof_global[idx]] + 1.0).cuda()
            it = torch.as_tensor(idx).cuda()
            asm._check(asm.lib.tfsi_halo_unpack(asm.h, xl.data_ptr(), it.data_ptr(), len(idx), buf.data_ptr(), None))
            assert np.array_equal(xl.cpu().numpy()[idx], xg[v.dof_global[idx]] + 1.0)
    assert covered == gs.num_dofs


@pytest.mark.parametrize("gdim", [2, 3])
def test_facet_operator_in_A_pre_and_residual(gdim, tube_params):
    """biharmonic "constrained_disp_vel" exterior-facet terms (biharmonic.py:38-41) through the C-ABI
    (tfsi_set_facet_operator) against the oracle, on boundaries that carry NO Dirichlet condition."""
    import torch
    from oracle.backend import OracleAssembler
    from oracle.facets import biharmonic_facet_matrix
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.boundary import Facets, biharmonic_facet_operator
    common, regions = tube_params
    common["extrapolation"] = "biharmonic"
    mesh, space, x, x1 = _setup(gdim, True, 3 if gdim == 3 else 6)
    facets = Facets(mesh)
    facets.mark(lambda xx, on: on, 5)
    degJ, degR = (4, 6) if gdim == 3 else (5, 10)
    pattern = space.pattern()
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], extrapolation="biharmonic",
                       quad_degree_jacobian=degJ, quad_degree_residual=degR, pattern=pattern)
    asm.set_facet_operator(*biharmonic_facet_operator(mesh, space, facets, [5]))
    ora = OracleAssembler(mesh, space, regions, common, degJ, degR, pattern=pattern, use_c_port=True)
    sel = np.nonzero(facets.on_boundary)[0]
    ora.set_facet_operator(*biharmonic_facet_matrix(mesh, space, facets.vertices[sel], facets.cell0[sel], facets.local0[sel]))
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d).cpu().numpy()
    b = asm.assemble_residual(xd, x1d).cpu().numpy()
    ref_pre, ref_b = ora.assemble_A_pre(x, x1), ora.assemble_residual(x, x1)
    base = OracleAssembler(mesh, space, regions, common, degJ, degR, pattern=pattern, use_c_port=True)
    assert np.abs(ref_pre - base.assemble_A_pre(x, x1)).max() > 0           # the term is really there
    assert _row_rel_err(A_pre, ref_pre, pattern[0]) < TOL
    assert np.abs(b - ref_b).max() < TOL * np.abs(ref_b).max()


@pytest.mark.parametrize("gdim", [2, 3])
def test_robin_and_pressure_load_facet_terms(gdim, tube_params):
    """Robin ds_s terms (solid.py:83-84: implicit half -> A_pre and b, explicit half -> b only) and the
    interior-facet load p(t) n('+').psi('+') dS (problem_aneu.py:214) through the C-ABI
    (tfsi_set_facet_operator[_prev], tfsi_set_facet_load[_scale]) against the oracle's closed-form facet
    matrices."""
    import torch
    from oracle.backend import OracleAssembler
    from oracle.facets import normal_load, robin_matrices
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.boundary import Facets, interior_facet_normal_load, robin_facet_operators
    common, regions = tube_params
    mesh, space, x, x1 = _setup(gdim, False, 3 if gdim == 3 else 6)
    facets = Facets(mesh)
    solid_facet = facets.on_boundary & (mesh.markers[facets.cell0] == 2)
    facets.markers[solid_facet] = 33                                           # Robin wall: boundary facets of solid cells
    interface = ~facets.on_boundary & (mesh.markers[facets.cell0] != mesh.markers[np.maximum(facets.cell1, 0)])
    facets.markers[interface] = 22                                             # fsi_id
    assert solid_facet.sum() > 0 and interface.sum() > 0
    degJ, degR = (4, 6) if gdim == 3 else (5, 10)
    pattern = space.pattern()
    k_s, c_s, p_t = [1.0e5], [1.0e2], 1.3e4
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"],
                       quad_degree_jacobian=degJ, quad_degree_residual=degR, pattern=pattern)
    now, prev = robin_facet_operators(mesh, space, facets, [33], k_s, c_s, common["theta"])
    asm.set_facet_operator(*now)
    asm.set_facet_operator_prev(*prev)
    asm.set_facet_load(*interior_facet_normal_load(mesh, space, facets, 22))
    asm.set_facet_load_scale(p_t)
    # oracle: closed-form facet mass matrices; "+" = the solid cell (higher marker)
    ora = OracleAssembler(mesh, space, regions, common, degJ, degR, pattern=pattern, use_c_port=True)
    sel = np.nonzero(solid_facet)[0]
    onow, oprev = robin_matrices(mesh, space, facets.cell0[sel], facets.local0[sel], k_s[0], c_s[0], common["theta"])
    ora.facet_op, ora.facet_op_prev = onow, oprev
    fi = np.nonzero(interface)[0]
    plus_is_1 = mesh.markers[facets.cell1[fi]] > mesh.markers[facets.cell0[fi]]
    pc = np.where(plus_is_1, facets.cell1[fi], facets.cell0[fi])
    pl = np.where(plus_is_1, facets.local1[fi], facets.local0[fi])
    assert (mesh.markers[pc] == 2).all()
    ora.facet_load, ora.facet_load_scale = normal_load(mesh, space, pc, pl), p_t
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    A_pre = asm.assemble_A_pre(xd, x1d).cpu().numpy()
    b = asm.assemble_residual(xd, x1d).cpu().numpy()
    ref_pre, ref_b = ora.assemble_A_pre(x, x1), ora.assemble_residual(x, x1)
    base = OracleAssembler(mesh, space, regions, common, degJ, degR, pattern=pattern, use_c_port=True)
    base_b = base.assemble_residual(x, x1)
    assert np.abs(ref_pre - base.assemble_A_pre(x, x1)).max() > 0 and np.abs(ref_b - base_b).max() > 1e-6 * np.abs(base_b).max()
    assert _row_rel_err(A_pre, ref_pre, pattern[0]) < TOL
    assert np.abs(b - ref_b).max() < TOL * np.abs(ref_b).max()
    # the load follows the scalar the problem file updates per step, and n = 0 removes the terms again
    asm.set_facet_load_scale(0.0)
    ora.facet_load_scale = 0.0
    assert np.abs(asm.assemble_residual(xd, x1d).cpu().numpy() - ora.assemble_residual(x, x1)).max() < TOL * np.abs(ref_b).max()
    z = np.zeros(0)
    asm.set_facet_operator(z, z, z)
    asm.set_facet_operator_prev(z, z, z)
    asm.set_facet_load(z, z)
    assert np.abs(asm.assemble_residual(xd, x1d).cpu().numpy() - base_b).max() < TOL * np.abs(base_b).max()
    with pytest.raises(RuntimeError, match="out of range"):
        asm.set_facet_load(np.array([space.num_dofs]), np.array([1.0]))


def test_abi_error_behaviour(tube_params):
    """the boundary fails loudly: a pattern that is not DOLFIN's dense-cell-block union, unsupported
    forms, oversized quadrature rules and null pointers give error codes + messages, never a silent path"""
    import ctypes as C
    import torch
    from turtlefsi_b200 import lib as L
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(3, False, 3)
    rp, ci = space.pattern()
    # (a) drop one off-diagonal entry from the pattern
    k = int(rp[5]) + 1
    rp2 = rp.copy()
    rp2[6:] -= 1
    ci2 = np.delete(ci, k)
    with pytest.raises(RuntimeError, match="CSR pattern"):
        FSIAssembler(mesh, space, regions, common["dt"], common["theta"], pattern=(rp2, ci2))
    # (b) unknown lifting names are refused, with the reference's own message for the laplace sub-type (laplace.py:42)
    with pytest.raises(RuntimeError, match="not available on the B200 path"):
        FSIAssembler(mesh, space, regions, common["dt"], common["theta"], extrapolation="plastic", pattern=(rp, ci))
    with pytest.raises(RuntimeError, match="Could not find extrapolation method"):
        FSIAssembler(mesh, space, regions, common["dt"], common["theta"], extrapolation="laplace",
                     extrapolation_sub_type="volume_chnage", pattern=(rp, ci))
    # (c) a quadrature rule beyond the table sizes of the kernels: error at creation
    with pytest.raises(RuntimeError, match="quadrature rule too large"):
        FSIAssembler(mesh, space, regions, common["dt"], common["theta"], quad_degree_jacobian=8, pattern=(rp, ci))
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"], pattern=(rp, ci))
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    # (c2) BC value arrays must match the registered dof set; the compressed-A_pre sentinel needs a registered copy
    asm.set_bc_dofs(np.array([0, 1, 2], dtype=np.int32))
    with pytest.raises(RuntimeError, match="BC values"):
        asm.apply_bc_vec(xd.clone(), torch.zeros(2, dtype=torch.float64, device="cuda"))
    with pytest.raises(RuntimeError, match="no compressed A_pre"):
        asm.assemble_jacobian(xd, x1d, "compressed")
    with pytest.raises(RuntimeError, match="needs a compressed A_pre"):
        asm.bind_matrix(asm.new_matrix())
    # (d) null pointers
    lib = L.load()
    assert lib.tfsi_assemble_residual(asm.h, None, None, None, None) == -1
    assert b"null" in lib.tfsi_last_error(asm.h)
    assert lib.tfsi_create(None, None) == -1
    # (e) facet-operator entries outside the pattern
    far = np.array([0], dtype=np.int32), np.array([space.num_dofs - 1], dtype=np.int32), np.array([1.0])
    with pytest.raises(RuntimeError, match="outside the CSR pattern"):
        asm.set_facet_operator(*far)


RCB_NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from turtlefsi_b200.assembler import FSIAssembler
from turtlefsi_b200.mesh import tube_mesh, synthetic_state
from turtlefsi_b200.partition_rcb import HaloExchangeRCB, RankView
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
world = dist.get_world_size()
regions = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3),
           dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703)]
mesh = tube_mesh(6)
view = RankView(mesh, rank, world)
asm = FSIAssembler(view.mesh, view.space, regions, 1e-3, 0.501, hmin=mesh.hmin())
xg = synthetic_state(mesh, view.gspace, 2024, h=0.3)
truth = view.scatter_global(xg)
x = torch.as_tensor(truth.copy()).cuda()
x[torch.as_tensor(np.nonzero(~view.dof_owned)[0]).cuda()] = -1.0          # poison the ghosts
halo = HaloExchangeRCB(asm, *view.halo_lists(), rank)
halo.exchange(x)
torch.cuda.synchronize()
assert np.array_equal(x.cpu().numpy(), truth), "NCCL halo exchange did not restore the ghost values"
# residual norm over owned dofs + all-reduce == norm of the single-domain residual (computed on rank 0)
x1 = torch.as_tensor(view.scatter_global(synthetic_state(mesh, view.gspace, 2025, h=0.3))).cuda()
b = asm.assemble_residual(x, x1)
ss = (b[torch.as_tensor(view.dof_owned).cuda()] ** 2).sum().reshape(1)
dist.all_reduce(ss)
if rank == 0:
    from turtlefsi_b200.space import MixedSpace
    g = FSIAssembler(mesh, view.gspace, regions, 1e-3, 0.501)
    gb = g.assemble_residual(torch.as_tensor(xg).cuda(), torch.as_tensor(synthetic_state(mesh, view.gspace, 2025, h=0.3)).cuda())
    ref = float((gb ** 2).sum())
    assert abs(ss.item() / ref - 1) < 1e-12, (ss.item(), ref)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_rcb_two_gpu_nccl_halo_and_norm(tmp_path):
    """needs two GPUs (skipped on a single-GPU box): RankView + HaloExchangeRCB over NCCL, distributed norm."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "rcb_nccl_worker.py"
    script.write_text(RCB_NCCL_WORKER % root)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29619", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


@pytest.mark.parametrize("gdim", [2, 3])
def test_free_stream_is_preserved_on_a_moving_mesh_cuda(gdim):
    """size-independent property through the C-ABI: uniform velocity + constant pressure on a moving mesh leave the
    interior momentum / continuity rows of the residual at round-off (see tests/test_known_answers.py)."""
    import torch
    from tests.helpers import square_mesh
    from tests.test_known_answers import _free_stream_state, _interior_flow_rows
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = square_mesh(6) if gdim == 2 else box_mesh(5)
    mesh.markers[:] = 1
    space = MixedSpace(mesh)
    regions = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3)]
    x, x1 = _free_stream_state(mesh, space, 31)
    asm = FSIAssembler(mesh, space, regions, 1e-3, 0.501, quad_degree_residual=10 if gdim == 2 else 6)
    b = asm.assemble_residual(torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()).cpu().numpy()
    rows = _interior_flow_rows(mesh, space)
    assert len(rows) > 20 and np.abs(b).max() > 1e-3
    assert np.abs(b[rows]).max() < 1e-12 * np.abs(b).max()


@pytest.mark.parametrize("gdim", [2, 3])
def test_solid_residual_and_jacobian_are_objective_cuda(gdim):
    """rigid rotation of any size at rest: the solid residual vanishes (E = 0), and the non-linear Jacobian block
    psi x delta d annihilates infinitesimal rigid motions of the rotated body (delta d = W R x, W skew)."""
    import torch
    from tests.helpers import square_mesh
    from tests.test_known_answers import _rigid_rotation_state
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.boundary import interpolate
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = square_mesh(5) if gdim == 2 else box_mesh(4)
    mesh.markers[:] = 2
    space = MixedSpace(mesh)
    regions = [dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703)]
    x, x1 = _rigid_rotation_state(mesh, space)
    asm = FSIAssembler(mesh, space, regions, 1e-3, 0.501, quad_degree_residual=10 if gdim == 2 else 6)
    xd, x1d = torch.as_tensor(x).cuda(), torch.as_tensor(x1).cuda()
    b = asm.assemble_residual(xd, x1d).cpu().numpy()
    assert np.abs(b).max() < 1e-12 * 3103448.0
    # J_nl (psi x delta d) applied to an infinitesimal rotation of the CURRENT configuration: delta d = W (x + d)
    W = np.array([[0.0, -1.0], [1.0, 0.0]]) if gdim == 2 else np.array([[0.0, -0.3, 0.2], [0.3, 0.0, -0.5], [-0.2, 0.5, 0.0]])
    X = space.node_coords(mesh)
    cur = X + np.stack([x[space.node_dof(0, a)] for a in range(gdim)], axis=1)
    dd = np.zeros(space.num_dofs)
    interp = cur @ W.T
    for a in range(gdim):
        dd[space.node_dof(0, a)] = interp[:, a]
    A = asm.assemble_jacobian(xd, x1d, None)
    y = asm.matvec(A, torch.as_tensor(dd).cuda()).cpu().numpy()
    stretch = np.zeros(space.num_dofs)
    stretch[space.node_dof(0, 0)] = cur[:, 0]
    ys = asm.matvec(A, torch.as_tensor(stretch).cuda()).cpu().numpy()
    assert np.abs(ys).max() > 1.0 and np.abs(y).max() < 1e-11 * np.abs(ys).max()


def test_translation_of_the_displacement_changes_nothing_cuda(tube_params):
    """only grad d and d - d1 enter the forms: a rigid translation added to d at both time levels leaves the CUDA
    residual and Jacobian unchanged (up to the order of the atomic additions)."""
    import torch
    from turtlefsi_b200.assembler import FSIAssembler
    common, regions = tube_params
    mesh, space, x, x1 = _setup(3, False, 4)
    xs, x1s = x.copy(), x1.copy()
    for a, c in enumerate((0.4, -0.7, 1.3)):
        xs[space.node_dof(0, a)] += c
        x1s[space.node_dof(0, a)] += c
    asm = FSIAssembler(mesh, space, regions, common["dt"], common["theta"])
    t = lambda v: torch.as_tensor(v).cuda()
    b, bs = asm.assemble_residual(t(x), t(x1)).cpu().numpy(), asm.assemble_residual(t(xs), t(x1s)).cpu().numpy()
    assert np.abs(b - bs).max() < 1e-12 * np.abs(b).max()
    A, As = asm.assemble_jacobian(t(x), t(x1), None).cpu().numpy(), asm.assemble_jacobian(t(xs), t(x1s), None).cpu().numpy()
    assert np.abs(A - As).max() < 1e-12 * np.abs(A).max()
