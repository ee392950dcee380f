"""DOLFIN-free KNOWN ANSWERS the reference itself ships (SURVEY 8c "Additional known answers"):
the Taylor-Green vortex (turtleFSI/problems/TaylorGreen2D.py: analytic u, p :134-155) driven through
the ORACLE and the reference's Newton loop. The discrete solution converges to the analytic one at
the Taylor-Hood rates, which pins the fluid forms N1-N5 (fluid.py:32-52), the theta scheme and the
time-dependent Dirichlet data independently of any FEniCS output."""
import os

import numpy as np
import pytest


def _run(N, **over):
    from tests.test_turek_pin import _run_oracle
    ov = dict(N=N, T=0.1, dt=0.01)
    ov.update(over)
    return _run_oracle("TaylorGreen2D", ov)


def test_rectangle_mesh_is_dolfins():
    from turtlefsi_b200.mesh import rectangle_mesh
    m = rectangle_mesh((-1, -1), (1, 1), 3, 2, "right")
    assert m.num_vertices == 12 and m.num_cells == 12
    assert np.allclose(m.coords[5], (-1 + 2 / 3, 0.0))                  # vertex iy*(nx+1)+ix
    assert m.cells[:2].tolist() == [[0, 1, 5], [0, 4, 5]]               # "right": (v0,v1,v3),(v0,v2,v3)
    left = rectangle_mesh((-1, -1), (1, 1), 3, 2, "left")
    assert left.cells[:2].tolist() == [[0, 1, 4], [1, 4, 5]]
    X = m.coords[m.cells]
    area = 0.5 * np.abs(np.linalg.det(np.swapaxes(X[:, 1:] - X[:, :1], 1, 2))).sum()
    assert abs(area - 4.0) < 1e-14


def test_taylor_green_converges_to_the_analytic_solution():
    a, b = _run(8), _run(16)
    ev = (a["error_v_list"][-1], b["error_v_list"][-1])
    ep = (a["error_p_list"][-1], b["error_p_list"][-1])
    # P2 velocity against its interpolant: better than third order; P1 pressure: second order
    assert ev[1] < 5e-3 and ev[0] / ev[1] > 8.0
    assert ep[1] < 7e-2 and ep[0] / ep[1] > 3.5
    # the pressure is pinned at the corner to the analytic value (pointwise BC)
    sp, mesh = b["DVP"], b["mesh"]
    corner = np.nonzero((np.abs(mesh.coords - 1.0) < 1e-14).all(axis=1))[0]
    x = b["dvp_"]["n"]
    assert np.allclose(x[sp.vertex_dof_p()[corner]], b["p_bc_val"](mesh.coords[corner]), rtol=0, atol=1e-14)
    # no mesh motion: the d rows are ident_zeros rows and d stays 0
    assert np.abs(x[sp.sub_dofs(0)]).max() == 0.0
    # Jacobian re-used over the whole run (recompute = recompute_tstep = 100): Newton still converges
    assert all(h[1] < 10 for h in b["residual_history"])
    assert b["total_error_v"] > 0 and abs(b["total_error_v"] - 0.01 * sum(b["error_v_list"])) < 1e-15


def _oracle_run(problem, **over):
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_oracle(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"],
                               ns.get("residual_quadrature_degree", 10), use_c_port=True)
    ov = dict(verbose=False)
    ov.update(over)
    return monolithic.run(problem, make_oracle, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs), overrides=ov)


STIFF = dict(mu_s=1e9 / (2 * 1.45), lambda_s=0.45 * 2 * (1e9 / (2 * 1.45)) / (1 - 2 * 0.45))


def test_robin_condition_reproduces_the_mass_spring_damper():
    """RobinBC_validation.py: a cylinder under gravity held by the Robin wall follows
    m x'' = F - c A x' - k A x (:96-108). With a stiff body (E = 1e9 instead of 1e6, so that the mid-point
    moves with the wall) the oracle's mid-point displacement matches the ODE to 6e-4 of its peak at
    dt = 5e-3 — inertia, gravity and both halves of the Robin term (solid.py:41,66-68,83-84)."""
    ns = _oracle_run("RobinBC_validation", dt=0.005, T=0.1, ncube=3, **STIFF)
    d = np.array(ns["d_list"])
    a = ns["default_variables"]["analytical_displacement"]
    assert len(d) == 21 and abs(a).max() > 5e-3
    assert np.abs(d - a).max() < 1e-3 * np.abs(a).max()
    p = ns["default_variables"]["ode_params"]
    assert abs(p["A"] / (2 * np.pi * 4.0) - 1) < 0.02 and abs(p["m"] / (1e3 * np.pi * 4.0) - 1) < 0.06   # polyhedral cylinder


def test_facet_builders_match_the_oracle_closed_forms():
    """host-side COO builders (quadrature) vs the oracle's closed-form P2 facet matrices, and the
    closed-surface identity  sum_i l_i = int n dS = 0  of the interior-facet load."""
    import scipy.sparse as sp
    from oracle.facets import normal_load, robin_matrices
    from tests.helpers import square_mesh
    from turtlefsi_b200.boundary import Facets, interior_facet_normal_load, robin_facet_operators
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    for mesh in (square_mesh(6), box_mesh(4)):
        g = mesh.gdim
        cen = mesh.coords[mesh.cells].mean(axis=1)
        mesh.markers = np.where(np.linalg.norm(cen - 0.5, axis=1) < 0.3, 2, 1).astype(np.int32)   # enclosed solid
        space = MixedSpace(mesh)
        N = space.num_dofs
        facets = Facets(mesh)
        facets.markers[facets.on_boundary] = 7
        interface = ~facets.on_boundary & (mesh.markers[facets.cell0] != mesh.markers[np.maximum(facets.cell1, 0)])
        facets.markers[interface] = 22
        (r, c, v), (rp, cp, vp) = robin_facet_operators(mesh, space, facets, [7], [3.0], [0.5], 0.6)
        sel = np.nonzero(facets.on_boundary)[0]
        onow, oprev = robin_matrices(mesh, space, facets.cell0[sel], facets.local0[sel], 3.0, 0.5, 0.6)
        now = sp.coo_matrix((v, (r, c)), shape=(N, N)).tocsr()
        prev = sp.coo_matrix((vp, (rp, cp)), shape=(N, N)).tocsr()
        assert abs(now - onow).max() < 1e-15 * abs(onow).max() * 10 and abs(prev - oprev).max() < 1e-14 * abs(oprev).max()
        # rows are psi (v-test) dofs only; k_s d + c_s v: the d- and v-column blocks are 3.0 : 0.5
        assert set(np.unique(r)) <= set(space.sub_dofs(1))
        ones_d, ones_v = np.zeros(N), np.zeros(N)
        ones_d[space.sub_dofs(0, 0)] = 1.0
        ones_v[space.sub_dofs(1, 0)] = 1.0
        length = facets.on_boundary.sum() and (now @ ones_d)[space.sub_dofs(1, 0)].sum() / (0.6 * 3.0)
        assert abs(length - (4.0 if g == 2 else 6.0)) < 1e-12                  # = measure of the boundary
        assert abs((now @ ones_v)[space.sub_dofs(1, 0)].sum() / (0.6 * 0.5) - length) < 1e-12
        # interior-facet load: "+" = the higher cell marker (the enclosed solid), n points out of it
        dofs, vals = interior_facet_normal_load(mesh, space, facets, 22)
        fi = np.nonzero(interface)[0]
        plus1 = mesh.markers[facets.cell1[fi]] > mesh.markers[facets.cell0[fi]]
        pc = np.where(plus1, facets.cell1[fi], facets.cell0[fi])
        pl = np.where(plus1, facets.local1[fi], facets.local0[fi])
        ref = normal_load(mesh, space, pc, pl)
        l = np.zeros(N)
        l[dofs] = vals
        assert np.abs(l - ref).max() < 1e-14 * np.abs(ref).max() and np.abs(ref).max() > 0
        for a in range(g):
            assert abs(l[space.sub_dofs(1, a)].sum()) < 1e-14                  # closed surface: int n dS = 0
        # divergence theorem with psi = x e_a: sum_i l_i x_i(a) = int n_a x_a dS = volume of the enclosed solid
        X = space.node_coords(mesh)
        Xc = mesh.coords[mesh.cells[mesh.markers == 2]]
        vol = np.abs(np.linalg.det(Xc[:, 1:] - Xc[:, :1])).sum() / (2 if g == 2 else 6)
        for a in range(g):
            assert abs(l[space.node_dof(1, a)] @ X[:, a] - vol) < 1e-13


def test_checkpoint_and_restart_continue_the_run(tmp_path):
    """problems/__init__.py:183-219, 361-381 + monolithic.py:44-50, 160-162: a run stopped after 4 steps and
    restarted from its checkpoint ends where the uninterrupted 7-step run ends (to the Newton tolerance: the
    restart re-assembles the Jacobian on its first step, newtonsolver.py:63)."""
    full = _run(8, T=0.06)
    folder = str(tmp_path / "tg")
    first = _run(8, T=0.03, results_folder=folder, checkpoint_step=1)
    assert len(first["error_v_list"]) == 4 and os.path.exists(os.path.join(folder, "Checkpoint", "checkpoint_fields.npy"))
    again = _run(8, T=0.06, restart_folder=folder)
    assert again["first_step_num"] == 3 and abs(again["t"] - full["t"]) < 1e-12 and len(again["error_v_list"]) == 3
    a, b = np.asarray(again["dvp_"]["n"]), np.asarray(full["dvp_"]["n"])
    assert np.abs(a - b).max() < 1e-8 * np.abs(b).max()
    assert abs(again["error_v_list"][-1] / full["error_v_list"][-1] - 1) < 1e-7
    # the first Newton iteration of the restarted run rebuilds the Jacobian (counter == first_step_num)
    assert again["residual_history"][0][:2] == (3, 0)
    bad = np.load(os.path.join(folder, "Checkpoint", "checkpoint_fields.npy"))
    bad[5] = np.nan
    np.save(os.path.join(folder, "Checkpoint", "checkpoint_fields.npy"), bad)
    import pytest
    with pytest.raises(ValueError, match="Nan values"):
        _run(8, T=0.06, restart_folder=folder)


def _free_stream_state(mesh, space, seed):
    """moving mesh (d != d1, both smooth, det F > 0), uniform velocity at both time levels, constant pressure"""
    from turtlefsi_b200.mesh import synthetic_state
    g = mesh.gdim
    x, x1 = synthetic_state(mesh, space, seed, h=0.3), synthetic_state(mesh, space, seed + 1, h=0.3)
    vel = np.array([0.3, -0.2, 0.5])[:g]
    for xx in (x, x1):
        for a in range(g):
            xx[space.node_dof(1, a)] = vel[a]
    x[space.vertex_dof_p()] = 7.0
    x1[space.vertex_dof_p()] = -3.0
    return x, x1


def _interior_flow_rows(mesh, space):
    from turtlefsi_b200.boundary import Facets
    f = Facets(mesh)
    bverts = np.unique(f.vertices[f.on_boundary])
    on_b = np.zeros(space.num_nodes, dtype=bool)
    on_b[space.vertex_node[bverts]] = True
    ev = space.edge_vertices
    isb = np.zeros(mesh.num_vertices, dtype=bool)
    isb[bverts] = True
    on_b[space.edge_node] = isb[ev[:, 0]] & isb[ev[:, 1]]        # conservative: edges between boundary vertices
    inner = ~on_b
    rows = [space.node_dof(1, a)[inner] for a in range(mesh.gdim)]
    vin = np.setdiff1d(np.arange(mesh.num_vertices), bverts)
    rows.append(np.asarray(space.vertex_dof_p())[vin])
    return np.concatenate(rows)


@pytest.mark.parametrize("gdim", [2, 3])
def test_free_stream_is_preserved_on_a_moving_mesh(gdim):
    """ALE consistency (a size-independent property of fluid.py:32-56): a uniform velocity with constant pressure
    solves the equations on ANY moving mesh — N1, N2, N6, L1, L2 vanish with grad v = 0 and v = v1, N5 is
    tr(adj(F) grad v) = 0, and -p adj(F)^T : grad psi integrates to zero for interior test functions (Piola
    identity). The momentum and continuity residual rows of interior nodes are therefore zero to round-off."""
    from oracle.backend import OracleAssembler
    from tests.conftest import TUBE_PARAMS
    from tests.helpers import square_mesh
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = square_mesh(6) if gdim == 2 else box_mesh(3)
    mesh.markers[:] = 1
    space = MixedSpace(mesh)
    regions = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3)]
    x, x1 = _free_stream_state(mesh, space, 31)
    rows = _interior_flow_rows(mesh, space)
    assert len(rows) > 20
    for use_c in (True, False):
        ora = OracleAssembler(mesh, space, regions, dict(TUBE_PARAMS), 4, 10 if gdim == 2 else 6, use_c_port=use_c)
        b = ora.assemble_residual(x, x1)
        scale = np.abs(b).max()
        assert scale > 0.1                                     # boundary rows carry the pressure load, d rows the lifting
        assert np.abs(b[rows]).max() < 1e-12 * scale


def _rigid_rotation_state(mesh, space):
    """d = (R - I) x + c for a LARGE rotation R, the same at both time levels, zero velocity"""
    from turtlefsi_b200.boundary import interpolate
    g = mesh.gdim
    if g == 2:
        t = 0.9
        R = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
    else:
        a, b = 0.7, -1.1
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        R = Rz @ Rx
    x = np.zeros(space.num_dofs)
    interpolate(space, mesh, x, 0, lambda X: X @ (R - np.eye(g)).T + 0.3)
    return x, x.copy()


@pytest.mark.parametrize("gdim", [2, 3])
def test_solid_residual_is_objective(gdim):
    """St.Venant-Kirchhoff stress (common.py:104-109) depends on E = (F^T F - I)/2 only: a rigid rotation of any
    size (F = R) at rest gives E = 0, P = 0 — the whole solid residual (solid.py:41-43, 62-63) vanishes."""
    from oracle.backend import OracleAssembler
    from tests.conftest import TUBE_PARAMS
    from tests.helpers import square_mesh
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = square_mesh(5) if gdim == 2 else box_mesh(3)
    mesh.markers[:] = 2
    space = MixedSpace(mesh)
    regions = [dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703)]
    x, x1 = _rigid_rotation_state(mesh, space)
    for use_c in (True, False):
        ora = OracleAssembler(mesh, space, regions, dict(TUBE_PARAMS), 4, 10 if gdim == 2 else 6, use_c_port=use_c)
        b = ora.assemble_residual(x, x1)
        # scale: the stress a 1 % stretch would produce on one cell face
        assert np.abs(b).max() < 1e-12 * 3103448.0
        # ... and a stretch is NOT free
        xs = x.copy()
        xs[space.node_dof(0, 0)] *= 1.01
        assert np.abs(ora.assemble_residual(xs, xs.copy())).max() > 1.0


def test_translation_of_the_displacement_changes_nothing():
    """only grad d and d - d1 enter the forms (fluid.py, solid.py, laplace.py): a rigid translation added to d at both
    time levels leaves residual and Jacobian unchanged (FSI tube, oracle C port)."""
    from oracle.backend import OracleAssembler
    from tests.conftest import TUBE_PARAMS, TUBE_REGIONS
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = tube_mesh(3)
    space = MixedSpace(mesh)
    x, x1 = synthetic_state(mesh, space, 5, h=0.5), synthetic_state(mesh, space, 6, h=0.5)
    xs, x1s = x.copy(), x1.copy()
    for a, c in enumerate((0.4, -0.7, 1.3)):
        xs[space.node_dof(0, a)] += c
        x1s[space.node_dof(0, a)] += c
    ora = OracleAssembler(mesh, space, TUBE_REGIONS, dict(TUBE_PARAMS), 4, 6, use_c_port=True)
    b, bs = ora.assemble_residual(x, x1), ora.assemble_residual(xs, x1s)
    assert np.abs(b - bs).max() < 1e-12 * np.abs(b).max()
    A, As = ora.assemble_jacobian(x, x1), ora.assemble_jacobian(xs, x1s)
    assert np.abs(A - As).max() < 1e-12 * np.abs(A).max()


def test_visualization_output_is_a_readable_xdmf_series(tmp_path):
    """problems/__init__.py:222-323 + monolithic.py:192-194: displacement / velocity / pressure series, one stored step
    per `save_step` steps; save_deg 1 = vertex values, save_deg 2 = the P2 fields as P1 fields on the refined mesh (the
    pressure stays on the original mesh). Heavy data is raw binary (no h5py here); read back with utils.xdmf.read_series."""
    from turtlefsi_b200.utils.xdmf import read_series
    for save_deg in (1, 2):
        folder = str(tmp_path / ("viz%d" % save_deg))
        ns = _run(4, T=0.03, visualization_folder=folder, save_step=2, save_deg=save_deg)
        mesh, space = ns["mesh"], ns["DVP"]
        x = np.asarray(ns["dvp_"]["n"])
        series = {k: read_series(os.path.join(folder, k + ".xdmf")) for k in ("displacement", "velocity", "pressure")}
        for k, (coords, cells, steps) in series.items():
            assert [round(t / 0.01) for t, _, _ in steps] == [1, 3]            # counters 0 and 2 of the four steps
            assert {name for _, name, _ in steps} == {k.capitalize()}
        coords, cells, steps = series["velocity"]
        if save_deg == 1:
            assert np.array_equal(coords, mesh.coords) and np.array_equal(cells, mesh.cells)
            nodes = space.vertex_node
        else:
            assert coords.shape[0] == space.num_nodes and cells.shape[0] == 4 * mesh.num_cells
            assert np.allclose(coords, space.node_coords(mesh))
            X = coords[cells]                                               # the sub-triangles tile every cell
            area = 0.5 * np.abs(np.linalg.det(X[:, 1:] - X[:, :1]))
            assert abs(area.sum() - 4.0) < 1e-12 and area.min() > 0.24 * area.max()
            nodes = np.arange(space.num_nodes)
        # ns holds the LAST step (counter 3); the stored ones are earlier, so compare structure and magnitudes only ...
        assert steps[-1][2].shape == (len(nodes), 3) and np.all(steps[-1][2][:, 2] == 0.0)
        assert np.abs(steps[-1][2][:, :2]).max() < 2 * np.abs(x[space.sub_dofs(1)]).max() + 1e-12
        pc, pcells, psteps = series["pressure"]
        assert np.array_equal(pc, mesh.coords) and np.array_equal(pcells, mesh.cells) and psteps[0][2].shape == (mesh.num_vertices, 1)
    # ... and the values exactly on a run whose last step IS a stored one
    folder = str(tmp_path / "exact")
    ns = _run(4, T=0.02, visualization_folder=folder, save_step=2, save_deg=2)
    x, space = np.asarray(ns["dvp_"]["n"]), ns["DVP"]
    _, _, steps = read_series(os.path.join(folder, "velocity.xdmf"))
    assert len(steps) == 2
    for a in range(2):
        assert np.array_equal(steps[-1][2][:, a], x[space.node_dof(1, a)])
    _, _, psteps = read_series(os.path.join(folder, "pressure.xdmf"))
    assert np.array_equal(psteps[-1][2][:, 0], x[space.vertex_dof_p()])
    import pytest
    with pytest.raises(ValueError, match="save_deg"):
        _run(4, T=0.01, visualization_folder=str(tmp_path / "bad"), save_step=1, save_deg=3)


def test_refined_visualization_cells_tile_the_p2_cells_in_3d():
    """save_deg = 2 in 3-D: every P2 tetrahedron becomes eight P1 tetrahedra on its ten nodes, each an eighth of it."""
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    from turtlefsi_b200.utils.xdmf import p2_refined_cells
    mesh = box_mesh(2)
    space = MixedSpace(mesh)
    sub = p2_refined_cells(space.cell_nodes, 3)
    X = space.node_coords(mesh)[sub]
    vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1])) / 6
    Xp = mesh.coords[mesh.cells]
    parent = np.abs(np.linalg.det(Xp[:, 1:] - Xp[:, :1])) / 6
    assert sub.shape == (8 * mesh.num_cells, 4) and abs(vol.sum() - 1.0) < 1e-13
    assert np.allclose(vol.reshape(8, -1), parent[None, :] / 8, rtol=1e-12)


def test_kill_and_pause_files_stop_and_hold_the_run(tmp_path):
    """problems/__init__.py:326-358 + monolithic.py:203-212: `killturtle` in the results folder (or `killtime`) ends the
    run cleanly behind a checkpoint, `pauseturtle` holds it until the file is removed."""
    import threading
    import time
    folder = tmp_path / "run"
    folder.mkdir()
    (folder / "killturtle").write_text("")
    ns = _run(4, T=0.05, results_folder=str(folder), checkpoint_step=1000)
    assert len(ns["error_v_list"]) == 1 and not (folder / "killturtle").exists()
    assert os.path.exists(folder / "Checkpoint" / "checkpoint_fields.npy")          # written by the kill, not by the step count
    again = _run(4, T=0.02, restart_folder=str(folder))
    assert again["first_step_num"] == 1
    timed = _run(4, T=0.05, results_folder=str(tmp_path / "timed"), killtime=0.0)
    assert len(timed["error_v_list"]) == 1
    hold = tmp_path / "hold"
    hold.mkdir()
    (hold / "pauseturtle").write_text("")
    threading.Timer(0.6, lambda: (hold / "pauseturtle").unlink()).start()
    t0 = time.time()
    ns = _run(4, T=0.005, results_folder=str(hold), pause_poll_seconds=0.05)
    assert time.time() - t0 > 0.5 and len(ns["error_v_list"]) == 1


def test_restart_keeps_and_merges_the_visualization_series(tmp_path):
    """problems/__init__.py:147-170, 414-461 + monolithic.py:224-227: a restarted run renames the series it finds
    (`*_run_1`), writes its own and joins both at the end: one file per field, steps in time order, the step the restart
    recomputed taken from the restarted run, equal to an uninterrupted run to the Newton tolerance."""
    from tests.test_turek_pin import _run_oracle
    from turtlefsi_b200.utils.xdmf import read_series
    ov = dict(N=4, dt=0.01, folder=str(tmp_path / "res"), sub_folder="a", save_step=1, checkpoint_step=2)
    _run_oracle("TaylorGreen2D", dict(ov, T=0.03), create_results_folder=True)                  # steps t = .01 ... .04
    viz = tmp_path / "res" / "a" / "Visualization"
    assert len(read_series(str(viz / "velocity.xdmf"))[2]) == 4
    again = _run_oracle("TaylorGreen2D", dict(ov, T=0.05, restart_folder=str(tmp_path / "res" / "a")), create_results_folder=True)
    assert again["run_number"] == 1 and again["first_step_num"] == 2       # checkpoint of counter 2: t = .03 is recomputed
    assert sorted(f.name for f in viz.iterdir()) == ["displacement.bin", "displacement.xdmf", "displacement_run_1.bin",
                                                     "pressure.bin", "pressure.xdmf", "pressure_run_1.bin",
                                                     "velocity.bin", "velocity.xdmf", "velocity_run_1.bin"]
    coords, cells, steps = read_series(str(viz / "velocity.xdmf"))
    assert np.allclose([t for t, _, _ in steps], [0.01, 0.02, 0.03, 0.04, 0.05, 0.06])
    full = _run_oracle("TaylorGreen2D", dict(N=4, dt=0.01, T=0.05, visualization_folder=str(tmp_path / "full"), save_step=1))
    ref = read_series(str(tmp_path / "full" / "velocity.xdmf"))[2]
    for (t, _, v), (tr, _, vr) in zip(steps, ref):
        assert abs(t - tr) < 1e-12 and np.abs(v - vr).max() < 1e-7 * np.abs(vr).max()
    assert len(read_series(str(viz / "pressure.xdmf"))[2]) == 6
