"""GPU: the re-hosted Turek problem files driven through the CUDA path (C-ABI) — Newton residual
histories and benchmark scalars against the stored ORACLE runs
(tests/golden/turek_oracle_histories.json, made by tests/golden/make_goldens.py) at 1e-8 relative
(BASELINE north_star), and against the reference's own goldens."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run_gpu(problem, **over):
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.linear_solver import HostLU

    def make_asm(mesh, space, regions, ns):
        return FSIAssembler(mesh, space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"], solid=ns["solid"],
                            extrapolation=ns["extrapolation"], extrapolation_sub_type=ns["extrapolation_sub_type"],
                            quad_degree_jacobian=ns["compiler_parameters"]["quadrature_degree"],
                            quad_degree_residual=ns.get("residual_quadrature_degree", 10), gravity=ns["gravity"])
    ov = dict(T=0.05, dt=0.01, verbose=False)
    ov.update(over)
    return monolithic.run(problem, make_asm,
                          lambda a: HostLU(a.rowptr.cpu().numpy(), a.colidx.cpu().numpy(), a.num_dofs), overrides=ov)


def _compare(ns, hist, keys, rtol=1e-8):
    got = ns["residual_history"]
    assert len(got) == len(hist["residual_history"])
    for (c, i, r, rr), (c0, i0, r0, rr0) in zip(got, hist["residual_history"]):
        assert (c, i) == (c0, i0)
        if r0 > 1e-3 and rr0 > 1e-3:   # converged iterates sit at round-off level: only the count is compared
            assert abs(r / r0 - 1) < rtol and abs(rr / rr0 - 1) < rtol
    names = dict(drag="drag_list", lift="lift_list", dis_x="displacement_x_list", dis_y="displacement_y_list")
    for k in keys:
        assert np.allclose(ns[names[k]], hist[k], rtol=rtol, atol=0), k


def _fixtures():
    return (json.load(open(os.path.join(HERE, "golden", "turek_oracle_histories.json"))),
            json.load(open(os.path.join(HERE, "golden", "turek_goldens.json"))))


def test_tf_fsi_file_default_biharmonic():
    hist, gold = _fixtures()
    ns = _run_gpu("TF_fsi", theta=0.51)
    _compare(ns, hist["tf_fsi_biharmonic_default"], ("drag", "lift", "dis_x"))
    assert np.allclose(ns["displacement_y_list"], hist["tf_fsi_biharmonic_default"]["dis_y"], rtol=1e-5)   # |dis_y| ~ 1e-9
    g = gold["TF_fsi"]
    assert abs(ns["drag_list"][-1] / g["drag"] - 1) < 1e-10            # reference golden, round-off level
    assert abs(ns["displacement_x_list"][-1] / g["dis_x"] - 1) < 1e-9
    assert abs(ns["lift_list"][-1] / g["lift"] - 1) < 1e-7
    # the Newton loop alternated between the two matrices of a pair, re-initialised from inside the fluid kernel
    assert ns["A_pair"] is not None and ns["assembler"].fused_launches() >= 5


def test_tf_fsi_laplace_constant():
    hist, gold = _fixtures()
    ns = _run_gpu("TF_fsi", theta=0.51, extrapolation="laplace", extrapolation_sub_type="constant")
    _compare(ns, hist["tf_fsi_laplace_constant"], ("drag", "lift", "dis_x"))
    g = gold["TF_fsi"]
    assert np.isclose(ns["drag_list"][-1], g["drag"]) and np.isclose(ns["lift_list"][-1], g["lift"], rtol=1e-4)


@pytest.mark.parametrize("ext,sub", [("laplace", "volume"), ("laplace", "small_constant"),
                                      ("biharmonic", "constrained_disp")])
def test_tf_fsi_other_lifting_variants_hit_reference_goldens(ext, sub):
    """tests/test_turtleFSI.py:61-103: every lifting variant shares one set of goldens."""
    _, gold = _fixtures()
    ns = _run_gpu("TF_fsi", theta=0.51, extrapolation=ext, extrapolation_sub_type=sub)
    g = gold["TF_fsi"]
    assert np.isclose(ns["displacement_x_list"][-1], g["dis_x"])
    assert np.isclose(ns["drag_list"][-1], g["drag"])
    assert np.isclose(ns["lift_list"][-1], g["lift"], rtol=1e-4 if ext == "laplace" else 1e-5)


def test_tf_fsi_laplace_volume_change():
    """tests/test_turtleFSI.py:61-80 with "volume_change" (alfa = 1/J(d), laplace.py:33-34): CUDA path vs the stored
    numpy-oracle run (the C port does not have this variant) and the reference's goldens."""
    hist, gold = _fixtures()
    ns = _run_gpu("TF_fsi", theta=0.51, extrapolation="laplace", extrapolation_sub_type="volume_change")
    _compare(ns, hist["tf_fsi_laplace_volume_change_numpy_oracle"], ("drag", "lift", "dis_x"))
    g = gold["TF_fsi"]
    assert np.isclose(ns["displacement_x_list"][-1], g["dis_x"]) and np.isclose(ns["displacement_y_list"][-1], g["dis_y"])
    assert np.isclose(ns["drag_list"][-1], g["drag"]) and np.isclose(ns["lift_list"][-1], g["lift"], rtol=1e-4)


def test_tf_fsi_elastic_lifting():
    """test_elastic of the reference (tests/test_turtleFSI.py:106-123: `--extrapolation elastic -et constant`, its own
    goldens :115-118): CUDA path (elastic.py:18-27 in the residual kernel, its Jacobian frozen into A_pre) vs the stored
    numpy-oracle run at 1e-8 and vs the reference's goldens."""
    hist, gold = _fixtures()
    ns = _run_gpu("TF_fsi", theta=0.51, extrapolation="elastic", extrapolation_sub_type="constant")
    h = hist["tf_fsi_elastic_numpy_oracle"]
    assert np.allclose(ns["drag_list"], h["drag"], rtol=1e-8, atol=0)
    assert np.allclose(ns["displacement_x_list"], h["dis_x"], rtol=1e-8, atol=0)
    assert np.allclose(ns["lift_list"], h["lift"], rtol=1e-6, atol=0)
    g = gold["TF_fsi_elastic"]
    assert np.isclose(ns["displacement_x_list"][-1], g["dis_x"]) and np.isclose(ns["displacement_y_list"][-1], g["dis_y"])
    assert np.isclose(ns["drag_list"][-1], g["drag"]) and np.isclose(ns["lift_list"][-1], g["lift"])
    assert abs(ns["drag_list"][-1] / g["drag"] - 1) < 1e-10


def test_tf_cfd():
    hist, gold = _fixtures()
    ns = _run_gpu("TF_cfd", theta=0.5)
    _compare(ns, hist["tf_cfd"], ("drag", "lift"))
    assert abs(ns["drag_list"][-1] / gold["TF_cfd"]["drag"] - 1) < 1e-10
    assert abs(ns["lift_list"][-1] / gold["TF_cfd"]["lift"] - 1) < 1e-7


def test_tf_csm():
    hist, gold = _fixtures()
    ns = _run_gpu("TF_csm")
    _compare(ns, hist["tf_csm"], ("dis_x", "dis_y"), rtol=1e-7)
    assert np.isclose(ns["displacement_x_list"][-1], gold["TF_csm"]["dis_x"])
    assert np.isclose(ns["displacement_y_list"][-1], gold["TF_csm"]["dis_y"])
    assert ns["A_pair"] is not None and ns["assembler"].fused_launches() == 0      # no fluid cells: foreground path


def test_tube3d_newton_matches_oracle():
    """BASELINE config 4 at test scale: 3-D tube FSI (problem_aneu parameters), Jacobian re-used over the
    iterations (recompute=3000): CUDA path vs the ORACLE (plain-C port) driven by the same Newton loop."""
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_oracle(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, 4, 6, use_c_port=True)
    ov = dict(T=0.002, dt=0.001, verbose=False, ncube=4)
    ref = monolithic.run("tube3d", make_oracle, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs), overrides=ov)
    ns = _run_gpu("tube3d", **ov)
    assert len(ns["residual_history"]) == len(ref["residual_history"]) > 6
    rmax = max(h[2] for h in ref["residual_history"])
    rrmax = max(h[3] for h in ref["residual_history"])
    for (c, i, r, rr), (c0, i0, r0, rr0) in zip(ns["residual_history"], ref["residual_history"]):
        assert (c, i) == (c0, i0)
        # with the Jacobian re-used, iterates after the first sit at the round-off floor of the
        # (delta = 1e7 scaled) linear solve, ~1e-8 of the first residual: only entries well above are compared
        if r0 > 1e-4 * rmax:
            assert abs(r / r0 - 1) < 1e-8
        if rr0 > 1e-4 * rrmax:
            assert abs(rr / rr0 - 1) < 1e-8
    assert np.allclose(np.array(ns["probe_v"]), np.array(ref["probe_v"]), rtol=1e-8, atol=1e-14)
    assert np.allclose(np.array(ns["probe_d"]), np.array(ref["probe_d"]), rtol=1e-7, atol=1e-16)
    assert abs(np.array(ns["probe_v"])[-1, 2]) > 1e-4          # the flow actually moves


def test_turtle_demo_two_steps_match_oracle():
    """BASELINE config 3 (the reference's default problem): turtle_demo, biharmonic "constrained_disp",
    Jacobian re-used (recompute = 15) — CUDA path vs the ORACLE driven by the same problem file."""
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_oracle(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10,
                               use_c_port=True)
    ref = monolithic.run("turtle_demo", make_oracle, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs),
                         overrides=dict(verbose=False), max_steps=2)
    from turtlefsi_b200.assembler import FSIAssembler

    def make_gpu(mesh, space, regions, ns):
        return FSIAssembler(mesh, space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"], solid=ns["solid"],
                            extrapolation=ns["extrapolation"], extrapolation_sub_type=ns["extrapolation_sub_type"],
                            quad_degree_jacobian=ns["compiler_parameters"]["quadrature_degree"], quad_degree_residual=10)
    ns = monolithic.run("turtle_demo", make_gpu,
                        lambda a: HostLU(a.rowptr.cpu().numpy(), a.colidx.cpu().numpy(), a.num_dofs),
                        overrides=dict(verbose=False), max_steps=2)
    assert ns["DVP"].num_dofs == 83012
    assert len(ns["residual_history"]) == len(ref["residual_history"])
    rmax = max(h[2] for h in ref["residual_history"])
    for (c, i, r, rr), (c0, i0, r0, rr0) in zip(ns["residual_history"], ref["residual_history"]):
        assert (c, i) == (c0, i0)
        if r0 > 1e-4 * rmax:
            assert abs(r / r0 - 1) < 1e-8
    x, x0 = ns["dvp_"]["n"].cpu().numpy(), ref["dvp_"]["n"]
    sp = ns["DVP"]
    for field in (0, 1, 2, 3):
        d = sp.sub_dofs(field)
        assert np.linalg.norm(x[d] - x0[d]) < 1e-8 * np.linalg.norm(x0[d]), field


def test_taylor_green_matches_oracle_and_the_analytic_solution():
    """TaylorGreen2D (the reference's DOLFIN-free known answer): CUDA path vs the oracle run of the same
    problem file, and the error against the analytic solution."""
    from tests.test_turek_pin import _run_oracle
    ov = dict(N=16, T=0.1, dt=0.01)
    ref = _run_oracle("TaylorGreen2D", ov)
    ns = _run_gpu("TaylorGreen2D", **ov)
    assert len(ns["residual_history"]) == len(ref["residual_history"])
    assert np.allclose(ns["error_v_list"], ref["error_v_list"], rtol=1e-8, atol=0)
    assert np.allclose(ns["error_p_list"], ref["error_p_list"], rtol=1e-8, atol=0)
    assert ns["error_v_list"][-1] < 5e-3 and ns["error_p_list"][-1] < 7e-2


def test_robin_validation_and_inner_pressure_match_oracle():
    """8f-1 facet terms end to end: RobinBC_validation (solid only, Robin wall, gravity) and the 3-D tube with
    the interior-facet pressure load of problem_aneu.py:212-214 — CUDA path vs oracle, same problem files."""
    from tests.test_known_answers import STIFF, _oracle_run
    ov = dict(dt=0.005, T=0.02, ncube=3, **STIFF)
    ref = _oracle_run("RobinBC_validation", **ov)
    ns = _run_gpu("RobinBC_validation", **ov)
    assert len(ns["residual_history"]) == len(ref["residual_history"])
    assert np.allclose(ns["d_list"], ref["d_list"], rtol=1e-8, atol=0)
    a = ns["default_variables"]["analytical_displacement"]
    assert np.abs(np.array(ns["d_list"]) - a).max() < 2e-3 * np.abs(a).max()      # the ODE itself
    ov = dict(T=0.003, dt=0.001, ncube=3, inner_pressure=True)
    ref = _oracle_run("tube3d", **ov)
    ns = _run_gpu("tube3d", **ov)
    base = _run_gpu("tube3d", **dict(ov, inner_pressure=False))
    assert len(ns["residual_history"]) == len(ref["residual_history"])
    assert np.allclose(np.array(ns["probe_v"]), np.array(ref["probe_v"]), rtol=1e-7, atol=1e-12)
    assert np.allclose(np.array(ns["probe_d"]), np.array(ref["probe_d"]), rtol=1e-6, atol=1e-14)
    assert np.abs(np.array(ns["probe_d"])[-1] - np.array(base["probe_d"])[-1]).max() > 1e-8   # the load acts


def test_tf_fsi_with_the_device_refactorisation_lu():
    """the linear solve on the device with the reference's re-use protocol — analysis and pivoting
    once on the host, numeric re-factorisation (cusolverRf) at every re-assembly, solve-only on re-use; every solve
    is residual-checked (tfsi_csr_matvec) and refined. Same Newton histories and goldens as with the host LU."""
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.linear_solver import DeviceRfLU
    made = {}

    def make_asm(mesh, space, regions, ns):
        return FSIAssembler(mesh, space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"], solid=ns["solid"],
                            extrapolation=ns["extrapolation"], extrapolation_sub_type=ns["extrapolation_sub_type"],
                            quad_degree_jacobian=ns["compiler_parameters"]["quadrature_degree"], quad_degree_residual=10,
                            gravity=ns["gravity"])

    def make_solver(a):
        made["lu"] = DeviceRfLU(a)
        return made["lu"]
    ns = monolithic.run("TF_fsi", make_asm, make_solver, overrides=dict(T=0.05, dt=0.01, theta=0.51, verbose=False))
    hist, gold = _fixtures()
    _compare(ns, hist["tf_fsi_biharmonic_default"], ("drag", "lift", "dis_x"))
    g = gold["TF_fsi"]
    assert abs(ns["drag_list"][-1] / g["drag"] - 1) < 1e-10 and abs(ns["lift_list"][-1] / g["lift"] - 1) < 1e-7
    lu = made["lu"]
    iters = len(ns["residual_history"])
    assert lu.solves == iters and lu.last_relative_residual < 1e-11
    assert lu.host_analyses <= 3 < iters                    # pivoting is NOT redone per iteration ...
    assert lu.refactorizations >= iters                     # ... the numeric factorisation is (TF_fsi: recompute = 1)
