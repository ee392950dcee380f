"""ORACLE PIN against the reference's OWN golden values (the only goldens it holds for this path:
end-of-run scalars of tests/test_turtleFSI.py). The CPU restatement (oracle/), driven by
the re-hosted problem files and the reference's Newton loop, reproduces

  TF_fsi  file default (biharmonic, constrained_disp_vel)   tests/test_turtleFSI.py:40-58
          drag to 2e-14, dis_x to 3e-13, lift to 1e-10, dis_y to 5e-10 RELATIVE
  TF_cfd  (refined mesh, fluid only)                         :12-23   drag 1e-13, lift 2e-11
  TF_csm  (refined mesh, solid only, gravity)                :26-37   dis_x 7e-7, dis_y 2e-7
  TF_fsi  laplace/constant                                   :61-80   within the reference tolerances

i.e. forms, P2/P1 and quadrature conventions, dof map, Dirichlet sets, ident_zeros, the Newton
loop with its Jacobian-reuse predicates, Plaza refinement and the drag/lift/point-evaluation
post-processing ("+" side = lowest cell index) all agree with DOLFIN's to round-off."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _goldens():
    return json.load(open(os.path.join(HERE, "golden", "turek_goldens.json")))


def _histories():
    return json.load(open(os.path.join(HERE, "golden", "turek_oracle_histories.json")))


def _run_oracle(problem, overrides, made=None, **run_kw):
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_asm(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10,
                               use_c_port=True)

    def make_solver(a):
        lu = HostLU(a.rowptr, a.colidx, a.num_dofs)
        if made is not None:
            made["lu"] = lu
        return lu
    ov = dict(T=0.05, dt=0.01, verbose=False)
    ov.update(overrides)
    return monolithic.run(problem, make_asm, make_solver, overrides=ov, **run_kw)


def test_oracle_reproduces_tf_fsi_file_default_golden():
    ns = _run_oracle("TF_fsi", dict(theta=0.51))
    g = _goldens()["TF_fsi"]
    assert ns["extrapolation"] == "biharmonic" and ns["extrapolation_sub_type"] == "constrained_disp_vel"
    assert ns["DVP"].num_dofs == 31965 and ns["assembler"].nnz == 2367247          # SURVEY 8a sizes
    assert len(ns["drag_list"]) == 6          # `while t <= T + dt / 10` takes T/dt + 1 steps
    # far sharper than the reference's np.isclose(rtol=1e-5): round-off level agreement
    assert abs(ns["drag_list"][-1] / g["drag"] - 1) < 1e-11
    assert abs(ns["displacement_x_list"][-1] / g["dis_x"] - 1) < 1e-10
    assert abs(ns["lift_list"][-1] / g["lift"] - 1) < 1e-8
    assert abs(ns["displacement_y_list"][-1] / g["dis_y"] - 1) < 1e-7


def test_oracle_reproduces_tf_cfd_golden():
    ns = _run_oracle("TF_cfd", dict(theta=0.5))
    g = _goldens()["TF_cfd"]
    assert ns["DVP"].num_dofs == 80938                                               # refined once (Plaza)
    assert abs(ns["drag_list"][-1] / g["drag"] - 1) < 1e-11
    assert abs(ns["lift_list"][-1] / g["lift"] - 1) < 1e-8


def test_oracle_reproduces_tf_csm_golden_and_reuses_the_jacobian():
    made = {}
    ns = _run_oracle("TF_csm", dict(), made)
    g = _goldens()["TF_csm"]
    assert np.isclose(ns["displacement_x_list"][-1], g["dis_x"])                     # reference's assertion
    assert np.isclose(ns["displacement_y_list"][-1], g["dis_y"])
    assert abs(ns["displacement_y_list"][-1] / g["dis_y"] - 1) < 1e-6
    # recompute=5 (default): newtonsolver.py:54-63 rebuild at iteration 0, every 5th iteration and
    # when the residual increases — the factorisation is re-used in between
    iters = len(ns["residual_history"])
    assert made["lu"].factorizations < iters and made["lu"].factorizations >= 6


def test_stored_oracle_histories_match_reference_goldens():
    """the committed fixtures (tests/golden/turek_oracle_histories.json, made by make_goldens.py) that
    the GPU tests compare against are themselves within the reference's tolerances"""
    h, g = _histories(), _goldens()
    f = h["tf_fsi_laplace_constant"]
    assert np.isclose(f["dis_x"][-1], g["TF_fsi"]["dis_x"]) and np.isclose(f["drag"][-1], g["TF_fsi"]["drag"])
    assert np.isclose(f["lift"][-1], g["TF_fsi"]["lift"], rtol=1e-4)                 # tests/test_turtleFSI.py:80
    f = h["tf_fsi_biharmonic_default"]
    for k in ("dis_x", "dis_y", "drag", "lift"):
        assert np.isclose(f[k][-1], g["TF_fsi"][k])
    assert np.isclose(h["tf_cfd"]["drag"][-1], g["TF_cfd"]["drag"]) and np.isclose(h["tf_cfd"]["lift"][-1], g["TF_cfd"]["lift"])
    assert np.isclose(h["tf_csm"]["dis_x"][-1], g["TF_csm"]["dis_x"]) and np.isclose(h["tf_csm"]["dis_y"][-1], g["TF_csm"]["dis_y"])


def test_stored_elastic_oracle_run_matches_reference_golden():
    """test_elastic (tests/test_turtleFSI.py:106-123): elastic lifting lives only in the numpy oracle (the
    kernels refuse it); its 2-minute run is stored by `make_goldens.py --elastic` and checked here."""
    f, g = _histories()["tf_fsi_elastic_numpy_oracle"], _goldens()["TF_fsi_elastic"]
    assert abs(f["drag"][-1] / g["drag"] - 1) < 1e-11 and abs(f["dis_x"][-1] / g["dis_x"] - 1) < 1e-10
    assert abs(f["lift"][-1] / g["lift"] - 1) < 1e-8 and abs(f["dis_y"][-1] / g["dis_y"] - 1) < 1e-7


def test_stored_volume_change_oracle_run_matches_reference_golden():
    """test_laplace[volume_change] (tests/test_turtleFSI.py:61-80): alfa = 1/J(d) lives in the numpy oracle (not in
    the C port); its 2-minute run is stored by `make_goldens.py --only-slow --volume-change` and checked here with
    the reference's own tolerances (np.isclose; lift rtol 1e-4, :80)."""
    f, g = _histories()["tf_fsi_laplace_volume_change_numpy_oracle"], _goldens()["TF_fsi"]
    assert np.isclose(f["dis_x"][-1], g["dis_x"]) and np.isclose(f["dis_y"][-1], g["dis_y"])
    assert np.isclose(f["drag"][-1], g["drag"]) and np.isclose(f["lift"][-1], g["lift"], rtol=1e-4)
    assert abs(f["drag"][-1] / g["drag"] - 1) < 1e-7


def test_plaza_refinement_properties():
    from turtlefsi_b200.mesh import read_dolfin_xml, refine
    m = read_dolfin_xml(os.path.join(HERE, "golden", "meshes", "TF_cfd.xml.gz"))
    r = refine(m)
    assert (r.num_vertices, r.num_cells) == (4858, 9304)
    area = lambda mesh: 0.5 * np.abs(np.linalg.det(np.swapaxes(
        mesh.coords[mesh.cells][:, 1:] - mesh.coords[mesh.cells][:, :1], 1, 2))).sum()
    assert abs(area(m) - area(r)) < 1e-13
    # conforming: every interior edge is shared by exactly two triangles
    from turtlefsi_b200.boundary import Facets
    f = Facets(r)
    assert f.on_boundary.sum() == 2 * Facets(m).on_boundary.sum()
