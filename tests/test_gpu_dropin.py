"""GPU twins of tests/test_dropin.py: problem files written against DOLFIN's API, run UNMODIFIED through
`python -m turtlefsi_b200` machinery with the CUDA assembler (turtlefsi_b200.run_turtle.make_b200) behind the hooks."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(problem, overrides, linear_solver="superlu"):
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.run_turtle import make_b200
    ma, ms = make_b200(linear_solver)
    return monolithic.run(problem, ma, ms, overrides=overrides)


def test_user_file_in_dolfin_dialect_on_the_cuda_path():
    """tests/dolfin_dialect/turek_flag.py: Turek FSI with the laplace lifting; drag / lift / tip displacement from the
    file's own `assemble(... ds(6))`, `dS(5)("+")` and point evaluation == the committed oracle history (1e-8) and the
    reference's goldens (its own tolerance)."""
    ns = _run(os.path.join(HERE, "dolfin_dialect", "turek_flag.py"), dict(T=0.05, dt=0.01, verbose=False))
    h = ns["history"]
    ref = json.load(open(os.path.join(HERE, "golden", "turek_oracle_histories.json")))["tf_fsi_laplace_constant"]
    g = json.load(open(os.path.join(HERE, "golden", "turek_goldens.json")))["TF_fsi"]
    assert len(h["drag"]) == 6
    assert np.allclose(h["drag"], ref["drag"], rtol=1e-8)
    assert np.allclose(h["lift"], ref["lift"], rtol=1e-6, atol=1e-12)
    assert np.allclose(h["dx"], ref["dis_x"], rtol=1e-7, atol=1e-16)
    assert np.isclose(h["drag"][-1], g["drag"]) and np.isclose(h["dx"][-1], g["dis_x"])
    assert np.isclose(h["lift"][-1], g["lift"], rtol=1e-4)                       # tests/test_turtleFSI.py:80
    assert ns["assembler"].launch_count() > 0


def test_form_edited_by_the_problem_file_on_the_cuda_path():
    """`F_solid_linear += p(t) * inner(n('+'), psi('+')) * dS(fsi_id)` in a DOLFIN-dialect file (problem_aneu.py:214) ==
    the hand-registered load of the native tube3d problem, both through the CUDA kernels (tfsi_set_facet_load)."""
    ov = dict(T=0.002, dt=0.001, verbose=False, ncube=3)
    a = _run(os.path.join(HERE, "dolfin_dialect", "tube_pressure.py"), ov)
    b = _run("tube3d", dict(ov, inner_pressure=True))
    ha, hb = np.array(a["residual_history"]), np.array(b["residual_history"])
    big = hb[:, 2] > 1e-4 * hb[:, 2].max()
    assert len(a["facet_loads"]) == 1 and ha.shape == hb.shape and big.sum() >= 3
    assert np.allclose(ha[big, 2], hb[big, 2], rtol=1e-9)
    assert np.allclose(np.array(a["probe_d"]), np.array(b["probe_d"]), rtol=1e-7, atol=1e-18)


def test_command_line_entry_point_on_the_cuda_path(tmp_path, monkeypatch):
    """`python -m turtlefsi_b200 --problem TF_cfd -dt 0.01 -T 0.02 --theta 0.5` (the shipped twin), device LU: runs, writes
    the result files, no fallback."""
    from turtlefsi_b200.run_turtle import main
    monkeypatch.chdir(tmp_path)
    ns = main(["--problem", "TF_cfd", "-dt", "0.01", "-T", "0.02", "--theta", "0.5", "--verbose", "False", "--sub-folder", "r"])
    ref = json.load(open(os.path.join(HERE, "golden", "turek_oracle_histories.json")))["tf_cfd"]
    assert np.allclose(ns["drag_list"], ref["drag"][:len(ns["drag_list"])], rtol=1e-8)
    assert (tmp_path / "TF_cfd_results" / "r" / "Drag.txt").exists()
    # the visualisation series of the first step (save_step = 10), read back from the raw-binary XDMF files
    from turtlefsi_b200.utils.xdmf import read_series
    coords, cells, steps = read_series(str(tmp_path / "TF_cfd_results" / "r" / "Visualization" / "velocity.xdmf"))
    assert cells.shape[0] == ns["mesh"].num_cells and len(steps) == 1 and np.isfinite(steps[0][2]).all()
    assert np.abs(steps[0][2]).max() > 0
