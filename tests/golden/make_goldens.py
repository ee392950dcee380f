"""Writes the golden fixtures of tests/golden (run here, where /root/reference is readable):
  turek_goldens.json          scalars pinned by the reference's own tests (transcribed, with the
                              reference file:line of each)
  tf_fsi_oracle_history.json  TF_fsi (laplace/constant, theta 0.51, dt 0.01, 6 steps) run through the
                              ORACLE backend (plain-C restatement, cross-checked against the numpy
                              complex-step forms): Newton residual history + per-step scalars.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

GOLDENS = {
    "source": "KVSlab/turtleFSI v2.4 tests/test_turtleFSI.py (np.isclose default rtol=1e-5, atol=1e-8 unless noted)",
    "TF_cfd": {"drag": 4.503203576965564, "lift": -0.03790359084395478, "ref": "tests/test_turtleFSI.py:19-20"},
    "TF_csm": {"dis_x": -3.313014369394714527e-05, "dis_y": -3.770127311444726199e-03, "ref": "tests/test_turtleFSI.py:33-34"},
    "TF_fsi": {"dis_x": -6.896013956339182e-06, "dis_y": 1.876355330341896e-09, "drag": 4.407481239804155,
               "lift": -0.005404703556977697, "ref": "tests/test_turtleFSI.py:50-53 (also :72-75 laplace x4, lift rtol 1e-4 at :80; :95-98 biharmonic x2)"},
    "TF_fsi_elastic": {"dis_x": -6.896144755254494e-06, "dis_y": 1.868651990487361e-09, "drag": 4.407488867909029,
                       "lift": -0.005404616050528832, "ref": "tests/test_turtleFSI.py:115-118"},
}


CASES = {
    # name: (problem, overrides, golden key)  — the command lines of the reference's tests
    "tf_fsi_laplace_constant": ("TF_fsi", dict(theta=0.51, extrapolation="laplace", extrapolation_sub_type="constant"), "TF_fsi"),
    "tf_fsi_biharmonic_default": ("TF_fsi", dict(theta=0.51), "TF_fsi"),          # tests/test_turtleFSI.py:40-58
    "tf_cfd": ("TF_cfd", dict(theta=0.5), "TF_cfd"),                              # :12-23
    "tf_csm": ("TF_csm", dict(), "TF_csm"),                                       # :26-37
}


def run_oracle(problem, overrides):
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_asm(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10,
                               use_c_port=True)
    ov = dict(T=0.05, dt=0.01, verbose=False)
    ov.update(overrides)
    ns = monolithic.run(problem, make_asm, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs), overrides=ov)
    out = {"problem": problem, "overrides": {k: v for k, v in ov.items()},
           "residual_history": [[int(c), int(i), float(r), float(rr)] for c, i, r, rr in ns["residual_history"]],
           "num_dofs": int(ns["DVP"].num_dofs), "nnz": int(ns["assembler"].nnz)}
    for key, name in (("drag", "drag_list"), ("lift", "lift_list"), ("dis_x", "displacement_x_list"),
                      ("dis_y", "displacement_y_list")):
        if name in ns:
            out[key] = [float(v) for v in ns[name]]
    return out


def run_oracle_tf_fsi():
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_asm(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10,
                               use_c_port=True)
    ns = monolithic.run("TF_fsi", make_asm, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs),
                        overrides=dict(T=0.05, dt=0.01, theta=0.51, extrapolation="laplace",
                                       extrapolation_sub_type="constant", verbose=False))
    return {
        "config": "TF_fsi --theta 0.51 -dt 0.01 -T 0.05 --extrapolation laplace --extrapolation-sub-type constant",
        "residual_history": [[int(c), int(i), float(r), float(rr)] for c, i, r, rr in ns["residual_history"]],
        "drag": [float(v) for v in ns["drag_list"]], "lift": [float(v) for v in ns["lift_list"]],
        "dis_x": [float(v) for v in ns["displacement_x_list"]], "dis_y": [float(v) for v in ns["displacement_y_list"]],
        "num_dofs": int(ns["DVP"].num_dofs), "nnz": int(ns["assembler"].nnz),
    }


def run_numpy_oracle_elastic():
    """test_elastic (tests/test_turtleFSI.py:106-123): elastic lifting exists only in the numpy oracle
    (oracle/forms.py, turtleFSI/modules/elastic.py:18-27); ~2 minutes."""
    import types
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU
    mod = types.ModuleType("turtlefsi_b200.modules.elastic")

    def extrapolate_setup(F_fluid_linear, extrapolation_sub_type="constant", **ns):
        t = dict(F_fluid_linear)
        t.update(extrapolation="elastic", extrapolation_sub_type=extrapolation_sub_type)
        return dict(F_fluid_linear=t)
    mod.extrapolate_setup = extrapolate_setup
    sys.modules["turtlefsi_b200.modules.elastic"] = mod

    def make_asm(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10)
    ns = monolithic.run("TF_fsi", make_asm, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs),
                        overrides=dict(T=0.05, dt=0.01, theta=0.51, extrapolation="elastic",
                                       extrapolation_sub_type="constant", verbose=False))
    return {"problem": "TF_fsi", "overrides": "-e elastic -et constant --theta 0.51", "oracle": "numpy complex-step",
            "drag": [float(v) for v in ns["drag_list"]], "lift": [float(v) for v in ns["lift_list"]],
            "dis_x": [float(v) for v in ns["displacement_x_list"]], "dis_y": [float(v) for v in ns["displacement_y_list"]]}


def run_numpy_oracle_volume_change():
    """test_laplace[volume_change] (tests/test_turtleFSI.py:61-80): alfa = 1/J(d) (laplace.py:33-34) is outside the
    C port; numpy complex-step oracle, ~2 minutes."""
    from oracle.backend import OracleAssembler
    from turtlefsi_b200 import monolithic
    from turtlefsi_b200.linear_solver import HostLU

    def make_asm(mesh, space, regions, ns):
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"])
        return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10)
    ns = monolithic.run("TF_fsi", make_asm, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs),
                        overrides=dict(T=0.05, dt=0.01, theta=0.51, extrapolation="laplace",
                                       extrapolation_sub_type="volume_change", verbose=False))
    return {"problem": "TF_fsi", "overrides": "-e laplace -et volume_change --theta 0.51", "oracle": "numpy complex-step",
            "residual_history": [[int(c), int(i), float(r), float(rr)] for c, i, r, rr in ns["residual_history"]],
            "drag": [float(v) for v in ns["drag_list"]], "lift": [float(v) for v in ns["lift_list"]],
            "dis_x": [float(v) for v in ns["displacement_x_list"]], "dis_y": [float(v) for v in ns["displacement_y_list"]]}


SLOW = {"tf_fsi_elastic_numpy_oracle": ("--elastic", run_numpy_oracle_elastic),
        "tf_fsi_laplace_volume_change_numpy_oracle": ("--volume-change", run_numpy_oracle_volume_change)}


if __name__ == "__main__":
    if "--only-slow" in sys.argv:            # regenerate just the flagged numpy-oracle runs inside the stored file
        path = os.path.join(HERE, "turek_oracle_histories.json")
        hist = json.load(open(path))
        for key, (flag, fn) in SLOW.items():
            if flag in sys.argv:
                hist[key] = fn()
        json.dump(hist, open(path, "w"), indent=1)
        print("written")
        sys.exit(0)
    json.dump(GOLDENS, open(os.path.join(HERE, "turek_goldens.json"), "w"), indent=1)
    json.dump(run_oracle_tf_fsi(), open(os.path.join(HERE, "tf_fsi_oracle_history.json"), "w"), indent=1)
    hist = {name: run_oracle(prob, ov) for name, (prob, ov, _) in CASES.items()}
    try:
        old = json.load(open(os.path.join(HERE, "turek_oracle_histories.json")))
    except Exception:
        old = {}
    for key, (flag, fn) in SLOW.items():     # keep the stored (slow to regenerate) numpy-oracle runs unless asked
        if flag in sys.argv:
            hist[key] = fn()
        elif key in old:
            hist[key] = old[key]
    json.dump(hist, open(os.path.join(HERE, "turek_oracle_histories.json"), "w"), indent=1)
    print("written")
