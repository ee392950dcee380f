"""The reference's own MPI test, mirrored: tests/test_turtleFSI.py:40-58 runs TF_fsi with `mpirun -np 1` and
`-np 2` against the same goldens. Here the problem file is run SPMD over 2 ranks (turtlefsi_b200.distributed:
recursive-bisection partition, ghost refresh before every assembly, owned rows only, gathered host LU):
  * CPU, gloo, oracle backend (3 steps) against the stored single-domain oracle history;
  * 2 GPUs, NCCL, CUDA path (6 steps) against the reference's goldens (skipped on a single-GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from turtlefsi_b200 import monolithic
from turtlefsi_b200.distributed import GatheredLU, make_distributed
gpu = %(gpu)r
if gpu:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl" if gpu else "gloo")
rank = dist.get_rank()

def make_asm(mesh, space, regions, ns):
    degJ = ns["compiler_parameters"]["quadrature_degree"]
    if gpu:
        from turtlefsi_b200.assembler import FSIAssembler
        mk = lambda view: FSIAssembler(view.mesh, view.space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"],
                                       solid=ns["solid"], extrapolation=ns["extrapolation"],
                                       extrapolation_sub_type=ns["extrapolation_sub_type"], quad_degree_jacobian=degJ,
                                       quad_degree_residual=10, gravity=ns["gravity"], hmin=mesh.hmin())
    else:
        from oracle.backend import OracleAssembler
        common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                      extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                      gravity=ns["gravity"], hmin=mesh.hmin())
        mk = lambda view: OracleAssembler(view.mesh, view.space, regions, common, degJ, 10, use_c_port=True)
    return make_distributed(mesh, space, mk, biharmonic=space.biharmonic, use_gpu_halo=gpu)

ns = monolithic.run("TF_fsi", make_asm, GatheredLU, overrides=dict(T=%(T)r, dt=0.01, theta=0.51, verbose=False,
                                                                   visualization_folder=%(viz)r, save_step=1))
x_global = np.asarray(ns["assembler"].to_host(ns["dvp_"]["n"]))          # collective
if rank == 0:
    from turtlefsi_b200.utils.xdmf import read_series
    coords, cells, steps = read_series(os.path.join(%(viz)r, "velocity.xdmf"))
    sp = ns["DVP"]
    viz_ok = (len(steps) == len(ns["drag_list"]) and cells.shape[0] == ns["mesh"].num_cells
              and all(np.array_equal(steps[-1][2][:, a], x_global[sp.node_dof(1, a)[sp.vertex_node]]) for a in range(2)))
    view = ns["assembler"].view
    out = dict(drag=ns["drag_list"], lift=ns["lift_list"], dis_x=ns["displacement_x_list"], dis_y=ns["displacement_y_list"],
               iters=[list(r[:2]) for r in ns["residual_history"]], residuals=[r[2] for r in ns["residual_history"]],
               owned_cells=int(view.cell_owned.sum()), local_cells=int(view.mesh.num_cells),
               factorizations=ns["up_sol"].factorizations, viz_ok=bool(viz_ok))
    json.dump(out, open(%(out)r, "w"))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _run(tmp_path, gpu, T, port):
    out = tmp_path / "result.json"
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, gpu=gpu, T=T, out=str(out), viz=str(tmp_path / "viz")))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=1500)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == 2
    return json.load(open(out))


def test_tf_fsi_two_ranks_oracle_gloo(tmp_path):
    r = _run(tmp_path, False, 0.02, 29641)
    h = json.load(open(os.path.join(ROOT, "tests", "golden", "turek_oracle_histories.json")))["tf_fsi_biharmonic_default"]
    n = len(r["drag"])
    assert n == 3 and 0 < r["owned_cells"] < r["local_cells"] < 2474            # really partitioned, with ghosts
    assert r["viz_ok"]            # the XDMF series rank 0 wrote holds the GATHERED global fields of every step
    assert np.allclose(r["drag"], h["drag"][:n], rtol=1e-11, atol=0)
    assert np.allclose(r["dis_x"], h["dis_x"][:n], rtol=1e-9, atol=0)
    assert np.allclose(r["lift"], h["lift"][:n], rtol=1e-8, atol=0)
    ref_it = [tuple(x[:2]) for x in h["residual_history"]]
    assert [tuple(x) for x in r["iters"]] == ref_it[:len(r["iters"])]             # same Newton iteration counts
    for got, ref in zip(r["residuals"], h["residual_history"]):
        if ref[2] > 1e-3:
            assert abs(got / ref[2] - 1) < 1e-8                                  # distributed norm == global norm


@pytest.mark.gpu
def test_tf_fsi_two_gpus_nccl_hits_reference_goldens(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run(tmp_path, True, 0.05, 29643)
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "turek_goldens.json")))["TF_fsi"]
    assert len(r["drag"]) == 6
    assert np.isclose(r["dis_x"][-1], g["dis_x"]) and np.isclose(r["dis_y"][-1], g["dis_y"])      # the reference's assertions
    assert np.isclose(r["drag"][-1], g["drag"]) and np.isclose(r["lift"][-1], g["lift"])
    assert abs(r["drag"][-1] / g["drag"] - 1) < 1e-10
