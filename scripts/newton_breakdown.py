"""Newton-step wall-time breakdown of the 2-D BASELINE configs on the GPU (SURVEY 8d metric (ii)):
assemble-J / assemble-R / BC + ident_zeros / linear solve (host SuperLU, not optimised, reported separately) /
update + norms, averaged over the Newton iterations of a few time steps. Development aid; writes one JSON line
per problem."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from turtlefsi_b200 import monolithic                    # noqa: E402
from turtlefsi_b200.assembler import FSIAssembler        # noqa: E402
from turtlefsi_b200.linear_solver import HostLU          # noqa: E402


class Timed:
    """wraps an object and accumulates device-synchronised wall time per method group."""

    def __init__(self, obj, groups, acc):
        self._o, self._g, self._acc = obj, groups, acc

    def __getattr__(self, name):
        attr = getattr(self._o, name)
        grp = self._g.get(name)
        if grp is None or not callable(attr):
            return attr

        def call(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = attr(*a, **k)
            torch.cuda.synchronize()
            self._acc[grp] = self._acc.get(grp, 0.0) + time.perf_counter() - t0
            self._acc["n_" + grp] = self._acc.get("n_" + grp, 0) + 1
            return r
        return call


GROUPS = dict(assemble_jacobian="assemble_J", assemble_residual="assemble_R", ident_zeros="bc_ident",
              apply_bc_matrix="bc_ident", apply_bc_rhs="bc_ident", apply_bc_vec="update_norms", axpy="update_norms",
              norm="update_norms", copy="update_norms")


def run(problem, steps, **over):
    acc = {}

    def make_asm(mesh, space, regions, ns):
        a = FSIAssembler(mesh, space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"], solid=ns["solid"],
                         extrapolation=ns["extrapolation"], extrapolation_sub_type=ns["extrapolation_sub_type"],
                         quad_degree_jacobian=ns["compiler_parameters"]["quadrature_degree"], quad_degree_residual=10,
                         gravity=ns["gravity"])
        return Timed(a, GROUPS, acc)

    def make_solver(a):
        lu = HostLU(a.rowptr.cpu().numpy(), a.colidx.cpu().numpy(), a.num_dofs)
        return Timed(lu, dict(solve="linear_solve", set_operator="linear_solve"), acc)
    ov = dict(verbose=False)
    ov.update(over)
    t0 = time.perf_counter()
    ns = monolithic.run(problem, make_asm, make_solver, overrides=ov, max_steps=steps)
    wall = time.perf_counter() - t0
    its = len(ns["residual_history"])
    out = dict(problem=problem, cells=int(ns["mesh"].num_cells), dofs=int(ns["DVP"].num_dofs), nnz=int(ns["assembler"].nnz),
               time_steps=steps, newton_iterations=its, jacobian_assemblies=acc.get("n_assemble_J", 0),
               ms_per_call=dict(assemble_J=1e3 * acc.get("assemble_J", 0) / max(1, acc.get("n_assemble_J", 0)),
                                assemble_R=1e3 * acc.get("assemble_R", 0) / max(1, acc.get("n_assemble_R", 0))),
               ms_per_newton_iteration={k: 1e3 * acc.get(k, 0.0) / its for k in
                                        ("assemble_J", "assemble_R", "bc_ident", "linear_solve", "update_norms")},
               wall_s=wall)
    print(json.dumps(out))


if __name__ == "__main__":
    run("TF_fsi", 3, T=1, dt=0.01, theta=0.51)
    run("TF_cfd", 3, T=1, dt=0.01, theta=0.5)
    run("turtle_demo", 2)
