"""development aid: DeviceRfLU (cusolverRf) against HostLU on the TF_fsi Newton matrices."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from turtlefsi_b200 import monolithic
from turtlefsi_b200.assembler import FSIAssembler
from turtlefsi_b200.linear_solver import DeviceRfLU, HostLU

def make_asm(mesh, space, regions, ns):
    return FSIAssembler(mesh, space, regions, ns["dt"], ns["theta"], fluid=ns["fluid"], solid=ns["solid"],
                        extrapolation=ns["extrapolation"], extrapolation_sub_type=ns["extrapolation_sub_type"],
                        quad_degree_jacobian=ns["compiler_parameters"]["quadrature_degree"], quad_degree_residual=10,
                        gravity=ns["gravity"])
made = {}
def make_solver(a):
    made["lu"] = DeviceRfLU(a)
    return made["lu"]
t0 = time.time()
ns = monolithic.run(sys.argv[1] if len(sys.argv) > 1 else "TF_fsi", make_asm, make_solver, overrides=dict(T=0.05, dt=0.01, theta=0.51, verbose=False))
lu = made["lu"]
print("wall", time.time() - t0, "iters", len(ns["residual_history"]), "host analyses", lu.host_analyses, "refactor", lu.refactorizations,
      "solves", lu.solves, "refinements", lu.refinements, "factor nnz", lu.factor_nnz, "last rel res", lu.last_relative_residual)
print("seconds", lu.seconds, "per refactor ms", 1e3 * lu.seconds["refactor"] / max(1, lu.refactorizations), "per solve ms", 1e3 * lu.seconds["solve"] / max(1, lu.solves))
if "drag_list" in ns: print("drag", ns["drag_list"][-1], "lift", ns["lift_list"][-1])
print([ (r[0], r[1], "%.3e" % r[2], "%.3e" % r[3]) for r in ns["residual_history"]][:8])
