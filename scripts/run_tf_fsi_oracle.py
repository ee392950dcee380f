"""run TF_fsi (laplace/constant) 6 steps through the ORACLE backend; print the final scalars."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from oracle.backend import OracleAssembler
from turtlefsi_b200 import monolithic
from turtlefsi_b200.linear_solver import HostLU

def make_asm(mesh, space, regions, ns):
    common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                  extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"],
                  gravity=ns["gravity"])
    return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10)

t0 = time.time()
ns = monolithic.run("TF_fsi", make_asm, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs),
                    overrides=dict(T=0.05, dt=0.01, theta=0.51, extrapolation="laplace", extrapolation_sub_type=sys.argv[1] if len(sys.argv) > 1 else "constant", verbose=True))
print("time %.1fs" % (time.time() - t0))
print("dis_x", ns["displacement_x_list"][-1], "golden -6.896013956339182e-06")
print("dis_y", ns["displacement_y_list"][-1], "golden 1.876355330341896e-09")
print("drag ", ns["drag_list"][-1], "golden 4.407481239804155")
print("lift ", ns["lift_list"][-1], "golden -0.005404703556977697")
