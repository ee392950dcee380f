"""TF_fsi with the file-default biharmonic lifting through the numpy ORACLE (slow, one-off)."""
import sys, time, json
sys.path.insert(0, ".")
from oracle.backend import OracleAssembler
from turtlefsi_b200 import monolithic
from turtlefsi_b200.linear_solver import HostLU
sub = sys.argv[1] if len(sys.argv) > 1 else "constrained_disp_vel"
def make_asm(mesh, space, regions, ns):
    common = dict(dt=ns["dt"], theta=ns["theta"], extrapolation=ns["extrapolation"],
                  extrapolation_sub_type=ns["extrapolation_sub_type"], fluid=ns["fluid"], solid=ns["solid"], gravity=ns["gravity"])
    return OracleAssembler(mesh, space, regions, common, ns["compiler_parameters"]["quadrature_degree"], 10)
t0 = time.time()
ns = monolithic.run("TF_fsi", make_asm, lambda a: HostLU(a.rowptr, a.colidx, a.num_dofs),
                    overrides=dict(T=0.05, dt=0.01, theta=0.51, extrapolation="biharmonic", extrapolation_sub_type=sub, verbose=False))
out = dict(sub=sub, time=time.time() - t0, dis_x=ns["displacement_x_list"], dis_y=ns["displacement_y_list"], drag=ns["drag_list"], lift=ns["lift_list"],
           residual_history=[[int(c), int(i), float(r), float(rr)] for c, i, r, rr in ns["residual_history"]], num_dofs=int(ns["DVP"].num_dofs), nnz=int(ns["assembler"].nnz))
json.dump(out, open("/tmp/tf_fsi_bih_%s.json" % sub, "w"))
print(sub, "time %.0fs" % out["time"], "dis_x", out["dis_x"][-1], "dis_y", out["dis_y"][-1], "drag", out["drag"][-1], "lift", out["lift"][-1], out["num_dofs"], out["nnz"])
