"""short run for ncu: a few Jacobian + residual assemblies on a tube mesh."""
import sys
import torch
sys.path.insert(0, ".")
from turtlefsi_b200.mesh import tube_mesh, synthetic_state
from turtlefsi_b200.space import MixedSpace
from turtlefsi_b200.assembler import FSIAssembler
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
regions = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3),
           dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703)]
mesh = tube_mesh(n); space = MixedSpace(mesh)
asm = FSIAssembler(mesh, space, regions, 1e-3, 0.501, verify_pattern=0)
x = torch.as_tensor(synthetic_state(mesh, space, 2024, h=2.0 / n)).cuda()
x1 = torch.as_tensor(synthetic_state(mesh, space, 2025, h=2.0 / n)).cuda()
A_pre = asm.assemble_A_pre(x, x1); A = asm.new_matrix(); b = asm.new_vector()
for _ in range(reps):
    asm.assemble_jacobian(x, x1, A_pre, A)
    asm.assemble_residual(x, x1, b)
torch.cuda.synchronize()
print("done", mesh.num_cells)
