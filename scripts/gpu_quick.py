"""quick on-GPU timing of the Jacobian/residual kernels (development aid, not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from turtlefsi_b200.mesh import tube_mesh, synthetic_state
from turtlefsi_b200.space import MixedSpace
from turtlefsi_b200.assembler import FSIAssembler, probe_fp64_peak

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
print("fp64 probe TFLOP/s, ms:", probe_fp64_peak())
regions = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3),
           dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703)]
t = time.time(); mesh = tube_mesh(n); space = MixedSpace(mesh); pat = space.pattern()
print("mesh", mesh.num_cells, "fluid", int((mesh.markers == 1).sum()), "dofs", space.num_dofs, "nnz", len(pat[1]), "host setup %.1fs" % (time.time() - t))
t = time.time()
asm = FSIAssembler(mesh, space, regions, 1e-3, 0.501, pattern=pat, verify_pattern=1)
torch.cuda.synchronize(); print("create %.2fs" % (time.time() - t))
x = torch.as_tensor(synthetic_state(mesh, space, 2024, h=2.0 / n)).cuda()
x1 = torch.as_tensor(synthetic_state(mesh, space, 2025, h=2.0 / n)).cuda()
A_pre = asm.assemble_A_pre(x, x1); A = asm.new_matrix(); b = asm.new_vector()
def timeit(f, k=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k
ms = timeit(lambda: asm.assemble_jacobian(x, x1, A_pre, A)); print("jacobian+A_pre ms %.3f  cells/s %.3e" % (ms, mesh.num_cells / ms * 1e3))
ms = timeit(lambda: asm.assemble_jacobian(x, x1, None, A)); print("jacobian (memset) ms %.3f  cells/s %.3e" % (ms, mesh.num_cells / ms * 1e3))
ms = timeit(lambda: asm.assemble_residual(x, x1, b)); print("residual ms %.3f  cells/s %.3e" % (ms, mesh.num_cells / ms * 1e3))
ms = timeit(lambda: asm.ident_zeros(A)); print("ident_zeros ms %.3f" % ms)
ms = timeit(lambda: A.copy_(A_pre)); print("copy ms %.3f GB/s %.1f" % (ms, 2 * A.numel() * 8 / ms / 1e6))
