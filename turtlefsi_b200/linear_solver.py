"""Linear solve of the Newton step — reported separately, NOT optimised (BASELINE north_star).

The reference uses PETSc LU through MUMPS and re-uses the factorisation while A is not
re-assembled (turtleFSI/modules/newtonsolver.py:51,93; docs/source/verif_perf.rst:24-35). The B200
stand-in named by the north_star is cuDSS, which is NOT in this image; this module
keeps the same set_operator/solve protocol with SciPy SuperLU on the host for problems of up to a
few 1e5 dofs (the 2-D benchmarks). It is outside the hot path of this repository."""
import numpy as np


class HostLU:
    def __init__(self, rowptr, colidx, n):
        self.rowptr = np.asarray(rowptr)
        self.colidx = np.asarray(colidx)
        self.n = n
        self._vals = None
        self._lu = None
        self.factorizations = 0

    def set_operator(self, A_vals):
        """A changed (or may have): the next solve re-factorises (PETSc operator-state tracking)."""
        self._vals = A_vals
        self._lu = None

    def solve(self, b):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        is_torch = hasattr(b, "cpu")
        if self._lu is None:
            vals = self._vals.cpu().numpy() if hasattr(self._vals, "cpu") else np.asarray(self._vals)
            A = sp.csr_matrix((vals, self.colidx, self.rowptr), shape=(self.n, self.n)).tocsc()
            self._lu = spla.splu(A, permc_spec="MMD_ATA")      # ~2x faster than COLAMD on these matrices
            self.factorizations += 1
        bh = b.cpu().numpy() if is_torch else np.asarray(b)
        xh = self._lu.solve(bh)
        if is_torch:
            import torch
            return torch.as_tensor(xh).to(b.device)
        return xh
