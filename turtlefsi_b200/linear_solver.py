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


def _cusolver_path():
    """the cuSOLVER that matches the cuBLAS already loaded by torch (its wheel bundles both); the toolkit's
    own copy only when torch does not ship one."""
    import glob
    import os

    import torch
    base = os.path.dirname(os.path.dirname(torch.__file__))
    hits = sorted(glob.glob(os.path.join(base, "nvidia", "cusolver", "lib", "libcusolver.so*")))
    return hits[-1] if hits else "libcusolver.so"


class DeviceRfLU:
    """Device-side LU with the reference's re-use protocol (newtonsolver.py:51,93; docs/source/verif_perf.rst:24-35),
    on cuSOLVER's re-factorisation module (cusolverRf, in the CUDA toolkit; cuDSS is not in this image):

      analysis + pivoting   ONCE, on the host (SciPy SuperLU): permutations P, Q and the patterns of L and U;
      re-factorisation      on the DEVICE whenever A was re-assembled (set_operator): same pattern, same pivots;
      solve                 on the DEVICE; nothing is re-factorised while A is re-used.

    Static pivoting is checked, not trusted: every solve is followed by a residual check r = b - A x
    (tfsi_csr_matvec) with iterative refinement; if that does not reach `tol`, pivoting is redone on the host
    for the current matrix. The linear solve is outside the hot path of this repository (BASELINE north_star:
    reported separately, not optimised); this class removes the host round trip of A from the Newton loop."""

    def __init__(self, asm, tol=1e-11, max_refine=3):
        import ctypes as C
        self.asm, self.tol, self.max_refine = asm, tol, max_refine
        self.n = asm.num_dofs
        self.lib = C.CDLL(_cusolver_path())
        self._C = C
        self.handle = None
        self._vals, self._dirty = None, False
        self.host_analyses = self.refactorizations = self.solves = self.refinements = 0
        self.seconds = dict(host_analysis=0.0, refactor=0.0, solve=0.0)     # device-synchronised wall time
        import torch
        self.rowptr32 = asm.rowptr.to(torch.int32)
        self.colidx32 = asm.colidx.to(torch.int32).contiguous()

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError("cusolverRf %s failed with status %d" % (what, rc))

    def set_operator(self, A_vals):
        self._vals, self._dirty = A_vals, True
        self._anorm = None

    def _norm_inf(self):
        """max row sum of |A| (for the normwise backward error of a solve)."""
        if self._anorm is None:
            import torch
            if getattr(self, "_row_of", None) is None:
                counts = (self.rowptr32[1:] - self.rowptr32[:-1]).long()
                self._row_of = torch.repeat_interleave(torch.arange(self.n, device=counts.device), counts)
            rs = torch.zeros(self.n, dtype=torch.float64, device=self._vals.device)
            rs.index_add_(0, self._row_of, self._vals.abs())
            self._anorm = float(rs.max())
        return self._anorm

    def _host_analysis(self):
        import time

        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        import torch
        C = self._C
        t0 = time.perf_counter()
        n = self.n
        rp, ci = self.rowptr32.cpu().numpy(), self.colidx32.cpu().numpy()
        vals = self._vals.cpu().numpy()
        A = sp.csr_matrix((vals, ci, rp), shape=(n, n))
        lu = spla.splu(A.tocsc(), permc_spec="MMD_ATA")     # pivoting + fill-reducing ordering on the real values
        P = np.argsort(lu.perm_r).astype(np.int32)          # (L U)[i, j] = A[P[i], Q[j]]
        Q = np.argsort(lu.perm_c).astype(np.int32)
        # SYMBOLIC pattern of L and U for that pivot order. SuperLU's numeric factors drop the fill that is
        # numerically zero (81 % of the pattern of A holds explicit zeros), but a re-factorisation with other
        # values needs all of it: factorise a generic diagonally dominant matrix with the SAME pattern and
        # order, without pivoting — no cancellation, hence the complete structural fill.
        pat = sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(n, n))[P][:, Q].tocsr()
        gen = pat.copy()
        gen.data = np.random.default_rng(1).uniform(0.5, 1.5, size=gen.nnz)
        gen = gen.tolil()
        gen.setdiag(10.0 * np.asarray(pat.sum(axis=1)).ravel())
        sym = spla.splu(gen.tocsc(), permc_spec="NATURAL", diag_pivot_thresh=0.0, options=dict(Equil=False))
        ident = np.arange(n)
        if not (np.array_equal(sym.perm_r, ident) and np.array_equal(sym.perm_c, ident)):
            raise RuntimeError("symbolic factorisation permuted the matrix")
        L, U = sym.L.tocsr(), sym.U.tocsr()
        L.sort_indices()
        U.sort_indices()
        if self.handle is not None:
            self.lib.cusolverRfDestroy(self.handle)
        self.handle = C.c_void_p()
        self._chk(self.lib.cusolverRfCreate(C.byref(self.handle)), "Create")
        self.lib.cusolverRfSetNumericProperties.argtypes = [C.c_void_p, C.c_double, C.c_double]
        self._chk(self.lib.cusolverRfSetNumericProperties(self.handle, 0.0, 0.0), "SetNumericProperties")
        self._chk(self.lib.cusolverRfSetResetValuesFastMode(self.handle, 1), "SetResetValuesFastMode")
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        keep = [i32(rp), i32(ci), f64(vals), i32(L.indptr), i32(L.indices), f64(L.data),
                i32(U.indptr), i32(U.indices), f64(U.data), P, Q]
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        self.lib.cusolverRfSetupHost.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + \
            [C.c_int] + [C.c_void_p] * 3 + [C.c_void_p] * 2 + [C.c_void_p]
        self._chk(self.lib.cusolverRfSetupHost(n, len(ci), ptr(keep[0]), ptr(keep[1]), ptr(keep[2]),
                                               L.nnz, ptr(keep[3]), ptr(keep[4]), ptr(keep[5]),
                                               U.nnz, ptr(keep[6]), ptr(keep[7]), ptr(keep[8]),
                                               ptr(P), ptr(Q), self.handle), "SetupHost")
        torch.cuda.synchronize()
        self.lib.cusolverRfAnalyze.argtypes = [C.c_void_p]
        self._chk(self.lib.cusolverRfAnalyze(self.handle), "Analyze")
        dev = self._vals.device
        self.dP, self.dQ = torch.as_tensor(P).to(dev), torch.as_tensor(Q).to(dev)
        self.temp = torch.empty(n, dtype=torch.float64, device=dev)
        self.factor_nnz = int(L.nnz + U.nnz)
        self.host_analyses += 1
        self.seconds["host_analysis"] += time.perf_counter() - t0

    def _refactor(self):
        import time

        import torch
        C = self._C
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p = lambda t: C.c_void_p(t.data_ptr())
        self.lib.cusolverRfResetValues.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6
        self._chk(self.lib.cusolverRfResetValues(self.n, self.colidx32.numel(), p(self.rowptr32), p(self.colidx32),
                                                 p(self._vals), p(self.dP), p(self.dQ), self.handle), "ResetValues")
        self.lib.cusolverRfRefactor.argtypes = [C.c_void_p]
        self._chk(self.lib.cusolverRfRefactor(self.handle), "Refactor")
        torch.cuda.synchronize()
        self.refactorizations += 1
        self.seconds["refactor"] += time.perf_counter() - t0

    def _rf_solve(self, rhs):
        C = self._C
        x = rhs.clone()
        p = lambda t: C.c_void_p(t.data_ptr())
        self.lib.cusolverRfSolve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self._chk(self.lib.cusolverRfSolve(self.handle, p(self.dP), p(self.dQ), 1, p(self.temp), self.n, p(x), self.n), "Solve")
        return x

    def solve(self, b):
        import torch
        if self._dirty:
            torch.cuda.synchronize()
            if self.handle is None:
                self._host_analysis()
            self._refactor()
            self._dirty = False
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for attempt in range(2):
            x = self._rf_solve(b)
            bn = float(torch.linalg.vector_norm(b))
            ok = False
            for it in range(self.max_refine + 1):
                r = b - self.asm.matvec(self._vals, x)
                rn = float(torch.linalg.vector_norm(r))
                if not np.isfinite(rn):
                    break
                # accepted: residual below tol |b|, or — when b itself is at round-off level of A x, as in the last
                # Newton iterations — a normwise backward error |r| / (|A| |x| + |b|) of a few ulps
                xn = float(torch.linalg.vector_norm(x, ord=float("inf")))
                if rn <= self.tol * bn or bn == 0.0 or rn <= 64 * 2.2e-16 * (self._norm_inf() * xn * np.sqrt(self.n) + bn):
                    ok = True
                    break
                if it < self.max_refine:
                    x = x + self._rf_solve(r)
                    self.refinements += 1
            if ok:
                break
            if attempt == 1:
                # MUMPS in the reference either solves or raises; never hand Newton an increment that failed its check
                raise RuntimeError("DeviceRfLU: linear solve did not reach the residual tolerance after re-pivoting "
                                   "(|b - A x| = %.3e, |b| = %.3e, tol = %.1e)" % (rn, bn, self.tol))
            t1 = time.perf_counter()
            self._host_analysis()                 # static pivots no longer adequate: new pivoting for this matrix
            self._refactor()
            t0 += time.perf_counter() - t1        # counted under host_analysis / refactor
        torch.cuda.synchronize()
        self.seconds["solve"] += time.perf_counter() - t0
        self.solves += 1
        self.last_relative_residual = rn / bn if bn else 0.0
        return x


def probe_cudss():
    """runtime probe for cuDSS, the sparse direct solver BASELINE's north_star names (`dlopen`, no link-time
    dependency). Returns (ctypes library | None, text). cuDSS is not part of the CUDA toolkit and is NOT in this
    image; where a deployment has it (LD_LIBRARY_PATH, the nvidia-cudss wheel), the probe finds it and reports its
    version."""
    import ctypes as C
    import glob
    import os
    import sys
    names = ["libcudss.so.0", "libcudss.so"]
    for base in sys.path:
        names += sorted(glob.glob(os.path.join(base, "nvidia", "cu*", "lib", "libcudss.so*")))
        names += sorted(glob.glob(os.path.join(base, "nvidia", "cudss", "lib", "libcudss.so*")))
    for name in names:
        try:
            lib = C.CDLL(name)
        except OSError:
            continue
        ver = []
        try:
            for prop in (0, 1, 2):                       # MAJOR_VERSION, MINOR_VERSION, PATCH_LEVEL (libraryPropertyType)
                v = C.c_int(0)
                if lib.cudssGetProperty(prop, C.byref(v)) != 0:
                    raise OSError("cudssGetProperty failed")
                ver.append(str(v.value))
        except Exception:
            ver = ["?"]
        return lib, "cuDSS %s found (%s)" % (".".join(ver), name)
    return None, "cuDSS not found by dlopen (libcudss.so[.0], nvidia-cudss wheel paths): not in this image"


def make_device_solver(asm, tol=1e-11):
    """(solver, name, note): the device LU behind the reference's set_operator / solve re-use protocol
    (newtonsolver.py:51,93). cuDSS is probed at run time; the path that exists and is tested in this image is
    DeviceRfLU (cusolverRf re-factorisation on the device after a one-off host analysis), so it is what is returned —
    the probe result is reported so that a deployment with cuDSS sees it named."""
    _, text = probe_cudss()
    return DeviceRfLU(asm, tol=tol), "cusolverRf (DeviceRfLU)", text + "; fallback in use: host analysis once (SciPy SuperLU), " \
        "numeric re-factorisation and solves on the device (cusolverRf), every solve residual-checked"
