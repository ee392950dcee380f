"""Python host side of the hot path: owns the C-ABI handle and hands torch device buffers to it.

Replaces, for turtleFSI's Newton solver (turtleFSI/modules/newtonsolver.py):
  assemble(J_linear, keep_diagonal=True)                     :26-27  -> assemble_A_pre
  assemble(J_nonlinear, tensor=A) ; A.axpy(1.0, A_pre, True) :75-79  -> assemble_jacobian
  A.ident_zeros()                                            :81     -> ident_zeros
  bc.apply(A) / bc.apply(b, x) / bc.apply(x)                 :82,91,98
  assemble(-F, tensor=b)                                     :88     -> assemble_residual
  b.norm('l2'), norm(dvp_res, 'l2')                          :105-106 -> norm
PyTorch is plumbing only (device memory, streams, torch.distributed)."""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .quadrature import default_rule

EXT_CODES = {"no_extrapolation": 0, "laplace": 1, "biharmonic": 2, "elastic": 3}
LAPLACE_SUB = {"constant": 0, "small_constant": 1, "volume": 2, "volume_change": 3}


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


class FSIAssembler:
    """One handle = one mesh partition on one GPU.

    regions: list of dicts {marker, kind ("fluid"|"solid"), rho, mu, lambda_} — what
    turtleFSI/modules/domain.py:54-103 derives from dx_f_id / dx_s_id and the property lists; solid regions may add
    the reference's material_model / C01 / C10 / C11 / Jm / viscoelasticity / mu_visc_s / lambda_visc_s entries.
    """

    def __init__(self, mesh, space, regions, dt, theta, fluid="fluid", solid="solid",
                 extrapolation="laplace", extrapolation_sub_type="constant",
                 quad_degree_jacobian=4, quad_degree_residual=None, gravity=None,
                 pattern=None, device=None, verify_pattern=1, hmin=None, deterministic=False):
        if not torch.cuda.is_available():
            raise RuntimeError("turtlefsi_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.mesh, self.space = mesh, space
        g = mesh.gdim
        if extrapolation not in EXT_CODES:
            raise RuntimeError("extrapolation '%s' is not available on the B200 path" % extrapolation)
        if extrapolation == "laplace" and extrapolation_sub_type not in LAPLACE_SUB:
            raise RuntimeError("Could not find extrapolation method {}".format(extrapolation_sub_type))
        if quad_degree_residual is None:
            # assemble(-F) uses the global form-compiler degree; "auto" estimates
            # 10 in 2-D; every 3-D problem file sets 6 (problem_aneu.py:12)
            quad_degree_residual = 10 if g == 2 else 6
        self.qJ = default_rule(g, quad_degree_jacobian)
        self.qR = default_rule(g, quad_degree_residual)
        if pattern is None:
            from .space import pattern_device
            pattern = pattern_device(space, self.device) if space.numbering == "interleaved" else space.pattern()
        rowptr, colidx = pattern
        dev = self.device

        def to_dev(a, dtype):
            if isinstance(a, torch.Tensor):
                return a.to(device=dev, dtype=dtype).contiguous()
            return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev)

        self.rowptr = to_dev(rowptr, torch.int64)
        self.colidx = to_dev(colidx, torch.int32)
        self.num_dofs = int(space.num_dofs)
        self.nnz = int(self.rowptr[-1].item())
        coords = np.ascontiguousarray(mesh.coords, dtype=np.float64)
        cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
        cdofs = np.ascontiguousarray(space.cell_dofs, dtype=np.int32)
        marker = np.ascontiguousarray(mesh.markers, dtype=np.int32)
        regs = (_lib.TfsiRegion * len(regions))()
        for i, r in enumerate(regions):
            regs[i].marker = int(r["marker"])
            regs[i].kind = 0 if r["kind"] == "fluid" else 1
            regs[i].rho = float(r["rho"])
            regs[i].mu = float(r["mu"])
            regs[i].lam = float(r.get("lambda_", 0.0))
            if r["kind"] == "solid":         # common.py:104-139, solid.py:46-59 (keys of the reference's solid_properties)
                model, visco = r.get("material_model", "StVenantKirchoff"), r.get("viscoelasticity")
                if model not in _lib.MATERIAL_MODELS:
                    raise RuntimeError('Invalid entry for material_model, choose from ["StVenantKirchoff",'
                                       '"StVenantKirchoffEnergy","NeoHookean","MooneyRivlin","Gent"]')
                regs[i].material = _lib.MATERIAL_MODELS[model]
                regs[i].viscoelasticity = _lib.VISCOELASTICITY.get(visco, 0)     # invalid entry: none, as solid.py:56-58
                for key in ("C01", "C10", "C11", "Jm"):
                    setattr(regs[i], key, float(r.get(key, 0.0)))
                regs[i].mu_visc = float(r.get("mu_visc_s", 0.0))
                regs[i].lambda_visc = float(r.get("lambda_visc_s", 0.0))
        p = _lib.TfsiProblem()
        p.gdim, p.biharmonic = g, int(space.biharmonic)
        p.num_vertices, p.num_cells, p.num_dofs = mesh.num_vertices, mesh.num_cells, self.num_dofs
        p.coords, p.cell_vertices, p.cell_dofs = coords.ctypes.data, cells.ctypes.data, cdofs.ctypes.data
        p.cell_marker = marker.ctypes.data
        p.rowptr, p.colidx = self.rowptr.data_ptr(), self.colidx.data_ptr()
        p.num_regions, p.regions = len(regions), regs
        p.dt, p.theta = float(dt), float(theta)
        p.fluid_on, p.solid_on = int(fluid == "fluid"), int(solid == "solid")
        p.extrapolation = EXT_CODES[extrapolation]
        p.extrapolation_sub_type = LAPLACE_SUB.get(extrapolation_sub_type, 0) if extrapolation == "laplace" else 0
        p.hmin = float(hmin if hmin is not None else mesh.hmin())
        p.gravity, p.has_gravity = (float(gravity), 1) if gravity is not None else (0.0, 0)
        p.nq_jacobian, p.qpts_jacobian, p.qwts_jacobian = len(self.qJ[1]), self.qJ[0].ctypes.data, self.qJ[1].ctypes.data
        p.nq_residual, p.qpts_residual, p.qwts_residual = len(self.qR[1]), self.qR[0].ctypes.data, self.qR[1].ctypes.data
        p.verify_pattern = int(verify_pattern)
        p.borrow_pattern = 1          # self.rowptr / self.colidx live as long as the handle: no second copy in HBM
        p.deterministic = int(bool(deterministic))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.tfsi_create(C.byref(p), C.byref(h))
        if rc != 0:
            raise RuntimeError("tfsi_create failed (%d): %s" % (rc, self.lib.tfsi_last_error(None).decode()))
        self.h = h
        self.bc_dofs = None
        self.num_colours = int(self.lib.tfsi_num_colours(self.h))

    # ------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("turtlefsi_b200 (%d): %s" % (rc, self.lib.tfsi_last_error(self.h).decode()))

    def _apre(self, A_pre):
        """A_pre argument of the C call: a tensor, None, or the string "compressed" (= TFSI_APRE_COMPRESSED: the copy
        registered by compress_A_pre, whatever became of the dense array)."""
        if isinstance(A_pre, str):
            if A_pre != "compressed":
                raise ValueError(A_pre)
            return C.c_void_p(_lib.APRE_COMPRESSED)
        return _ptr(A_pre)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.tfsi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def new_vector(self):
        return torch.zeros(self.num_dofs, dtype=torch.float64, device=self.device)

    def new_matrix(self):
        return torch.zeros(self.nnz, dtype=torch.float64, device=self.device)

    # vector plumbing used by modules/newtonsolver.py (torch = device memory only)
    def to_device(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def to_host(self, t):
        return t.detach().cpu().numpy()

    def copy(self, dst, src):
        dst.copy_(src)

    def axpy(self, x, alpha, y):
        x.add_(y, alpha=float(alpha))

    def set_profiling(self, on=True):
        self._check(self.lib.tfsi_set_profiling(self.h, int(bool(on))))

    def kernel_time(self, kind):
        """(total ms, launches) of kernel kind 0 = Jacobian/fluid, 1 = Jacobian/solid, 2 = residual."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        self._check(self.lib.tfsi_kernel_time(self.h, int(kind), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self):
        return int(self.lib.tfsi_launch_count(self.h))

    # ------------------------------------------------------------------------------------
    def assemble_A_pre(self, x, x1, out=None):
        out = self.new_matrix() if out is None else out
        self._check(self.lib.tfsi_assemble_A_pre(self.h, _ptr(x), _ptr(x1), _ptr(out), self._stream()))
        return out

    def set_facet_operator(self, rows, cols, vals):
        """fixed linear facet terms as a COO operator (turtlefsi_b200.boundary.biharmonic_facet_operator)."""
        r = np.ascontiguousarray(rows, dtype=np.int32)
        c = np.ascontiguousarray(cols, dtype=np.int32)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self._check(self.lib.tfsi_set_facet_operator(self.h, _ptr(r), _ptr(c), _ptr(v), len(v)))

    def set_facet_operator_prev(self, rows, cols, vals):
        """linear facet terms acting on dvp_["n-1"] (Robin, solid.py:84): residual only."""
        r = np.ascontiguousarray(rows, dtype=np.int32)
        c = np.ascontiguousarray(cols, dtype=np.int32)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self._check(self.lib.tfsi_set_facet_operator_prev(self.h, _ptr(r), _ptr(c), _ptr(v), len(v)))

    def set_facet_load(self, dofs, vals, scale=1.0):
        """state-independent facet load l (problem_aneu.py:212-214): b -= scale * l."""
        d = np.ascontiguousarray(dofs, dtype=np.int32)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        self._check(self.lib.tfsi_set_facet_load(self.h, _ptr(d), _ptr(v), len(v)))
        self.set_facet_load_scale(scale)

    def set_facet_load_scale(self, scale):
        self._check(self.lib.tfsi_set_facet_load_scale(self.h, float(scale)))

    def matvec(self, A, x, out=None):
        """y = A x on the CSR pattern (residual check around the linear solve)."""
        out = self.new_vector() if out is None else out
        self._check(self.lib.tfsi_csr_matvec(self.h, _ptr(A), _ptr(x), _ptr(out), self._stream()))
        return out

    def compress_A_pre(self, A_pre):
        """register the non-zeros of A_pre so that A := A_pre costs a zero-fill + sparse overwrite."""
        self._check(self.lib.tfsi_compress_A_pre(self.h, _ptr(A_pre), self._stream()))
        return int(self.lib.tfsi_compressed_nnz(self.h))

    def invalidate_A_pre(self):
        """forget the compressed copy (A_pre was edited): the dense array is read again."""
        self._check(self.lib.tfsi_invalidate_A_pre(self.h))

    def bind_matrix(self, A):
        """A = Matrix(A_pre) (newtonsolver.py:28): A becomes the persistent solver matrix; later assemblies into it
        rewrite only the rows the non-linear forms touch."""
        self._check(self.lib.tfsi_bind_matrix(self.h, _ptr(A), self._stream()))
        self._bound = A            # keep the tensor alive: the handle holds its address
        return int(self.lib.tfsi_bound_init_entries(self.h))

    def bind_matrix_pair(self, A0, A1, A_pre="compressed"):
        """two persistent matrices used alternately: the rows of the one not being assembled are re-initialised by the
        fluid kernel of the running assembly itself (tfsi_bind_matrix_pair). A_pre: the array given to compress_A_pre,
        or "compressed" once that array has been freed."""
        ap = C.c_void_p(_lib.APRE_COMPRESSED) if isinstance(A_pre, str) else _ptr(A_pre)
        self._check(self.lib.tfsi_bind_matrix_pair(self.h, _ptr(A0), _ptr(A1), ap, self._stream()))
        self._bound = (A0, A1)
        return int(self.lib.tfsi_bound_init_entries(self.h))

    def pair_fits(self):
        """a second matrix (plus the zero-fill lists) fits into the free HBM; TFSI_MATRIX_PAIR=0 switches the pair off."""
        import os
        if os.environ.get("TFSI_MATRIX_PAIR", "1") == "0":
            return False
        free, _ = torch.cuda.mem_get_info(self.device)
        return 8 * self.nnz * 1.35 < free

    def fused_launches(self):
        """fluid launches that carried re-initialisation work for the matrix pair so far."""
        return int(self.lib.tfsi_fused_launches(self.h))

    def unbind_matrix(self):
        self._check(self.lib.tfsi_unbind_matrix(self.h))
        self._bound = None

    def release_pattern(self):
        """drop the column array after set-up (A_pre assembled, facet operators set): the hot path does not read it."""
        self._check(self.lib.tfsi_release_pattern(self.h))
        self.colidx = None

    def assemble_jacobian(self, x, x1, A_pre=None, out=None, phase=0):
        """A := A_pre + J_nl(x, x1). phase 1 / 2: only the x-independent initialisation / only the cell kernels
        (a caller that exchanges halos runs phase 1 under the transfer)."""
        out = self.new_matrix() if out is None else out
        if phase:
            self._check(self.lib.tfsi_assemble_jacobian_phase(self.h, _ptr(x), _ptr(x1), self._apre(A_pre), _ptr(out),
                                                              self._stream(), int(phase)))
        else:
            self._check(self.lib.tfsi_assemble_jacobian(self.h, _ptr(x), _ptr(x1), self._apre(A_pre), _ptr(out), self._stream()))
        return out

    def assemble_residual(self, x, x1, out=None):
        out = self.new_vector() if out is None else out
        self._check(self.lib.tfsi_assemble_residual(self.h, _ptr(x), _ptr(x1), _ptr(out), self._stream()))
        return out

    def ident_zeros(self, A):
        self._check(self.lib.tfsi_ident_zeros(self.h, _ptr(A), self._stream()))

    def set_bc_dofs(self, dofs):
        d = np.ascontiguousarray(np.asarray(dofs), dtype=np.int32)
        self._check(self.lib.tfsi_set_bc_dofs(self.h, C.c_void_p(d.ctypes.data), len(d)))
        self.bc_dofs = d

    def apply_bc_matrix(self, A):
        self._check(self.lib.tfsi_apply_bc_matrix(self.h, _ptr(A), self._stream()))

    def apply_bc_rhs(self, b, x, g):
        self._check(self.lib.tfsi_apply_bc_rhs(self.h, _ptr(b), _ptr(x), _ptr(g), 0 if g is None else len(g), self._stream()))

    def apply_bc_vec(self, x, g):
        self._check(self.lib.tfsi_apply_bc_vec(self.h, _ptr(x), _ptr(g), 0 if g is None else len(g), self._stream()))

    def sumsq(self, v):
        out = C.c_double(0.0)
        self._check(self.lib.tfsi_sumsq(self.h, _ptr(v), v.numel(), C.byref(out), self._stream()))
        return out.value

    def norm(self, v):
        return float(np.sqrt(self.sumsq(v)))

    def newton_assemble_host(self, x_host, x1_host, g_host, A_pre, A, b_host, recompute=True):
        """reference-facing end-to-end call with HOST vectors (pinned torch tensors or numpy). x1_host=None re-uses
        the previous time level uploaded by the last call (it changes once per time step)."""
        ss = C.c_double(0.0)
        self._check(self.lib.tfsi_newton_assemble_host(
            self.h, _ptr(x_host), _ptr(x1_host), _ptr(g_host), self._apre(A_pre), _ptr(A), _ptr(b_host),
            C.byref(ss), int(bool(recompute)), self._stream()))
        return ss.value


def probe_fp64_peak():
    """measured FP64 FMA throughput of the device (TFLOP/s) — the compute roofline denominator."""
    lib = _lib.load()
    tf, ms = C.c_double(0.0), C.c_double(0.0)
    rc = lib.tfsi_probe_fp64_peak(C.byref(tf), C.byref(ms))
    if rc != 0:
        raise RuntimeError("fp64 probe failed: %s" % lib.tfsi_last_error(None).decode())
    return tf.value, ms.value
