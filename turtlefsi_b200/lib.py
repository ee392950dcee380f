"""ctypes binding of libturtlefsi_b200.so (include/turtlefsi_b200.h). No compute happens in
Python; a missing library or a missing GPU raises — there is no fallback path."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFSI_LIB") or os.path.join(_HERE, "libturtlefsi_b200.so")   # TFSI_LIB: A/B builds (scripts/)
CSRC = os.path.join(_HERE, "csrc")

EXPORTS = [
    "tfsi_create", "tfsi_destroy", "tfsi_last_error", "tfsi_nnz", "tfsi_ndof_cell",
    "tfsi_assemble_A_pre", "tfsi_assemble_jacobian", "tfsi_assemble_residual", "tfsi_ident_zeros",
    "tfsi_set_bc_dofs", "tfsi_apply_bc_matrix", "tfsi_apply_bc_rhs", "tfsi_apply_bc_vec",
    "tfsi_sumsq", "tfsi_newton_assemble_host", "tfsi_halo_pack", "tfsi_halo_unpack",
    "tfsi_launch_count", "tfsi_probe_fp64_peak", "tfsi_expand_node_pattern", "tfsi_set_profiling", "tfsi_kernel_time", "tfsi_compress_A_pre", "tfsi_compressed_nnz", "tfsi_set_facet_operator", "tfsi_set_facet_operator_prev", "tfsi_set_facet_load", "tfsi_set_facet_load_scale", "tfsi_csr_matvec",
    "tfsi_invalidate_A_pre", "tfsi_bind_matrix", "tfsi_unbind_matrix", "tfsi_bound_init_entries", "tfsi_release_pattern",
    "tfsi_num_colours", "tfsi_assemble_jacobian_phase", "tfsi_bind_matrix_pair", "tfsi_fused_launches",
]

APRE_COMPRESSED = 1     # TFSI_APRE_COMPRESSED sentinel (include/turtlefsi_b200.h)


class TfsiRegion(C.Structure):
    _fields_ = [("marker", C.c_int32), ("kind", C.c_int32), ("rho", C.c_double),
                ("mu", C.c_double), ("lam", C.c_double),
                ("material", C.c_int32), ("viscoelasticity", C.c_int32),
                ("C01", C.c_double), ("C10", C.c_double), ("C11", C.c_double), ("Jm", C.c_double),
                ("mu_visc", C.c_double), ("lambda_visc", C.c_double)]


# solid_properties["material_model"] (common.py:104-121) -> TFSI_MAT_*; ["viscoelasticity"] (solid.py:46-59)
MATERIAL_MODELS = {"StVenantKirchoff": 0, "StVenantKirchoffEnergy": 1, "NeoHookean": 2, "MooneyRivlin": 3, "Gent": 4}
VISCOELASTICITY = {None: 0, "None": 0, "Form1": 1, "Form2": 2}


class TfsiProblem(C.Structure):
    _fields_ = [
        ("gdim", C.c_int32), ("biharmonic", C.c_int32),
        ("num_vertices", C.c_int64), ("num_cells", C.c_int64), ("num_dofs", C.c_int64),
        ("coords", C.c_void_p), ("cell_vertices", C.c_void_p), ("cell_dofs", C.c_void_p),
        ("cell_marker", C.c_void_p), ("rowptr", C.c_void_p), ("colidx", C.c_void_p),
        ("num_regions", C.c_int32), ("regions", C.POINTER(TfsiRegion)),
        ("dt", C.c_double), ("theta", C.c_double),
        ("fluid_on", C.c_int32), ("solid_on", C.c_int32),
        ("extrapolation", C.c_int32), ("extrapolation_sub_type", C.c_int32),
        ("hmin", C.c_double), ("gravity", C.c_double), ("has_gravity", C.c_int32),
        ("nq_jacobian", C.c_int32), ("qpts_jacobian", C.c_void_p), ("qwts_jacobian", C.c_void_p),
        ("nq_residual", C.c_int32), ("qpts_residual", C.c_void_p), ("qwts_residual", C.c_void_p),
        ("verify_pattern", C.c_int32), ("borrow_pattern", C.c_int32), ("deterministic", C.c_int32),
    ]


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a of the single translation unit (in-tree)."""
    srcs = [os.path.join(CSRC, f) for f in ("tfsi_abi.cu", "tfsi_kernels.cuh", "tfsi_math.cuh")]
    srcs.append(os.path.join(_HERE, "..", "include", "turtlefsi_b200.h"))
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH, srcs[0]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libturtlefsi_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "turtlefsi_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, dbl, i64, i32 = C.c_void_p, C.c_double, C.c_int64, C.c_int32
    lib.tfsi_create.argtypes = [C.POINTER(TfsiProblem), C.POINTER(vp)]
    lib.tfsi_destroy.argtypes = [vp]
    lib.tfsi_destroy.restype = None
    lib.tfsi_last_error.argtypes = [vp]
    lib.tfsi_last_error.restype = C.c_char_p
    lib.tfsi_nnz.argtypes = [vp]
    lib.tfsi_nnz.restype = i64
    lib.tfsi_ndof_cell.argtypes = [vp]
    lib.tfsi_launch_count.argtypes = [vp]
    lib.tfsi_launch_count.restype = i64
    lib.tfsi_assemble_A_pre.argtypes = [vp, vp, vp, vp, vp]
    lib.tfsi_assemble_jacobian.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.tfsi_assemble_residual.argtypes = [vp, vp, vp, vp, vp]
    lib.tfsi_assemble_jacobian_phase.argtypes = [vp, vp, vp, vp, vp, vp, i32]
    lib.tfsi_ident_zeros.argtypes = [vp, vp, vp]
    lib.tfsi_set_bc_dofs.argtypes = [vp, vp, i64]
    lib.tfsi_apply_bc_matrix.argtypes = [vp, vp, vp]
    lib.tfsi_apply_bc_rhs.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.tfsi_apply_bc_vec.argtypes = [vp, vp, vp, i64, vp]
    lib.tfsi_sumsq.argtypes = [vp, vp, i64, C.POINTER(dbl), vp]
    lib.tfsi_newton_assemble_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(dbl), i32, vp]
    lib.tfsi_halo_pack.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.tfsi_halo_unpack.argtypes = [vp, vp, vp, i64, vp, vp]
    lib.tfsi_expand_node_pattern.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tfsi_set_facet_operator.argtypes = [vp, vp, vp, vp, i64]
    lib.tfsi_set_facet_operator_prev.argtypes = [vp, vp, vp, vp, i64]
    lib.tfsi_set_facet_load.argtypes = [vp, vp, vp, i64]
    lib.tfsi_set_facet_load_scale.argtypes = [vp, C.c_double]
    lib.tfsi_csr_matvec.argtypes = [vp, vp, vp, vp, vp]
    lib.tfsi_compress_A_pre.argtypes = [vp, vp, vp]
    lib.tfsi_compressed_nnz.argtypes = [vp]
    lib.tfsi_compressed_nnz.restype = i64
    lib.tfsi_invalidate_A_pre.argtypes = [vp]
    lib.tfsi_bind_matrix.argtypes = [vp, vp, vp]
    lib.tfsi_unbind_matrix.argtypes = [vp]
    lib.tfsi_fused_launches.argtypes = [vp]
    lib.tfsi_fused_launches.restype = i64
    lib.tfsi_bind_matrix_pair.argtypes = [vp, vp, vp, vp, vp]
    lib.tfsi_bound_init_entries.argtypes = [vp]
    lib.tfsi_bound_init_entries.restype = i64
    lib.tfsi_release_pattern.argtypes = [vp]
    lib.tfsi_num_colours.argtypes = [vp]
    lib.tfsi_set_profiling.argtypes = [vp, i32]
    lib.tfsi_kernel_time.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(i64)]
    lib.tfsi_probe_fp64_peak.argtypes = [C.POINTER(dbl), C.POINTER(dbl)]
    _lib = lib
    return lib
