"""ORACLE (test infrastructure only): ctypes loader of oracle/liboracle_assemble.so, the plain-C
restatement used as the timed CPU baseline ("port") and as a second checker."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle_assemble.so")


class OracleRegion(C.Structure):
    _fields_ = [("marker", C.c_int32), ("kind", C.c_int32), ("rho", C.c_double), ("mu", C.c_double),
                ("lam", C.c_double)]


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _PATH


_native = None


def build_native():
    """-march=native build for the TIMED CPU baseline, compiled on the machine that runs it (the portable
    liboracle_assemble.so travels to the GPU box with the repo snapshot; a binary tuned for this container's CPU must
    not). Falls back to the portable build when the compiler is missing."""
    global _native
    if _native is not None:
        return _native
    import hashlib
    import tempfile
    try:
        flags = open("/proc/cpuinfo").read().split("flags", 1)[1].split("\n", 1)[0]
    except Exception:
        flags = "unknown"
    src = os.path.join(_HERE, "c", "oracle_assemble.c")
    tag = hashlib.sha1((flags + open(src).read()).encode()).hexdigest()[:12]
    out = os.path.join(tempfile.gettempdir(), "tfsi_oracle_native_%s.so" % tag)
    if not os.path.exists(out):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        tmp = out + ".%d.tmp" % os.getpid()
        res = subprocess.run([cc, "-O3", "-march=native", "-fopenmp", "-fPIC", "-std=c11", "-shared", "-o", tmp, src, "-lm"],
                             capture_output=True)
        if res.returncode != 0:
            build()
            return _PATH
        os.replace(tmp, out)
    _native = out
    return out


def load():
    if _native is not None:
        return C.CDLL(_native)
    if not os.path.exists(_PATH):
        build()
    return C.CDLL(_PATH)


def _regs(regions):
    arr = (OracleRegion * len(regions))()
    for i, r in enumerate(regions):
        arr[i].marker, arr[i].kind = int(r["marker"]), 0 if r["kind"] == "fluid" else 1
        arr[i].rho, arr[i].mu, arr[i].lam = float(r["rho"]), float(r["mu"]), float(r.get("lambda_", 0.0))
    return arr


def _p(a):
    return C.c_void_p(a.ctypes.data)


def max_threads():
    return int(load().tfsi_oracle_max_threads())


def assemble_jacobian(mesh, space, regions, dt, theta, x, x1, rule, pattern, A_vals, c0=0, c1=None, nthreads=0,
                      part="nl", ext=1):
    lib = load()
    pts, wts = rule
    c1 = mesh.num_cells if c1 is None else c1
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
    cd = np.ascontiguousarray(space.cell_dofs, dtype=np.int32)
    mk = np.ascontiguousarray(mesh.markers, dtype=np.int32)
    rowptr = np.ascontiguousarray(pattern[0], dtype=np.int64)
    colidx = np.ascontiguousarray(pattern[1], dtype=np.int32)
    regs = _regs(regions)
    rc = lib.tfsi_oracle_assemble_jacobian(
        C.c_int(mesh.gdim), C.c_int(len(wts)), _p(pts), _p(wts), C.c_int64(c0), C.c_int64(c1),
        _p(mesh.coords), _p(cells), _p(cd), _p(mk), C.c_int(space.ndof_cell), C.c_int(len(regions)), regs,
        C.c_double(dt), C.c_double(theta), _p(x), _p(x1), _p(rowptr), _p(colidx), _p(A_vals), C.c_int(nthreads),
        C.c_int({"nl": 0, "lin": 1, "all": 2}[part]), C.c_int(ext))
    if rc != 0:
        raise RuntimeError("oracle C assembly failed (%d)" % rc)
    return A_vals


def assemble_residual(mesh, space, regions, dt, theta, x, x1, rule, b, c0=0, c1=None, nthreads=0, ext=1,
                      gravity=0.0):
    lib = load()
    pts, wts = rule
    c1 = mesh.num_cells if c1 is None else c1
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
    cd = np.ascontiguousarray(space.cell_dofs, dtype=np.int32)
    mk = np.ascontiguousarray(mesh.markers, dtype=np.int32)
    regs = _regs(regions)
    rc = lib.tfsi_oracle_assemble_residual(
        C.c_int(mesh.gdim), C.c_int(len(wts)), _p(pts), _p(wts), C.c_int64(c0), C.c_int64(c1),
        _p(mesh.coords), _p(cells), _p(cd), _p(mk), C.c_int(space.ndof_cell), C.c_int(len(regions)), regs,
        C.c_double(dt), C.c_double(theta), _p(x), _p(x1), _p(b), C.c_int(nthreads), C.c_int(ext), C.c_double(gravity))
    if rc != 0:
        raise RuntimeError("oracle C residual failed (%d)" % rc)
    return b
