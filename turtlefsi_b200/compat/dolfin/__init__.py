"""`dolfin` stand-in for running turtleFSI problem files UNMODIFIED on the B200 path (SURVEY 8b "Problem-file API").

turtleFSI problem files start with `from dolfin import *` (turtleFSI/problems/TF_fsi.py:14, TF_cfd.py:13,
TF_csm.py:11, turtle_demo.py:21) and use DOLFIN for: reading the mesh, marking sub-domains, Dirichlet conditions,
user expressions, scalar functionals (`assemble(... * ds(6))`) and point evaluation. This module provides those
names on top of turtlefsi_b200's host layer (mesh.py, boundary.py, space.py, ufl_lite.py); the hot path itself —
residual / Jacobian assembly — never goes through here. It is put on `sys.path` by `python -m turtlefsi_b200` (and
`turtlefsi_b200.compat.install()`) ONLY when no real DOLFIN can be imported; with a real DOLFIN the binding of
INTEGRATION.md §A applies instead.

Not provided (raise NotImplementedError): C++ string `Expression`s, symbolic `derivative`, `solve`, `project`,
XDMF `write_checkpoint`, MPI-parallel meshes (multi-GPU runs shard through turtlefsi_b200.distributed instead)."""
import copy as _copy
import os as _os

import numpy as np

from turtlefsi_b200 import ufl_lite as _u
from turtlefsi_b200 import boundary as _bnd
from turtlefsi_b200 import mesh as _mesh
from turtlefsi_b200.space import MixedSpace as _MixedSpace
from turtlefsi_b200.ufl_lite import (Identity, as_matrix, as_tensor, as_vector, derivative, det, dev, diff, div,  # noqa: F401
                                     cos, cosh, dot, exp, grad, inner, inv, ln, nabla_grad, outer, sin, sinh, skew, sqrt,
                                     sym, tan, tanh, tr, transpose, variable)

DOLFIN_EPS = 3.0e-16
DOLFIN_PI = float(np.pi)
pi = float(np.pi)

# `parameters["form_compiler"]["quadrature_degree"]`, `parameters["ghost_mode"]` (TF_fsi.py:22, problem_aneu.py:12-18)
parameters = {
    "form_compiler": {"quadrature_degree": -1, "optimize": True, "cpp_optimize": True, "representation": "auto"},
    "ghost_mode": "none", "allow_extrapolation": False, "reorder_dofs_serial": True, "std_out_all_processes": True,
}

_CTX = {}          # set by the driver: mesh (compat), space (native), assembler


def near(x, x0, eps=DOLFIN_EPS):
    return (x0 - eps <= x) & (x <= x0 + eps)


def between(x, interval):
    return (interval[0] <= x) & (x <= interval[1])


def set_log_level(level):
    pass


def set_log_active(flag):
    pass


def list_linear_solver_methods():
    print("device LU (cusolverRf after a host analysis) | host SuperLU — see turtlefsi_b200/linear_solver.py")


def _unsupported(name, why):
    def f(*a, **k):
        raise NotImplementedError("dolfin.%s is not available on the B200 path: %s" % (name, why))
    f.__name__ = name
    return f


Expression = _unsupported("Expression", "C++ string expressions need DOLFIN's JIT; write a UserExpression")
CompiledSubDomain = _unsupported("CompiledSubDomain", "needs DOLFIN's JIT; use AutoSubDomain")
solve = _unsupported("solve", "only the monolithic Newton solver is re-hosted")
project = _unsupported("project", "needs a mass-matrix solve outside the hot path")
File = _unsupported("File", "visualisation output is out of scope (SURVEY 8f-4)")


# ------------------------------------------------------------------------------------------------------
# MPI (one process per GPU; the problem files only ask for the rank and reduce point evaluations)
class _Comm:
    def __init__(self):
        self.rank = int(_os.environ.get("RANK", "0"))
        self.size = int(_os.environ.get("WORLD_SIZE", "1"))

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def tompi4py(self):
        return self

    def Allreduce(self, src, dst, op=None):
        # problem-file hooks see GLOBAL vectors on every rank (turtlefsi_b200.distributed), so a reduction over
        # ranks of a point evaluation is the identity
        np.copyto(dst, np.asarray(src, dtype=dst.dtype))

    def Barrier(self):
        pass

    barrier = Barrier


class MPI:
    comm_world = _Comm()
    comm_self = comm_world

    @staticmethod
    def rank(comm=None):
        return MPI.comm_world.rank

    @staticmethod
    def size(comm=None):
        return MPI.comm_world.size

    @staticmethod
    def barrier(comm=None):
        pass

    @staticmethod
    def max(comm, v):
        return v

    @staticmethod
    def min(comm, v):
        return v

    @staticmethod
    def sum(comm, v):
        return v


class Timer:
    def __init__(self, name=""):
        self.name, self._t0, self._el = name, None, 0.0

    def start(self):
        import time
        self._t0 = time.time()

    def stop(self):
        import time
        self._el = time.time() - (self._t0 or time.time())
        return self._el

    def elapsed(self):
        return (self._el, self._el, self._el)


# ------------------------------------------------------------------------------------------------------
# mesh
class _Geometry:
    def __init__(self, m):
        self._m = m

    def dim(self):
        return self._m.native.gdim


class _Topology(_Geometry):
    pass


class Mesh:
    """Mesh("file.xml[.gz]") (TF_fsi.py:68), Mesh() to be filled by XDMFFile.read (turtle_demo.py:64-66), Mesh(mesh)."""

    def __init__(self, arg=None):
        self.native = None
        self._facets = None
        if isinstance(arg, Mesh):
            self.native = _copy.deepcopy(arg.native)
        elif isinstance(arg, _mesh.Mesh):
            self.native = arg
        elif isinstance(arg, (str, _os.PathLike)):
            self.native = _mesh.read_dolfin_xml(str(arg))
        elif arg is not None and not isinstance(arg, _Comm):
            raise TypeError("Mesh(%r)" % (arg,))

    def geometry(self):
        return _Geometry(self)

    def topology(self):
        return _Topology(self)

    def ufl_cell(self):
        return "triangle" if self.native.gdim == 2 else "tetrahedron"

    def mpi_comm(self):
        return MPI.comm_world

    def num_cells(self):
        return self.native.num_cells

    def num_vertices(self):
        return self.native.num_vertices

    def num_entities_global(self, dim):
        g = self.native.gdim
        return {0: self.native.num_vertices, g: self.native.num_cells}.get(dim, self.facets().num if dim == g - 1 else None)

    num_entities = num_entities_global

    def coordinates(self):
        return self.native.coords

    def cells(self):
        return self.native.cells

    def hmin(self):
        return self.native.hmin()

    def init(self, *a):
        pass

    def facets(self):
        """facet topology of the mesh, built once (turtlefsi_b200.boundary.Facets)."""
        if self._facets is None:
            self._facets = _bnd.Facets(self.native)
        return self._facets


class Point:
    """Point(x, y[, z]) (TaylorGreen2D.py:123)."""

    def __init__(self, *c):
        self._c = np.asarray(c[0] if len(c) == 1 and np.ndim(c[0]) else c, dtype=float)

    def array(self):
        return self._c

    def x(self):
        return float(self._c[0])

    def y(self):
        return float(self._c[1])

    def __getitem__(self, i):
        return float(self._c[i])


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    """RectangleMesh(Point, Point, nx, ny, "right" | "left") in DOLFIN's vertex and cell order (TaylorGreen2D.py:123)."""
    if isinstance(p0, _Comm):
        raise NotImplementedError("RectangleMesh(comm, ...): meshes are built on every rank and cut by turtlefsi_b200.distributed")
    a, b = Point(p0).array() if not isinstance(p0, Point) else p0.array(), Point(p1).array() if not isinstance(p1, Point) else p1.array()
    return Mesh(_mesh.rectangle_mesh(tuple(a[:2]), tuple(b[:2]), int(nx), int(ny), diagonal))


def UnitSquareMesh(nx, ny, diagonal="right"):
    return RectangleMesh(Point(0.0, 0.0), Point(1.0, 1.0), nx, ny, diagonal)


def refine(mesh, *a, **k):
    """uniform Plaza refinement (TF_cfd.py:49, TF_csm.py:54)."""
    return Mesh(_mesh.refine(mesh.native))


class MeshFunction:
    """MeshFunction("size_t", mesh, dim[, value]): cell markers (dim = gdim) or facet markers (dim = gdim - 1)."""

    def __init__(self, kind, mesh, dim, value=None):
        if not isinstance(dim, (int, np.integer)):
            if isinstance(dim, MeshValueCollection):
                mvc = dim
                self.__init__(kind, mesh, mvc.dim())
                mvc._fill(self)
                return
            raise TypeError("MeshFunction(kind, mesh, dim)")
        g = mesh.native.gdim
        self._mesh, self._dim, self.kind = mesh, int(dim), kind
        if dim == g:
            self._a = np.zeros(mesh.native.num_cells, dtype=np.int64)
            self.facets = None
        elif dim == g - 1:
            self.facets = _copy.copy(mesh.facets())             # shared topology, own marker array
            self.facets.markers = np.zeros(self.facets.num, dtype=np.int64)
            self._a = self.facets.markers
        else:
            raise NotImplementedError("MeshFunction of dimension %d (only cells and facets are used on this path)" % dim)
        if value is not None:
            self._a[:] = value

    def set_all(self, v):
        self._a[:] = v

    def array(self):
        return self._a

    def dim(self):
        return self._dim

    def mesh(self):
        return self._mesh

    def size(self):
        return len(self._a)

    def where_equal(self, v):
        return np.nonzero(self._a == v)[0]

    def __getitem__(self, i):
        return self._a[int(i) if not isinstance(i, slice) else i]

    def __setitem__(self, i, v):
        self._a[i] = v

    def rename(self, *a):
        pass


class MeshValueCollection:
    """MeshValueCollection("size_t", mesh, dim): (entity, value) pairs read from a file (turtle_demo.py:76-79)."""

    def __init__(self, kind, mesh=None, dim=None):
        self._mesh, self._dim, self._entities, self._values = mesh, dim, None, None

    def dim(self):
        return self._dim

    def _fill(self, mf):
        if self._entities is None:
            return
        fac = mf.facets
        nv = self._mesh.native.num_vertices
        ent = np.sort(np.asarray(self._entities, dtype=np.int64), axis=1)
        key = lambda v: sum(v[:, c] * nv ** (v.shape[1] - 1 - c) for c in range(v.shape[1]))
        fkey = key(fac.vertices)
        order = np.argsort(fkey)
        pos = np.searchsorted(fkey[order], key(ent))
        if (pos >= len(order)).any() or (fkey[order][np.minimum(pos, len(order) - 1)] != key(ent)).any():
            raise RuntimeError("MeshValueCollection: an entity of the file is not a facet of the mesh")
        mf.array()[order[pos]] = self._values


class _cpp_mesh:
    @staticmethod
    def MeshFunctionSizet(mesh, mvc):
        return MeshFunction("size_t", mesh, mvc)


class cpp:
    mesh = _cpp_mesh


class XDMFFile:
    """reader for the shipped turtle_demo files (turtle_demo.py:58-80). h5py is absent: the raw HDF5 reader of
    turtlefsi_b200.mesh.read_turtle_demo knows the fixed layout of `mc.h5` / `mf.h5` that sit beside the .xdmf files;
    any other XDMF file is refused."""

    def __init__(self, *args):
        self.path = str(args[-1])
        self._data = None
        self._series = None
        self.parameters = {}          # flush_output / rewrite_function_mesh (problems/__init__.py:235-237): always so here

    def _load(self):
        if self._data is None:
            self._data = _mesh.read_turtle_demo(_os.path.dirname(self.path))
        return self._data

    def read(self, obj, *a):
        m, fverts, fmark = self._load()
        if isinstance(obj, Mesh):
            obj.native = m
            obj._facets = None
        elif isinstance(obj, MeshFunction):
            if obj.dim() != m.gdim:
                raise NotImplementedError("XDMFFile.read of a facet MeshFunction: use a MeshValueCollection")
            obj.array()[:] = m.markers
        elif isinstance(obj, MeshValueCollection):
            obj._entities, obj._values = fverts, fmark
        else:
            raise TypeError("XDMFFile.read(%r)" % type(obj).__name__)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def write(self, f, t=0.0):
        """XDMFFile.write(function, t) (problems/__init__.py:319-323): the VERTEX values of a sub-function (d, v, p or one
        component) as a step of a time series — DOLFIN's `write` stores CG2 functions by their vertex values too. Heavy
        data is raw binary beside the .xdmf file (turtlefsi_b200.utils.xdmf; no h5py in this image)."""
        from turtlefsi_b200.utils.xdmf import XDMFSeries
        if not isinstance(f, Function) or f.field is None:
            raise NotImplementedError("XDMFFile.write of %r: write the sub-functions of the mixed function" % type(f).__name__)
        space, x = f.space, np.asarray(f._host(), dtype=float)
        if self._series is None:
            m = f._V.mesh().native
            _os.makedirs(_os.path.dirname(self.path) or ".", exist_ok=True)
            self._series = XDMFSeries(self.path, m.coords, m.cells)
        if f.field == 2:
            vals = x[space.vertex_dof_p()]
        elif f.comp is not None:
            vals = x[space.node_dof(f.field, f.comp)[space.vertex_node]]
        else:
            vals = np.stack([x[space.node_dof(f.field, a)[space.vertex_node]] for a in range(space.gdim)], axis=1)
        self._series.write(getattr(f, "_name", "f"), vals, t)

    write_checkpoint = _unsupported("XDMFFile.write_checkpoint", "checkpoints are numpy files (turtlefsi_b200/problems/__init__.py)")


class HDF5File:
    """HDF5File(comm, "mesh.h5", "r").read(mesh | MeshFunction, "/name"[, False]) — the DOLFIN mesh files problem files
    read (Test_Cylinder/problem_aneu.py:58-66): `/mesh/{coordinates,topology}`, cell functions `/<name>/{cell_indices,
    values}`, facet functions `/<name>/{topology,values}` (matched to the mesh facets by their vertices). Read-only: h5py is
    not in this image (turtlefsi_b200.utils.hdf5_lite parses the file), fields are stored as numpy files."""

    def __init__(self, comm, path, mode="r"):
        if mode != "r":
            raise NotImplementedError("HDF5File(%r): only reading (checkpoints are numpy files, turtlefsi_b200/problems/__init__.py)" % mode)
        from turtlefsi_b200.utils.hdf5_lite import read_datasets
        self._d = read_datasets(str(path))

    def _get(self, name, key):
        k = "/" + name.strip("/") + "/" + key
        if k not in self._d:
            raise KeyError("HDF5File: no dataset %s (file holds %s)" % (k, sorted(self._d)))
        return self._d[k]

    def has_dataset(self, name):
        n = "/" + name.strip("/")
        return any(k == n or k.startswith(n + "/") for k in self._d)

    def read(self, obj, name, *a):
        if isinstance(obj, Mesh):
            X = np.array(self._get(name, "coordinates"), dtype=float)
            cells = np.sort(np.asarray(self._get(name, "topology"), dtype=np.int64), axis=1).astype(np.int32)
            obj.native = _mesh.Mesh(X, cells, np.zeros(len(cells), dtype=np.int32))
            obj._facets = None
        elif isinstance(obj, MeshFunction):
            m = obj.mesh().native
            vals = np.asarray(self._get(name, "values")).astype(np.int64)
            topo = np.sort(np.asarray(self._get(name, "topology"), dtype=np.int64), axis=1)
            nv = m.num_vertices
            key = lambda v: sum(v[:, c] * nv ** (v.shape[1] - 1 - c) for c in range(v.shape[1]))
            ents = np.asarray(m.cells, dtype=np.int64) if obj.dim() == m.gdim else np.asarray(obj.facets.vertices, dtype=np.int64)
            if topo.shape[1] != ents.shape[1]:
                raise ValueError("HDF5File.read: %s holds entities of %d vertices, the MeshFunction has dimension %d"
                                 % (name, topo.shape[1], obj.dim()))
            ekey = key(ents)
            order = np.argsort(ekey)
            pos = np.minimum(np.searchsorted(ekey[order], key(topo)), len(order) - 1)
            if (ekey[order][pos] != key(topo)).any():
                raise RuntimeError("HDF5File.read: an entity of %s is not an entity of the mesh" % name)
            obj.array()[order[pos]] = vals
        else:
            raise TypeError("HDF5File.read(%r)" % type(obj).__name__)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    write = _unsupported("HDF5File.write", "fields are stored as numpy files (turtlefsi_b200/problems/__init__.py)")


# ------------------------------------------------------------------------------------------------------
# sub-domains
def _call_inside(fn, nargs, x, on_boundary):
    return bool(fn(x, on_boundary) if nargs >= 2 else fn(x))


class SubDomain:
    """user sub-domain: override `inside(self, x, on_boundary)`. mark() follows DOLFIN: an entity is marked when
    all its vertices AND its midpoint are inside."""
    _nargs = 2

    def inside(self, x, on_boundary):
        return False

    def _inside_many(self, X, flags):
        """inside() at the points X [n, g] with per-point on_boundary flags — vectorised when the user's function
        allows it (verified on a sample), point by point otherwise (python `and` / `or` / chained comparisons)."""
        n = len(X)
        fn, na = self.inside, self._nargs
        try:
            r = fn(X.T, flags) if na >= 2 else fn(X.T)
            r = np.asarray(r)
            if r.shape == (n,) and r.dtype == bool:
                for i in np.unique(np.linspace(0, n - 1, 7).astype(int)) if n else []:
                    if bool(r[i]) != _call_inside(fn, na, X[i], bool(flags[i])):
                        raise ValueError("vectorised inside() disagrees")
                return r
        except Exception:
            pass
        return np.fromiter((_call_inside(fn, na, X[i], bool(flags[i])) for i in range(n)), dtype=bool, count=n)

    def mark(self, mf, value, check_midpoint=True):
        mesh = mf.mesh().native
        X = mesh.coords
        if mf.facets is None:
            ents = mesh.cells
            flags = np.zeros(len(ents), dtype=bool)
        else:
            ents = mf.facets.vertices
            flags = mf.facets.on_boundary
        ok = np.ones(len(ents), dtype=bool)
        for f in (False, True):                     # vertex results are shared by all entities with the same flag
            sel = flags == f
            if not sel.any():
                continue
            verts = np.unique(ents[sel])
            inv = np.zeros(mesh.num_vertices, dtype=bool)
            inv[verts] = self._inside_many(X[verts], np.full(len(verts), f))
            ok[sel] &= inv[ents[sel]].all(axis=1)
        if check_midpoint:
            cand = np.nonzero(ok)[0]
            if len(cand):
                ok[cand] &= self._inside_many(X[ents[cand]].mean(axis=1), flags[cand])
        mf.array()[ok] = value


class AutoSubDomain(SubDomain):
    def __init__(self, fn):
        import inspect
        self._fn = fn
        try:
            self._nargs = len(inspect.signature(fn).parameters)
        except (TypeError, ValueError):
            self._nargs = 1

    def inside(self, x, on_boundary=False):
        return self._fn(x, on_boundary) if self._nargs >= 2 else self._fn(x)


class DomainBoundary(SubDomain):
    def inside(self, x, on_boundary):
        return on_boundary


# ------------------------------------------------------------------------------------------------------
# coefficients
class Constant(_u.Expr):
    def __init__(self, value, cell=None, name=None):
        self.value = np.asarray(value, dtype=float)
        self.shape = self.value.shape

    def assign(self, v):
        self.value = np.asarray(float(v) if not self.shape else getattr(v, "value", v), dtype=float).reshape(self.shape)

    def values(self):
        return np.atleast_1d(self.value).ravel()

    def evaluate(self, ps, side):
        return np.broadcast_to(self.value, (ps.m, ps.nq) + self.shape)

    def __float__(self):
        return float(self.value)

    def __call__(self, x):
        return _u.Restricted(self, x) if isinstance(x, str) else self.value


class UserExpression(_u.Expr):
    """subclass with `eval(self, value, x)` and `value_shape(self)` (TF_fsi.py:120-140). Usable as a Dirichlet value,
    as a coefficient of a facet form, and called at points."""

    def __init__(self, *args, degree=None, element=None, **kwargs):
        self._degree = 2 if degree is None else int(degree)

    def value_shape(self):
        return ()

    @property
    def shape(self):
        return tuple(self.value_shape())

    def degree(self):
        return self._degree

    def ufl_shape(self):
        return self.shape

    def _eval_many(self, X):
        """values at the points X [n, g] -> [n, ncomp]. `eval` is tried once on whole arrays (x[i] = coordinate
        arrays, value[i] = component rows) and checked on a sample against point-by-point calls; anything else falls
        back to the loop DOLFIN would run."""
        X = np.atleast_2d(np.asarray(X, dtype=float))
        n = len(X)
        nv = int(np.prod(self.shape)) if self.shape else 1

        def one(i):
            v = np.zeros(nv)
            self.eval(v, X[i])
            return v
        try:
            V = np.zeros((nv, n))
            self.eval(V, X.T)
            for i in np.unique(np.linspace(0, n - 1, 5).astype(int)) if n else []:
                if not np.array_equal(V[:, i], one(i)):
                    raise ValueError("vectorised eval() disagrees")
            return V.T.copy()
        except Exception:
            pass
        return np.stack([one(i) for i in range(n)]) if n else np.zeros((0, nv))

    def evaluate(self, ps, side):
        X = ps.x.reshape(-1, ps.x.shape[-1])
        return self._eval_many(X).reshape((ps.m, ps.nq) + self.shape)

    def __call__(self, x, *a):
        if isinstance(x, str):
            return _u.Restricted(self, x)
        X = np.asarray(x, dtype=float)
        if X.ndim == 2:
            return self._eval_many(X)                   # Dirichlet protocol of turtlefsi_b200.boundary: [n, g] -> [n, ncomp]
        v = self._eval_many(X[None, :])[0]
        return v if self.shape else float(v[0])


def FacetNormal(mesh):
    return _u.FacetNormalExpr(mesh.native.gdim)


def SpatialCoordinate(mesh):
    return _u.SpatialCoordinateExpr(mesh.native.gdim)


def CellVolume(mesh):
    raise NotImplementedError("CellVolume in a problem-file form (the lifting operators use it inside the kernels)")


# ------------------------------------------------------------------------------------------------------
# elements and spaces: the path has ONE space, DVP = [P2^g, P2^g, P1 (, P2^g)] (turtleFSI/monolithic.py:77-92)
class FiniteElement:
    def __init__(self, family, cell=None, degree=None, **k):
        self.family, self.cell, self.deg, self.vector = family, cell, degree, False

    def degree(self):
        return self.deg


class VectorElement(FiniteElement):
    def __init__(self, family, cell=None, degree=None, dim=None, **k):
        super().__init__(family, cell, degree)
        self.vector = True


class MixedElement:
    def __init__(self, *elements):
        self.elements = list(elements[0]) if len(elements) == 1 and isinstance(elements[0], (list, tuple)) else list(elements)


def _check_dvp(el):
    e = el.elements
    ok = len(e) in (3, 4) and all(x.vector and x.deg == 2 for x in (e[0], e[1]) + tuple(e[3:])) and \
        (not e[2].vector and e[2].deg == 1) and all(x.family in ("CG", "Lagrange", "P") for x in e)
    if not ok:
        raise NotImplementedError("only the Taylor-Hood mixed space [P2^g, P2^g, P1 (, P2^g)] of turtleFSI/monolithic.py:"
                                  "77-92 exists on the B200 path (the kernels are written for it)")
    return len(e) == 4


class FunctionSpace:
    def __init__(self, mesh, element=None, degree=None, native=None):
        self._mesh = mesh
        if native is not None:
            self.native = native
        elif isinstance(element, MixedElement):
            self.native = _MixedSpace(mesh.native, biharmonic=_check_dvp(element))
        else:
            raise NotImplementedError("FunctionSpace(mesh, %r): only the monolithic mixed space is available" % (element,))
        self.field, self.comp = None, None

    def sub(self, i):
        s = _copy.copy(self)
        if self.field is None:
            if not 0 <= i < (4 if self.native.biharmonic else 3):
                raise IndexError("sub-space %d" % i)
            s.field = i
        elif self.comp is None and self.field != 2:
            if not 0 <= i < self.native.gdim:
                raise IndexError("component %d" % i)
            s.comp = i
        else:
            raise IndexError("no sub-space below a scalar space")
        return s

    def mesh(self):
        return self._mesh

    def dim(self):
        return self.native.num_dofs if self.field is None else len(self.native.sub_dofs(self.field, self.comp))

    def num_sub_spaces(self):
        if self.field is None:
            return 4 if self.native.biharmonic else 3
        return 0 if (self.field == 2 or self.comp is not None) else self.native.gdim

    def dofmap(self):
        return _DofMap(self)

    def collapse(self):
        return self

    def ufl_element(self):
        return self


class _DofMap:
    def __init__(self, V):
        self.V = V

    def dofs(self):
        V = self.V
        return np.arange(V.native.num_dofs) if V.field is None else np.sort(V.native.sub_dofs(V.field, V.comp))

    def cell_dofs(self, c):
        return self.V.native.cell_dofs[c]


class Vector:
    """GenericVector view of an assembler vector (device tensor on the GPU path)."""

    def __init__(self, owner):
        self._o = owner

    def get_local(self):
        return np.array(self._o._host(), dtype=float)

    def set_local(self, a):
        self._o._store(np.asarray(a, dtype=float))

    def __getitem__(self, i):
        return self._o._host()[i]

    def __setitem__(self, i, v):
        a = np.array(self._o._host(), dtype=float)
        a[i] = v
        self._o._store(a)

    def zero(self):
        self._o._store(np.zeros(self.size()))

    def axpy(self, alpha, other):
        self._o._store(self.get_local() + alpha * other.get_local())

    def norm(self, kind="l2"):
        a = self._o._host()
        return float({"l2": np.linalg.norm(a), "linf": np.abs(a).max(initial=0.0), "l1": np.abs(a).sum()}[kind])

    def size(self):
        return len(self._o._host())

    def apply(self, mode=""):
        pass

    def max(self):
        return float(self._o._host().max())

    def min(self):
        return float(self._o._host().min())


class Function(_u.FieldExpr):
    """Function(DVP) — the monolithic state (dvp_["n"] ...) — or one of its sub-functions.

    `vec` is the assembler's vector (a device tensor on the GPU path, never copied here unless a hook asks for
    values); sub-functions created with `deepcopy=True` hold a host snapshot, the others read the parent."""

    def __init__(self, V, vec=None, assembler=None, _parent=None, _field=None, _comp=None, _snapshot=None):
        self._V = V
        self._parent, self._snapshot = _parent, _snapshot
        self.assembler = assembler if assembler is not None else _CTX.get("assembler")
        self.vec = vec
        if _parent is None and vec is None:
            self.vec = self.assembler.new_vector() if self.assembler is not None else np.zeros(V.native.num_dofs)
        space = V.native
        if _field is None:
            # the mixed function itself: no tensor shape of its own in forms; use split() / sub()
            self.space, self.field, self.comp, self.gdim = space, None, None, space.gdim
            self.getvec = self._host
            self._shape = (space.ndof_cell,)
        else:
            super().__init__(space, self._host, _field, _comp)
            self._shape = self.shape

    # --- storage ---
    def _host(self):
        if self._snapshot is not None:
            return self._snapshot
        if self._parent is not None:
            return self._parent._host()
        key = id(self)
        if key in _HOST_CACHE:
            return _HOST_CACHE[key]
        a = self.vec if isinstance(self.vec, np.ndarray) else np.asarray(self.assembler.to_host(self.vec))
        if _HOST_CACHE_ON[0]:
            _HOST_CACHE[key] = a
        return a

    def _store(self, a):
        if self._parent is not None or self._snapshot is not None:
            raise NotImplementedError("writing through a sub-function: assign to the mixed function's vector()")
        if isinstance(self.vec, np.ndarray):
            self.vec[:] = a
        else:
            self.assembler.copy(self.vec, self.assembler.to_device(a))
        _HOST_CACHE.pop(id(self), None)

    # --- dolfin API ---
    def vector(self):
        if self.field is not None:
            raise NotImplementedError("vector() of a sub-function")
        return Vector(self)

    def function_space(self):
        return self._V

    def value_shape(self):
        return self._shape

    def ufl_shape(self):
        return self._shape

    def sub(self, i, deepcopy=False):
        if self.field is None:
            V = self._V.sub(i)
            snap = np.array(self._host(), dtype=float) if deepcopy else None
            return Function(V, assembler=self.assembler, _parent=None if deepcopy else self, _field=i, _snapshot=snap)
        if self.comp is None and self.field != 2:
            return Function(self._V.sub(i), assembler=self.assembler, _parent=self._parent, _field=self.field, _comp=i,
                            _snapshot=self._snapshot)
        raise IndexError("no sub-function below a scalar function")

    def split(self, deepcopy=False):
        return tuple(self.sub(i, deepcopy) for i in range(self._V.num_sub_spaces()))

    def copy(self, deepcopy=True):
        if self.field is not None:
            return Function(self._V, assembler=self.assembler, _field=self.field, _comp=self.comp,
                            _snapshot=np.array(self._host(), dtype=float))
        f = Function(self._V, assembler=self.assembler)
        f._store(np.array(self._host(), dtype=float))
        return f

    def rename(self, name, *a):
        self._name = str(name)

    def __call__(self, *x):
        if len(x) == 1 and isinstance(x[0], str):
            return _u.Restricted(self, x[0])
        pt = np.asarray(x[0] if len(x) == 1 else x, dtype=float).ravel()
        if self.field is None:
            raise NotImplementedError("point evaluation of the mixed function: evaluate a sub-function")
        mesh = self._V.mesh().native
        xh = self._host()
        if self.field == 2:
            return float(_point_eval_p(mesh, self.space, xh, pt))
        v = _bnd.point_eval(mesh, self.space, xh, pt, field=self.field)     # raises RuntimeError outside the mesh
        return v if self.comp is None else float(v[self.comp])


def _point_eval_p(mesh, space, x, point):
    g = mesh.gdim
    X = mesh.coords[mesh.cells]
    Jg = np.swapaxes(X[:, 1:] - X[:, :1], 1, 2)
    xi = np.linalg.solve(Jg, (np.asarray(point)[None, :] - X[:, 0])[:, :, None])[:, :, 0]
    lam = np.concatenate([1 - xi.sum(axis=1, keepdims=True), xi], axis=1)
    inside = np.nonzero((lam > -1e-10).all(axis=1))[0]
    if len(inside) == 0:
        raise RuntimeError("point outside the mesh")
    c = inside[0]
    n2 = space.n2
    return lam[c] @ x[space.cell_dofs[c]][2 * g * n2:2 * g * n2 + g + 1]


_HOST_CACHE = {}
_HOST_CACHE_ON = [False]


def split(f):
    return f.split()


def TestFunctions(V):
    return tuple(_u.TestFunctionExpr(V.native, i) for i in range(V.num_sub_spaces()))


def TestFunction(V):
    if V.field is None:
        raise NotImplementedError("TestFunction of the mixed space: use TestFunctions(DVP)")
    return _u.TestFunctionExpr(V.native, V.field)


def TrialFunction(V):
    return None          # only ever handed to derivative(), which the kernels replace


def assign(dst, src):
    """assign(dvp_["n"].sub(i), f) (TaylorGreen2D.py:189-192): the dofs of sub-space i of the mixed function := those of f
    (a function of the collapsed sub-space, as interpolate() returns it)."""
    if not isinstance(dst, Function) or dst.field is None or dst._parent is None:
        raise NotImplementedError("assign(): the target must be a sub-function view, e.g. dvp_['n'].sub(1)")
    if not isinstance(src, Function) or src.field != dst.field or src.comp != dst.comp:
        raise NotImplementedError("assign(): source and target must be the same sub-space")
    dofs = dst.space.sub_dofs(dst.field, dst.comp)
    x = np.array(dst._parent._host(), dtype=float)
    x[dofs] = np.asarray(src._host(), dtype=float)[dofs]
    dst._parent._store(x)


def errornorm(u, uh, norm_type="L2", degree_rise=3, mesh=None):
    """errornorm(ue, uh, "L2") (TaylorGreen2D.py:222-223) for two functions of the SAME sub-space: sqrt(int |ue - uh|^2 dx)
    with the exact element mass matrices. (DOLFIN first interpolates both into a space of higher degree; for two
    members of one Lagrange space the difference is a member of it and the norm is the same.)"""
    if str(norm_type).upper() != "L2":
        raise NotImplementedError("errornorm: only the L2 norm")
    if not (isinstance(u, Function) and isinstance(uh, Function)) or u.field is None or u.field != uh.field:
        raise NotImplementedError("errornorm: two functions of the same sub-space (interpolate the exact solution first)")
    space = uh.space
    return float(_bnd.l2_norm(uh._V.mesh().native, space, np.asarray(u._host(), dtype=float) - np.asarray(uh._host(), dtype=float), uh.field))


def interpolate(expr, V):
    """interpolate(UserExpression, DVP.sub(i).collapse()) -> nodal values written into a fresh mixed vector's field."""
    if V.field is None:
        raise NotImplementedError("interpolate into the mixed space")
    f = Function(FunctionSpace(V.mesh(), native=V.native))
    x = np.zeros(V.native.num_dofs)
    _bnd.interpolate(V.native, V.mesh().native, x, V.field, lambda X: expr(X))
    f._store(x)
    return f.sub(V.field, deepcopy=True)


# ------------------------------------------------------------------------------------------------------
# measures and assemble
class Measure(_u.Measure):
    def __init__(self, kind, domain=None, subdomain_data=None, subdomain_id=None, metadata=None):
        super().__init__(kind, domain, subdomain_data, subdomain_id, metadata)


dx = Measure("dx")
ds = Measure("ds")
dS = Measure("dS")


def _measure_context(form):
    for it in form.expression_terms():
        sd, dom = it.measure.subdomain_data, it.measure.domain
        if isinstance(sd, MeshFunction):
            return sd.mesh(), sd.facets
        if isinstance(dom, Mesh):
            return dom, None
    return _CTX.get("mesh"), None


def assemble(form, tensor=None, form_compiler_parameters=None, **k):
    """scalar functionals only (drag / lift / areas / fluxes: TF_fsi.py:229-233, TF_cfd.py:129-131,
    problem_aneu.py:158-160). Matrices and vectors of the Newton solver are assembled by the CUDA kernels through
    the C-ABI, never here. '+' of an interior facet = the incident cell with the lower index (DOLFIN's first
    incident cell — the convention the reference's TF_fsi goldens pin)."""
    if not isinstance(form, _u.Form):
        raise TypeError("assemble() expects integrand * measure")
    if form.kernel_terms() or any(it.test_function() is not None for it in form.expression_terms()):
        raise NotImplementedError("assemble() of a linear / bilinear form: the residual and the Jacobian are assembled "
                                  "by the sm_100a kernels (turtlefsi_b200.modules.newtonsolver)")
    mesh, fac = _measure_context(form)
    if mesh is None:
        raise RuntimeError("assemble(): no mesh in context")
    if fac is None:
        fac = mesh.facets() if not isinstance(_CTX.get("boundaries"), MeshFunction) else _CTX["boundaries"].facets
    deg = (form_compiler_parameters or {}).get("quadrature_degree", parameters["form_compiler"].get("quadrature_degree", -1))
    _HOST_CACHE_ON[0] = True
    try:
        return _u.assemble_scalar(form, mesh.native, _CTX.get("space"), fac, deg if deg is not None else -1, plus="lower_index")
    finally:
        _HOST_CACHE_ON[0] = False
        _HOST_CACHE.clear()


# ------------------------------------------------------------------------------------------------------
# Dirichlet conditions
class DirichletBC(_bnd.DirichletBC):
    """DirichletBC(DVP.sub(i)[.sub(j)], value, boundaries, marker) — topological (TF_fsi.py:143-185) — or
    DirichletBC(V, value, sub_domain, method="pointwise") (TaylorGreen2D.py:157-178)."""

    def __init__(self, V, value, *args, method="topological", **k):
        if V.field is None:
            raise NotImplementedError("DirichletBC on the whole mixed space")
        if isinstance(value, Constant):
            value = tuple(value.values()) if value.shape else float(value)
        elif isinstance(value, Function):
            raise NotImplementedError("DirichletBC with a Function value")
        mesh = V.mesh().native
        self._V = V
        if len(args) == 2 and isinstance(args[0], MeshFunction):
            if args[0].facets is None:
                raise ValueError("DirichletBC needs a facet MeshFunction")
            super().__init__(V.native, mesh, args[0].facets, V.field, value, int(args[1]), comp=V.comp)
        elif len(args) >= 1 and (isinstance(args[0], SubDomain) or callable(args[0])):
            sd = args[0] if isinstance(args[0], SubDomain) else AutoSubDomain(args[0])
            if method == "pointwise":
                p = _bnd.PointwiseBC(V.native, mesh, V.field, value,
                                     lambda X: sd._inside_many(X, np.zeros(len(X), dtype=bool)), comp=V.comp)
                self.__dict__.update(p.__dict__)
            else:
                mf = MeshFunction("size_t", V.mesh(), mesh.gdim - 1, 0)
                sd.mark(mf, 1)
                super().__init__(V.native, mesh, mf.facets, V.field, value, 1, comp=V.comp)
        else:
            raise TypeError("DirichletBC(V, value, boundaries, marker) or DirichletBC(V, value, sub_domain[, method])")

    def function_space(self):
        return self._V

    def apply(self, *a):
        raise NotImplementedError("bc.apply: Dirichlet rows are imposed on the device (tfsi_apply_bc_*)")


class LUSolver:
    def __init__(self, *a, **k):
        raise NotImplementedError("the linear solver is created by the driver (turtlefsi_b200.linear_solver)")


Matrix = PETScMatrix = _unsupported("Matrix", "matrices live on the device (tfsi_bind_matrix)")
