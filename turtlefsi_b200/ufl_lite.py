"""A small form language for the HOST side of the drop-in — what the reference's problem files and physics
modules write in UFL, without UFL/FFC/DOLFIN.

What the reference does with UFL on this path, and what stands in for it here:

  * `fluid_setup / solid_setup / extrapolate_setup` return `F_*_linear`, `F_*_nonlinear` as sums of cell integrals
    (turtleFSI/modules/fluid.py:58, solid.py:89, laplace.py:48) which `solver_setup` adds up and differentiates
    (newtonsolver.py:18-27). Here those cell integrals are `KernelIntegral` terms — NAMES of the integrals the
    sm_100a kernels implement, with the region marker and its constants — and `Form` is their sum. The assembler is
    configured from the terms that are present (`regions_of`), so a `--fluid/--solid/--extrapolation` module decides
    what is assembled by what it returns, as in the reference.
  * problem files add facet terms to those forms (Test_Cylinder/problem_aneu.py:214:
    `F_solid_linear += p(t) * inner(n('+'), psi('+')) * dS(fsi_id)`) and evaluate scalar functionals
    (TF_fsi.py:229-233: `assemble((sigma(v, p, d, mu_f) * n)[0] * ds(6))`, `("+")` restrictions on `dS`). Those are
    real expression trees here (`Expr`), evaluated by quadrature on facets / cells with numpy: `assemble_scalar`,
    `assemble_load`. Facet terms that depend on the unknown are not accepted (the shipped modules' own ones — Robin,
    biharmonic ds — are fixed operators, modules/solid.py, modules/biharmonic.py).

Conventions [dep: UFL/FFC 2019.1]: quadrature degree of a functional = estimated polynomial degree (sum over
products, max over sums, quotient = sum, grad lowers by one on affine cells) unless
`parameters["form_compiler"]["quadrature_degree"]` is set; facet rule = Gauss-Legendre on edges, the default
triangle scheme on faces; `*` is UFL's (scalar scaling, matrix-vector, matrix-matrix)."""
import numpy as np

from .element import tabulate

# ------------------------------------------------------------------------------------------------------
# expression tree


def as_expr(v):
    if isinstance(v, Expr):
        return v
    if isinstance(v, (int, float, np.integer, np.floating)):
        return Const(float(v))
    if isinstance(v, (list, tuple, np.ndarray)):
        if all(isinstance(c, (int, float, np.integer, np.floating)) for c in np.ravel(np.asarray(v, dtype=object))):
            return Const(np.asarray(v, dtype=float))
        return ListTensor([as_expr(c) for c in v])
    raise TypeError("cannot use %r in a form" % (v,))


def _bcast(a, ra, b, rb):
    """align two evaluated values (array, tensor rank) for an elementwise operation of equal tensor rank, or of a
    scalar with a tensor."""
    if ra < rb:
        a = a.reshape(a.shape + (1,) * (rb - ra))
    elif rb < ra:
        b = b.reshape(b.shape + (1,) * (ra - rb))
    return a, b


class Expr:
    """node of the tree: `shape` is the tensor shape; `evaluate(ps, side)` returns an array
    [..., m, nq, *shape] (leading axes appear only below a test function); `degree()` the estimated degree."""
    shape = ()

    @property
    def rank(self):
        return len(self.shape)

    # --- algebra --------------------------------------------------------------------------------------
    def __add__(self, o):
        if isinstance(o, (Form, Measure)):
            return NotImplemented
        return Sum(self, as_expr(o), 1.0)

    def __radd__(self, o):
        return Sum(as_expr(o), self, 1.0)

    def __sub__(self, o):
        return Sum(self, as_expr(o), -1.0)

    def __rsub__(self, o):
        return Sum(as_expr(o), self, -1.0)

    def __neg__(self):
        return Product(Const(-1.0), self)

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([Integral(self, o, 1.0)])
        if isinstance(o, Form):
            return NotImplemented
        return Product(self, as_expr(o))

    def __rmul__(self, o):
        return Product(as_expr(o), self)

    def __truediv__(self, o):
        return Division(self, as_expr(o))

    def __rtruediv__(self, o):
        return Division(as_expr(o), self)

    def __pow__(self, e):
        return Power(self, as_expr(e))

    def __getitem__(self, idx):
        return Indexed(self, idx if isinstance(idx, tuple) else (idx,))

    def __len__(self):
        if not self.shape:
            raise TypeError("scalar expression has no len()")
        return self.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __call__(self, side):
        if side not in ("+", "-"):
            raise ValueError("restriction must be '+' or '-'")
        return Restricted(self, side)

    @property
    def T(self):
        return Transposed(self)

    def dx(self, i):
        return Indexed(Grad(self), (Ellipsis, i))

    def children(self):
        return ()

    def walk(self):
        yield self
        for c in self.children():
            yield from c.walk()

    def degree(self):
        return 0


class Const(Expr):
    def __init__(self, value):
        self.value = np.asarray(value, dtype=float)
        self.shape = self.value.shape

    def evaluate(self, ps, side):
        return np.broadcast_to(self.value, (ps.m, ps.nq) + self.shape)

    def __float__(self):
        return float(self.value)


class Sum(Expr):
    def __init__(self, a, b, sign):
        if a.shape != b.shape:
            raise ValueError("shape mismatch in sum: %s vs %s" % (a.shape, b.shape))
        self.a, self.b, self.sign, self.shape = a, b, sign, a.shape

    def children(self):
        return (self.a, self.b)

    def evaluate(self, ps, side):
        return self.a.evaluate(ps, side) + self.sign * self.b.evaluate(ps, side)

    def degree(self):
        return max(self.a.degree(), self.b.degree())


class Product(Expr):
    """UFL `*`: scalar * anything, or contraction of the last index of a with the first of b."""

    def __init__(self, a, b):
        self.a, self.b = a, b
        if a.rank == 0 or b.rank == 0:
            self.shape = a.shape or b.shape
        else:
            if a.shape[-1] != b.shape[0]:
                raise ValueError("cannot multiply shapes %s and %s" % (a.shape, b.shape))
            if a.rank == 1 and b.rank == 1:
                raise ValueError("vector * vector is not defined: use dot or inner")
            self.shape = a.shape[:-1] + b.shape[1:]

    def children(self):
        return (self.a, self.b)

    def evaluate(self, ps, side):
        A, B = self.a.evaluate(ps, side), self.b.evaluate(ps, side)
        if self.a.rank == 0 or self.b.rank == 0:
            A, B = _bcast(A, self.a.rank, B, self.b.rank)
            return A * B
        return _contract(A, self.a.rank, B, self.b.rank)

    def degree(self):
        return self.a.degree() + self.b.degree()


def _contract(A, ra, B, rb):
    """sum over the last tensor index of A and the first tensor index of B (batch axes broadcast)."""
    A = A.reshape(A.shape + (1,) * (rb - 1))                                         # [..., a.., k, 1..]
    B = B.reshape(B.shape[:B.ndim - rb] + (1,) * (ra - 1) + B.shape[B.ndim - rb:])   # [..., 1.., k, b..]
    return (A * B).sum(axis=-rb)


class Division(Expr):
    def __init__(self, a, b):
        if b.rank:
            raise ValueError("division by a non-scalar")
        self.a, self.b, self.shape = a, b, a.shape

    def children(self):
        return (self.a, self.b)

    def evaluate(self, ps, side):
        A, B = _bcast(self.a.evaluate(ps, side), self.a.rank, self.b.evaluate(ps, side), 0)
        return A / B

    def degree(self):
        return self.a.degree() + self.b.degree()


class Power(Expr):
    def __init__(self, a, e):
        if a.rank or e.rank:
            raise ValueError("power of non-scalars")
        self.a, self.e = a, e

    def children(self):
        return (self.a, self.e)

    def evaluate(self, ps, side):
        return self.a.evaluate(ps, side) ** self.e.evaluate(ps, side)

    def degree(self):
        if isinstance(self.e, Const) and float(self.e) >= 0 and float(self.e) == int(float(self.e)):
            return self.a.degree() * int(float(self.e))
        return self.a.degree() + 2


class MathFunction(Expr):
    def __init__(self, fn, a):
        if a.rank:
            raise ValueError("math function of a non-scalar")
        self.fn, self.a = fn, a

    def children(self):
        return (self.a,)

    def evaluate(self, ps, side):
        return self.fn(self.a.evaluate(ps, side))

    def degree(self):
        return self.a.degree() + 2


class Indexed(Expr):
    def __init__(self, a, idx):
        self.a = a
        if Ellipsis in idx:
            k = idx.index(Ellipsis)
            idx = idx[:k] + (slice(None),) * (a.rank - len(idx) + 1) + idx[k + 1:]
        if len(idx) > a.rank:
            raise IndexError("too many indices")
        idx = tuple(int(i) if not isinstance(i, slice) else i for i in idx) + (slice(None),) * (a.rank - len(idx))
        for i, n in zip(idx, a.shape):
            if not isinstance(i, slice) and not (-n <= i < n):
                raise IndexError("index %d out of range for shape %s" % (i, a.shape))
        self.idx = idx
        self.shape = tuple(n for i, n in zip(idx, a.shape) if isinstance(i, slice))

    def children(self):
        return (self.a,)

    def evaluate(self, ps, side):
        return self.a.evaluate(ps, side)[(Ellipsis,) + self.idx]

    def degree(self):
        return self.a.degree()


class ListTensor(Expr):
    def __init__(self, comps):
        self.comps = list(comps)
        if any(c.shape != self.comps[0].shape for c in self.comps):
            raise ValueError("ragged tensor")
        self.shape = (len(self.comps),) + self.comps[0].shape

    def children(self):
        return tuple(self.comps)

    def evaluate(self, ps, side):
        vals = np.broadcast_arrays(*[c.evaluate(ps, side) for c in self.comps])
        r = self.comps[0].rank
        return np.stack(vals, axis=vals[0].ndim - r)

    def degree(self):
        return max(c.degree() for c in self.comps)


class Transposed(Expr):
    def __init__(self, a):
        if a.rank != 2:
            raise ValueError("transpose of a rank-%d tensor" % a.rank)
        self.a, self.shape = a, a.shape[::-1]

    def children(self):
        return (self.a,)

    def evaluate(self, ps, side):
        return np.swapaxes(self.a.evaluate(ps, side), -1, -2)

    def degree(self):
        return self.a.degree()


class Restricted(Expr):
    def __init__(self, a, side):
        self.a, self.side, self.shape = a, side, a.shape
        if hasattr(a, "evaluate_grad"):                      # grad(v("+")) == grad(v)("+")
            self.gdim = a.gdim
            self.evaluate_grad = lambda ps, side: a.evaluate_grad(ps, self.side)

    def children(self):
        return (self.a,)

    def evaluate(self, ps, side):
        return self.a.evaluate(ps, self.side)

    def degree(self):
        return self.a.degree()


class IdentityExpr(Const):
    def __init__(self, n):
        super().__init__(np.eye(n))


class MatrixOp(Expr):
    """inv / det / tr / sym / skew / dev of a square matrix."""

    def __init__(self, op, a):
        if a.rank != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("%s needs a square matrix" % op)
        self.op, self.a = op, a
        self.shape = () if op in ("det", "tr") else a.shape

    def children(self):
        return (self.a,)

    def evaluate(self, ps, side):
        A = self.a.evaluate(ps, side)
        if self.op == "inv":
            return np.linalg.inv(A)
        if self.op == "det":
            return np.linalg.det(A)
        if self.op == "tr":
            return np.trace(A, axis1=-2, axis2=-1)
        if self.op == "sym":
            return 0.5 * (A + np.swapaxes(A, -1, -2))
        if self.op == "skew":
            return 0.5 * (A - np.swapaxes(A, -1, -2))
        if self.op == "dev":
            n = A.shape[-1]
            return A - np.trace(A, axis1=-2, axis2=-1)[..., None, None] * np.eye(n) / n
        raise ValueError(self.op)

    def degree(self):
        n, d = self.a.shape[0], self.a.degree()
        if self.op == "det":
            return n * d
        if self.op == "inv":                       # adj / det after UFL's algebra lowering
            return (n - 1) * d + n * d
        return d


class Contraction(Expr):
    """dot (last index with first index) and inner (all indices)."""

    def __init__(self, kind, a, b):
        self.kind, self.a, self.b = kind, a, b
        if kind == "inner":
            if a.shape != b.shape:
                raise ValueError("inner: shapes %s and %s" % (a.shape, b.shape))
            self.shape = ()
        elif kind == "outer":
            self.shape = a.shape + b.shape
        else:
            if a.rank == 0 or b.rank == 0:
                raise ValueError("dot needs tensors of rank >= 1")
            if a.shape[-1] != b.shape[0]:
                raise ValueError("dot: shapes %s and %s" % (a.shape, b.shape))
            self.shape = a.shape[:-1] + b.shape[1:]

    def children(self):
        return (self.a, self.b)

    def evaluate(self, ps, side):
        A, B = self.a.evaluate(ps, side), self.b.evaluate(ps, side)
        if self.kind == "inner":
            r = self.a.rank
            return (A * B).sum(axis=tuple(range(-r, 0))) if r else A * B
        if self.kind == "outer":
            A = A.reshape(A.shape + (1,) * self.b.rank)
            B = B.reshape(B.shape[:B.ndim - self.b.rank] + (1,) * self.a.rank + B.shape[B.ndim - self.b.rank:])
            return A * B
        return _contract(A, self.a.rank, B, self.b.rank)

    def degree(self):
        return self.a.degree() + self.b.degree()


class Grad(Expr):
    def __init__(self, a):
        if not hasattr(a, "evaluate_grad"):
            raise NotImplementedError("grad() of a derived expression is not supported on this path: "
                                      "apply grad to the fields (d, v, p, w, test functions, x) directly")
        self.a = a
        self.shape = a.shape + (a.gdim,)

    def children(self):
        return (self.a,)

    def evaluate(self, ps, side):
        return self.a.evaluate_grad(ps, side)

    def degree(self):
        return max(self.a.degree() - 1, 0)


class FacetNormalExpr(Expr):
    def __init__(self, gdim):
        self.gdim, self.shape = gdim, (gdim,)

    def evaluate(self, ps, side):
        if ps.normal is None:
            raise ValueError("FacetNormal in a cell integral")
        return np.broadcast_to(ps.normal[side][:, None, :], (ps.m, ps.nq, self.gdim))


class SpatialCoordinateExpr(Expr):
    def __init__(self, gdim):
        self.gdim, self.shape = gdim, (gdim,)

    def evaluate(self, ps, side):
        return ps.x

    def evaluate_grad(self, ps, side):
        return np.broadcast_to(np.eye(self.gdim), (ps.m, ps.nq, self.gdim, self.gdim))

    def degree(self):
        return 1


class PointFunctionExpr(Expr):
    """a coefficient given by a callable of the physical points, X [n, g] -> [n, *shape] (Expression, Constant
    that changes in time): evaluated afresh at every assembly."""

    def __init__(self, fn, shape, degree=2):
        self.fn, self.shape, self._deg = fn, tuple(shape), degree

    def evaluate(self, ps, side):
        X = ps.x.reshape(-1, ps.x.shape[-1])
        return np.asarray(self.fn(X), dtype=float).reshape((ps.m, ps.nq) + self.shape)

    def degree(self):
        return self._deg


class FieldExpr(Expr):
    """sub-function `field` (0 = d, 1 = v, 2 = p, 3 = w; optionally one component) of a state vector of the mixed
    space. `getvec()` returns the current host copy of the vector."""

    def __init__(self, space, getvec, field, comp=None):
        self.space, self.getvec, self.field, self.comp = space, getvec, field, comp
        self.gdim = space.gdim
        self.shape = () if (field == 2 or comp is not None) else (space.gdim,)

    def _coeffs(self, ps, side):
        g, n2 = self.space.gdim, self.space.n2
        cd = self.space.cell_dofs[ps.cell[side]]
        x = self.getvec()
        if self.field == 2:
            return x[cd[:, 2 * g * n2:2 * g * n2 + g + 1]]                       # [m, n1]
        off = {0: 0, 1: g * n2, 3: 2 * g * n2 + g + 1}[self.field]
        co = x[cd[:, off:off + g * n2]].reshape(len(cd), g, n2)                  # [m, g, n2]
        return co if self.comp is None else co[:, self.comp]

    def evaluate(self, ps, side):
        co = self._coeffs(ps, side)
        phi2, _, phi1 = ps.tab(side)
        if self.field == 2:
            return np.einsum('mn,mqn->mq', co, phi1)
        if self.comp is not None:
            return np.einsum('mn,mqn->mq', co, phi2)
        return np.einsum('man,mqn->mqa', co, phi2)

    def evaluate_grad(self, ps, side):
        co = self._coeffs(ps, side)
        if self.field == 2:
            gp = ps.grad_phi1(side)
            return np.einsum('mn,mqnb->mqb', co, gp)
        gp = ps.grad_phi2(side)
        if self.comp is not None:
            return np.einsum('mn,mqnb->mqb', co, gp)
        return np.einsum('man,mqnb->mqab', co, gp)

    def degree(self):
        return 1 if self.field == 2 else 2


class TestFunctionExpr(Expr):
    """test function of sub-space `field` of the mixed space (phi: 0, psi: 1, gamma: 2, beta: 3). Evaluates to ALL
    basis functions of the cell at once: a leading axis of length nb = g * n2 (index a * n2 + i) or n1."""

    def __init__(self, space, field):
        self.space, self.field, self.gdim = space, field, space.gdim
        self.shape = () if field == 2 else (space.gdim,)

    def evaluate(self, ps, side):
        ps.arg = (self.field, side)
        phi2, _, phi1 = ps.tab(side)
        if self.field == 2:
            return np.moveaxis(phi1, -1, 0)                                      # [n1, m, nq]
        g, n2 = self.gdim, self.space.n2
        out = np.zeros((g, n2, ps.m, ps.nq, g))
        for a in range(g):
            out[a, :, :, :, a] = np.moveaxis(phi2, -1, 0)
        return out.reshape(g * n2, ps.m, ps.nq, g)

    def evaluate_grad(self, ps, side):
        ps.arg = (self.field, side)
        if self.field == 2:
            return np.moveaxis(ps.grad_phi1(side), 2, 0)                         # [n1, m, nq, g]
        g, n2 = self.gdim, self.space.n2
        gp = np.moveaxis(ps.grad_phi2(side), 2, 0)                               # [n2, m, nq, g]
        out = np.zeros((g, n2, ps.m, ps.nq, g, g))
        for a in range(g):
            out[a, :, :, :, a, :] = gp
        return out.reshape(g * n2, ps.m, ps.nq, g, g)

    def degree(self):
        return 1 if self.field == 2 else 2


# free functions with UFL's names
def grad(f):
    return Grad(as_expr(f))


def nabla_grad(f):
    g = Grad(as_expr(f))
    return Transposed(g) if g.rank == 2 else g


def div(f):
    f = as_expr(f)
    if f.rank != 1:
        raise NotImplementedError("div of rank-%d" % f.rank)
    return MatrixOp("tr", Grad(f))


def dot(a, b):
    return Contraction("dot", as_expr(a), as_expr(b))


def inner(a, b):
    return Contraction("inner", as_expr(a), as_expr(b))


def outer(a, b):
    return Contraction("outer", as_expr(a), as_expr(b))


def inv(a):
    return MatrixOp("inv", as_expr(a))


def det(a):
    return MatrixOp("det", as_expr(a))


def tr(a):
    return MatrixOp("tr", as_expr(a))


def sym(a):
    return MatrixOp("sym", as_expr(a))


def skew(a):
    return MatrixOp("skew", as_expr(a))


def dev(a):
    return MatrixOp("dev", as_expr(a))


def transpose(a):
    return Transposed(as_expr(a))


def Identity(n):
    return IdentityExpr(n)


def as_vector(comps):
    return ListTensor([as_expr(c) for c in comps])


def as_tensor(rows):
    return ListTensor([as_vector(r) if isinstance(r, (list, tuple)) else as_expr(r) for r in rows])


def as_matrix(rows):
    return as_tensor(rows)


def _math(fn, a):
    """UFL's math functions evaluate plain numbers at once (UserExpression.eval of a problem file does
    `value[0] = sin(pi * x[0]) * exp(-t)`, TaylorGreen2D.py:143-144) and build an expression node otherwise."""
    if isinstance(a, (int, float, np.integer, np.floating)):
        return float(fn(a))
    return MathFunction(fn, as_expr(a))


def sqrt(a):
    return _math(np.sqrt, a)


def exp(a):
    return _math(np.exp, a)


def ln(a):
    return _math(np.log, a)


def sin(a):
    return _math(np.sin, a)


def cos(a):
    return _math(np.cos, a)


def tan(a):
    return _math(np.tan, a)


def sinh(a):
    return _math(np.sinh, a)


def cosh(a):
    return _math(np.cosh, a)


def tanh(a):
    return _math(np.tanh, a)


def _no_symbolic(name):
    def f(*a, **k):
        raise NotImplementedError("%s(): symbolic differentiation is not part of the B200 path — the Jacobian of the "
                                  "built-in forms is hand-derived inside the CUDA kernels" % name)
    f.__name__ = name
    return f


derivative = _no_symbolic("derivative")
diff = _no_symbolic("diff")
variable = _no_symbolic("variable")


# ------------------------------------------------------------------------------------------------------
# measures, integrals, forms
class Measure:
    """Measure("dx" | "ds" | "dS", domain=mesh, subdomain_data=markers); `ds(6)` selects marker 6."""

    def __init__(self, kind, domain=None, subdomain_data=None, subdomain_id=None, metadata=None):
        if kind not in ("dx", "ds", "dS"):
            raise ValueError("unknown measure %r" % kind)
        self.kind, self.domain, self.subdomain_data = kind, domain, subdomain_data
        self.subdomain_id, self.metadata = subdomain_id, metadata

    def __call__(self, subdomain_id=None, domain=None, subdomain_data=None, metadata=None, degree=None):
        md = dict(self.metadata or {})
        md.update(metadata or {})
        if degree is not None:
            md["quadrature_degree"] = degree
        return Measure(self.kind, domain if domain is not None else self.domain,
                       subdomain_data if subdomain_data is not None else self.subdomain_data,
                       subdomain_id if subdomain_id is not None else self.subdomain_id, md or None)

    def __rmul__(self, integrand):
        return Form([Integral(as_expr(integrand), self, 1.0)])


class Integral:
    def __init__(self, integrand, measure, scale):
        self.integrand, self.measure, self.scale = integrand, measure, float(scale)

    def scaled(self, s):
        return Integral(self.integrand, self.measure, self.scale * s)

    def test_function(self):
        args = [n for n in self.integrand.walk() if isinstance(n, TestFunctionExpr)]
        return args[0] if args else None

    def depends_on_state(self):
        return any(isinstance(n, FieldExpr) for n in self.integrand.walk())


class KernelIntegral:
    """a cell integral the sm_100a kernels implement, BY NAME: `kind` in {"fluid", "solid", "laplace", "biharmonic",
    "elastic"}, `part` in {"linear", "nonlinear"} (the reference's F_*_linear / F_*_nonlinear split,
    newtonsolver.py:18-24), `marker` = the dx sub-domain id, `params` = the constants baked into the kernel."""

    def __init__(self, kind, part, marker, scale=1.0, **params):
        self.kind, self.part, self.marker, self.scale, self.params = kind, part, int(marker), float(scale), params

    def scaled(self, s):
        return KernelIntegral(self.kind, self.part, self.marker, self.scale * s, **self.params)

    def __repr__(self):
        return "%s%s[%s]*dx(%d)" % ("-" if self.scale < 0 else "", self.kind, self.part, self.marker)


class Form:
    """a sum of integrals (Integral / KernelIntegral). `Form() + x`, `0 + form`, `-form`, `a * form` as in UFL."""

    def __init__(self, integrals=()):
        self.integrals = list(integrals)

    def __add__(self, o):
        if isinstance(o, (int, float)) and o == 0:
            return self
        if isinstance(o, (Integral, KernelIntegral)):
            return Form(self.integrals + [o])
        if not isinstance(o, Form):
            return NotImplemented
        return Form(self.integrals + o.integrals)

    __radd__ = __add__

    def __neg__(self):
        return Form([i.scaled(-1.0) for i in self.integrals])

    def __sub__(self, o):
        return self + (-o)

    def __rmul__(self, s):
        return Form([i.scaled(float(s)) for i in self.integrals])

    __mul__ = __rmul__

    def kernel_terms(self):
        return [i for i in self.integrals if isinstance(i, KernelIntegral)]

    def expression_terms(self):
        return [i for i in self.integrals if isinstance(i, Integral)]

    def empty(self):
        return not self.integrals

    def __repr__(self):
        return "Form(" + " + ".join(repr(i) if isinstance(i, KernelIntegral) else "<%s integral>" % i.measure.kind
                                    for i in self.integrals) + ")"


def regions_of(form):
    """(regions, switches) of the assembler from the kernel terms that are PRESENT in the form: one region entry per
    fluid / solid marker with its constants (turtleFSI/modules/domain.py:54-103 semantics) and the switches
    `fluid`, `solid`, `extrapolation`, `extrapolation_sub_type` in the reference's vocabulary. A lifting term over a
    marker without fluid term (structure + mesh motion only) still makes that marker a fluid region."""
    regions, lifting, kinds = {}, None, set()
    for t in form.kernel_terms():
        if t.scale != 1.0:
            raise NotImplementedError("kernel integral %r scaled by %g: the kernels assemble the reference's forms with "
                                      "unit weight" % (t, t.scale))
        if t.kind in ("fluid", "solid"):
            kinds.add(t.kind)
            r = regions.setdefault(t.marker, dict(marker=t.marker, kind=t.kind))
            if r["kind"] != t.kind:
                raise ValueError("cell marker %d carries fluid and solid terms" % t.marker)
            for k, v in t.params.items():
                if k in r and r[k] != v:
                    raise ValueError("region %d: conflicting %s (%r vs %r)" % (t.marker, k, r[k], v))
                r[k] = v
        elif t.kind in ("laplace", "biharmonic", "elastic"):
            sub = t.params.get("extrapolation_sub_type", "constant")
            if lifting is not None and lifting[:2] != (t.kind, sub):
                raise ValueError("two different lifting operators in one form: %r and %r" % (lifting[:2], (t.kind, sub)))
            lifting = (t.kind, sub, (lifting[2] if lifting else ()) + (t.marker,))
        else:
            raise NotImplementedError("no kernel implements the cell integral %r" % t.kind)
    if lifting is not None:
        for m in lifting[2]:
            r = regions.setdefault(m, dict(marker=m, kind="fluid"))
            if r["kind"] != "fluid":
                raise ValueError("mesh lifting over the solid region %d" % m)
    fluid_markers = [m for m, r in regions.items() if r["kind"] == "fluid"]
    if "fluid" in kinds and any("rho" not in regions[m] for m in fluid_markers):
        raise NotImplementedError("fluid terms on some fluid regions only: the kernels switch the fluid on or off globally")
    for r in regions.values():
        r.setdefault("rho", 0.0)
        r.setdefault("mu", 0.0)
    switches = dict(fluid="fluid" if "fluid" in kinds else "no_fluid", solid="solid" if "solid" in kinds else "no_solid",
                    extrapolation=lifting[0] if lifting else "no_extrapolation",
                    extrapolation_sub_type=lifting[1] if lifting else "constant")
    return list(regions.values()), switches


# ------------------------------------------------------------------------------------------------------
# quadrature point sets
class PointSet:
    """a batch of m entities (facets or cells) with nq quadrature points each."""

    def __init__(self, mesh, space, cell, lfacet, xi_by_k, x, w, normal):
        self.mesh, self.space = mesh, space
        self.cell, self.lfacet, self._xi_by_k = cell, lfacet, xi_by_k       # dicts side -> [m]
        self.x, self.w, self.normal = x, w, normal
        self.m, self.nq = w.shape
        self.arg = None
        self._tab, self._K = {}, {}

    def tab(self, side):
        """(phi2 [m, nq, n2], dphi2 [m, nq, n2, g] REFERENCE gradients, phi1 [m, nq, n1]) in the side's cell."""
        if side not in self._tab:
            g = self.mesh.gdim
            t = [tabulate(g, xi) for xi in self._xi_by_k]
            k = self.lfacet[side]
            self._tab[side] = (np.stack([a[0] for a in t])[k], np.stack([a[1] for a in t])[k],
                               np.stack([a[2] for a in t])[k])
        return self._tab[side]

    def K(self, side):
        if side not in self._K:
            X = self.mesh.coords[self.mesh.cells[self.cell[side]]]
            self._K[side] = np.linalg.inv(np.swapaxes(X[:, 1:] - X[:, :1], 1, 2))      # K[m, k, b] = d xi_k / d x_b
        return self._K[side]

    def grad_phi2(self, side):
        return np.einsum('mqnk,mkb->mqnb', self.tab(side)[1], self.K(side))

    def grad_phi1(self, side):
        g = self.mesh.gdim
        dphi1 = np.vstack([-np.ones(g), np.eye(g)])                                      # P1 reference gradients
        return np.broadcast_to(np.einsum('nk,mkb->mnb', dphi1, self.K(side))[:, None], (self.m, self.nq, g + 1, g))


def _facet_rule(g, degree):
    from .boundary import _facet_rule as rule
    return rule(g, degree)


def facet_point_sets(mesh, space, facets, fids, degree, plus="lower_index", chunk=8192):
    """PointSets over the facets `fids`. '+' side: "lower_index" = the incident cell with the lower index (DOLFIN's
    first incident cell; pinned by the TF_fsi drag/lift goldens), "higher_marker" = the incident cell with the higher
    cell marker (forms that carry cell sub-domains, boundary.interior_facet_normal_load)."""
    g = mesh.gdim
    bary, fw = _facet_rule(g, degree)
    ref_vertices = np.vstack([np.zeros(g), np.eye(g)])
    xi_by_k = [bary @ ref_vertices[[v for v in range(g + 1) if v != k]] for k in range(g + 1)]
    fids = np.asarray(fids, dtype=np.int64)
    for s in range(0, len(fids), chunk):
        f = fids[s:s + chunk]
        c0, k0, c1, k1 = facets.cell0[f], facets.local0[f], facets.cell1[f].copy(), facets.local1[f].copy()
        interior = c1 >= 0
        if plus == "higher_marker" and mesh.markers is not None:
            swap = interior & (mesh.markers[np.maximum(c1, 0)] > mesh.markers[c0])
            c0, c1 = np.where(swap, c1, c0), np.where(swap, c0, c1)
            k0, k1 = np.where(swap, k1, k0), np.where(swap, k0, k1)
        c1 = np.where(interior, c1, c0)               # exterior facets: '-' falls back to the only cell
        k1 = np.where(interior, k1, k0)
        V = mesh.coords[facets.vertices[f]]                                  # [m, g, g] ascending vertex ids
        x = np.einsum('qv,mvb->mqb', bary, V)
        if g == 2:
            t = V[:, 1] - V[:, 0]
            meas = np.linalg.norm(t, axis=1)
            nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / meas[:, None]
        else:
            cr = np.cross(V[:, 1] - V[:, 0], V[:, 2] - V[:, 0])
            nl = np.linalg.norm(cr, axis=1)
            meas = 0.5 * nl
            nrm = cr / nl[:, None]
        opp = mesh.coords[mesh.cells[c0, k0]]                                # vertex of the '+' cell off the facet
        flip = np.einsum('mb,mb->m', nrm, opp - V[:, 0]) > 0
        nrm = np.where(flip[:, None], -nrm, nrm)
        yield PointSet(mesh, space, {"+": c0, "-": c1}, {"+": k0, "-": k1}, xi_by_k, x, meas[:, None] * fw[None, :],
                       {"+": nrm, "-": -nrm})


def cell_point_sets(mesh, space, cids, degree, chunk=8192):
    from .quadrature import default_rule
    g = mesh.gdim
    P, W = default_rule(g, degree)
    cids = np.asarray(cids, dtype=np.int64)
    for s in range(0, len(cids), chunk):
        c = cids[s:s + chunk]
        X = mesh.coords[mesh.cells[c]]
        Jg = np.swapaxes(X[:, 1:] - X[:, :1], 1, 2)
        detJ = np.abs(np.linalg.det(Jg))
        x = X[:, None, 0, :] + np.einsum('qk,mbk->mqb', P, Jg)
        z = np.zeros(len(c), dtype=np.int64)
        yield PointSet(mesh, space, {"+": c, "-": c}, {"+": z, "-": z}, [P], x, detJ[:, None] * W[None, :], None)


def _entities(integral, mesh, facets):
    """ids of the facets / cells an integral runs over."""
    ms = integral.measure
    data = ms.subdomain_data
    sid = ms.subdomain_id
    ids = None if sid is None or sid == "everywhere" else (list(sid) if isinstance(sid, (tuple, list)) else [sid])
    if ms.kind == "dx":
        markers = np.asarray(data.array() if hasattr(data, "array") else (data if data is not None else mesh.markers))
        sel = np.ones(mesh.num_cells, dtype=bool) if ids is None else np.isin(markers, ids)
        return np.nonzero(sel)[0]
    markers = np.asarray(data.array() if hasattr(data, "array") else (data if data is not None else facets.markers))
    sel = facets.on_boundary if ms.kind == "ds" else ~facets.on_boundary
    if ids is not None:
        sel = sel & np.isin(markers, ids)
    return np.nonzero(sel)[0]


def _degree(integral, default_degree):
    md = integral.measure.metadata or {}
    if md.get("quadrature_degree") is not None:
        return int(md["quadrature_degree"])
    if default_degree is not None and int(default_degree) >= 0:
        return int(default_degree)
    return max(int(integral.integrand.degree()), 1)


def assemble_scalar(form, mesh, space, facets, default_degree=None, plus="lower_index"):
    """value of a functional (no test function): sum over its integrals."""
    total = 0.0
    for it in form.expression_terms():
        if it.integrand.rank:
            raise ValueError("integrand of a functional must be scalar, got shape %s" % (it.integrand.shape,))
        if it.test_function() is not None:
            raise ValueError("assemble_scalar: the form has a test function")
        ids = _entities(it, mesh, facets)
        deg = _degree(it, default_degree)
        sets = cell_point_sets(mesh, space, ids, deg) if it.measure.kind == "dx" else \
            facet_point_sets(mesh, space, facets, ids, deg, plus)
        for ps in sets:
            total += it.scale * float((it.integrand.evaluate(ps, "+") * ps.w).sum())
    return total


def assemble_load(integrals, mesh, space, facets, default_degree=None, plus="higher_marker"):
    """vector [num_dofs] of integrals that are linear in ONE test function and do not depend on the unknown."""
    b = np.zeros(space.num_dofs)
    g, n2 = space.gdim, space.n2
    off = {0: 0, 1: g * n2, 2: 2 * g * n2, 3: 2 * g * n2 + g + 1}
    for it in integrals:
        if it.integrand.rank:
            raise ValueError("integrand must be scalar, got shape %s" % (it.integrand.shape,))
        if it.test_function() is None:
            raise ValueError("assemble_load: integral without a test function")
        ids = _entities(it, mesh, facets)
        deg = _degree(it, default_degree)
        sets = cell_point_sets(mesh, space, ids, deg) if it.measure.kind == "dx" else \
            facet_point_sets(mesh, space, facets, ids, deg, plus)
        for ps in sets:
            val = it.integrand.evaluate(ps, "+")                              # [nb, m, nq]
            if ps.arg is None or val.ndim != 3:
                raise ValueError("the integrand is not linear in a test function")
            field, side = ps.arg
            nb = val.shape[0]
            loc = (val * ps.w[None]).sum(axis=2) * it.scale                   # [nb, m]
            dofs = space.cell_dofs[ps.cell[side]][:, off[field]:off[field] + nb]   # [m, nb]
            np.add.at(b, dofs.T.ravel(), loc.ravel())
    return b
