"""ORACLE (test infrastructure only): numpy assembler with the interface that
turtlefsi_b200/modules/newtonsolver.py drives — the CPU restatement of the whole Newton hot path
(forms: oracle/forms.py; assembly + BC + ident_zeros: oracle/assembly.py). Used by the tests to
run the reference's Newton loop on the CPU and compare the CUDA path against it."""
import numpy as np

from . import assembly as oa


class OracleAssembler:
    def __init__(self, mesh, space, regions, common, quad_degree_jacobian, quad_degree_residual, pattern=None,
                 use_c_port=False):
        """use_c_port: assemble with the plain-C restatement (oracle/c, laplace/constant only; it is
        checked against the numpy complex-step forms in tests/test_oracle_forms.py) — 100x faster."""
        self.use_c = use_c_port
        ext = common.get("extrapolation", "laplace")
        sub = common.get("extrapolation_sub_type", "constant")
        self._c_ext = {"no_extrapolation": 0, "laplace": 1, "biharmonic": 2}.get(ext, -1)
        if use_c_port and (self._c_ext < 0 or (ext == "laplace" and sub != "constant")):
            raise ValueError("the C port covers lifting none | laplace/constant | biharmonic only")
        self.mesh, self.space, self.regions, self.common = mesh, space, regions, dict(common)
        if use_c_port:
            # no_fluid / no_solid: the C port has no switch; drop the region instead (same zero forms)
            self.regions = [r for r in regions if not ((r["kind"] == "fluid" and common.get("fluid", "fluid") != "fluid" and self._c_ext == 0)
                                                        or (r["kind"] == "solid" and common.get("solid", "solid") != "solid"))]
        self.common.setdefault("hmin", mesh.hmin())
        self.degJ, self.degR = quad_degree_jacobian, quad_degree_residual
        self.pattern = pattern if pattern is not None else space.pattern()
        self.rowptr, self.colidx = self.pattern
        self.num_dofs, self.nnz = space.num_dofs, len(self.colidx)
        self.bc_dofs = np.zeros(0, dtype=np.int32)
        self.facet_op = None
        self.facet_op_prev = None
        self.facet_load, self.facet_load_scale = None, 1.0

    def set_facet_operator(self, rows, cols, vals):
        """fixed linear facet terms (COO); added to A_pre, and b -= A_f x (biharmonic.py:38-41)."""
        import scipy.sparse as sp
        n = self.num_dofs
        self.facet_op = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr() if len(vals) else None

    def set_facet_operator_prev(self, rows, cols, vals):
        """linear facet terms on the previous time level (solid.py:84): b -= A_p x1."""
        import scipy.sparse as sp
        n = self.num_dofs
        self.facet_op_prev = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr() if len(vals) else None

    def set_facet_load(self, dofs, vals, scale=1.0):
        """state-independent facet load (problem_aneu.py:214): b -= scale * l."""
        l = np.zeros(self.num_dofs)
        np.add.at(l, np.asarray(dofs, dtype=np.int64), vals)
        self.facet_load, self.facet_load_scale = (l if len(vals) else None), float(scale)

    def set_facet_load_scale(self, scale):
        self.facet_load_scale = float(scale)

    def _facet_vals(self):
        keys = oa._pattern_keys(self.rowptr, self.colidx, self.num_dofs)
        M = self.facet_op.tocoo()
        pos = np.searchsorted(keys, M.row.astype(np.int64) * self.num_dofs + M.col)
        v = np.zeros(self.nnz)
        np.add.at(v, pos, M.data)
        return v

    def new_vector(self):
        return np.zeros(self.num_dofs)

    def new_matrix(self):
        return np.zeros(self.nnz)

    def to_device(self, a):
        return np.array(a, dtype=np.float64)

    def to_host(self, a):
        return np.asarray(a)

    def copy(self, dst, src):
        dst[:] = src

    def axpy(self, x, alpha, y):
        x += alpha * y

    def _c_matrix(self, x, x1, part):
        from . import cport, fem_tables as ft
        v = np.zeros(self.nnz)
        cport.assemble_jacobian(self.mesh, self.space, self.regions, self.common["dt"], self.common["theta"],
                                np.ascontiguousarray(x), np.ascontiguousarray(x1),
                                ft.quadrature(self.mesh.gdim, self.degJ), self.pattern, v, part=part, ext=self._c_ext)
        return v

    def assemble_A_pre(self, x, x1, out=None):
        v = self._c_matrix(x, x1, "lin") if self.use_c else oa.assemble_matrix(self.mesh, self.space, self.regions, self.common, x, x1, "lin", self.degJ, self.pattern)
        if self.facet_op is not None:
            v = v + self._facet_vals()
        if out is not None:
            out[:] = v
            return out
        return v

    def assemble_jacobian(self, x, x1, A_pre=None, out=None):
        v = self._c_matrix(x, x1, "nl") if self.use_c else oa.assemble_matrix(self.mesh, self.space, self.regions, self.common, x, x1, "nl", self.degJ, self.pattern)
        if A_pre is not None:
            v = v + A_pre
        if out is not None:
            out[:] = v
            return out
        return v

    def assemble_residual(self, x, x1, out=None):
        if self.use_c:
            from . import cport, fem_tables as ft
            b = np.zeros(self.num_dofs)
            cport.assemble_residual(self.mesh, self.space, self.regions, self.common["dt"], self.common["theta"],
                                    np.ascontiguousarray(x), np.ascontiguousarray(x1),
                                    ft.quadrature(self.mesh.gdim, self.degR), b, ext=self._c_ext,
                                    gravity=float(self.common.get("gravity") or 0.0))
        else:
            b = oa.assemble_vector(self.mesh, self.space, self.regions, self.common, x, x1, self.degR)
        if self.facet_op is not None:
            b = b - self.facet_op @ np.asarray(x)
        if self.facet_op_prev is not None:
            b = b - self.facet_op_prev @ np.asarray(x1)
        if self.facet_load is not None:
            b = b - self.facet_load_scale * self.facet_load
        if out is not None:
            out[:] = b
            return out
        return b

    def ident_zeros(self, A):
        oa.ident_zeros(A, self.pattern)

    def set_bc_dofs(self, dofs):
        self.bc_dofs = np.asarray(dofs, dtype=np.int64)

    def apply_bc_matrix(self, A):
        oa.apply_bc_matrix(A, self.pattern, self.bc_dofs)

    def apply_bc_rhs(self, b, x, g):
        oa.apply_bc_rhs(b, x, self.bc_dofs, g)

    def apply_bc_vec(self, x, g):
        oa.apply_bc_vec(x, self.bc_dofs, g)

    def norm(self, v):
        return float(np.linalg.norm(v))
