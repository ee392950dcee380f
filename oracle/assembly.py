"""ORACLE (test infrastructure only — never imported by the product path).

Global assembly of the oracle cell forms (oracle/forms.py) into the fixed CSR pattern, plus the
host restatement of the matrix/vector operations of turtleFSI/modules/newtonsolver.py:
  A_pre = assemble(J_linear)            :26-27        -> assemble_matrix(part="lin")
  assemble(J_nonlinear, tensor=A)       :75-77        -> assemble_matrix(part="nl")
  A.ident_zeros()                       :81  [dep DOLFIN: rows with sum|a_ij| < DOLFIN_EPS]
  bc.apply(A) / (b, x) / (x)            :82, :91, :98 [dep DOLFIN DirichletBC]
  assemble(-F)                          :88           -> assemble_vector
"""
import numpy as np

from .forms import CellForms

DOLFIN_EPS = 3.0e-16


def region_params(common, region):
    """merge the scheme-level parameters with one region's properties."""
    prm = dict(common)
    if region["kind"] == "fluid":
        prm.update(rho_f=region["rho"], mu_f=region["mu"])
    else:
        prm.update(rho_s=region["rho"], mu_s=region["mu"], lambda_s=region.get("lambda_", 0.0))
        for key in ("material_model", "C01", "C10", "C11", "Jm", "viscoelasticity", "mu_visc_s", "lambda_visc_s"):
            if key in region:
                prm[key] = region[key]
    return prm


def _pattern_keys(rowptr, colidx, n):
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
    return rows * n + colidx.astype(np.int64)


def assemble_matrix(mesh, space, regions, common, x, x1, part, degree, pattern, chunk=20000):
    """values [nnz] aligned with pattern=(rowptr, colidx). part in {"lin", "nl", "all"}."""
    rowptr, colidx = pattern
    n = space.num_dofs
    keys = _pattern_keys(rowptr, colidx, n)
    vals = np.zeros(len(colidx))
    cf = CellForms(mesh.gdim, space.biharmonic, degree)
    for reg in regions:
        prm = region_params(common, reg)
        sel = np.nonzero(mesh.markers == reg["marker"])[0]
        for s in range(0, len(sel), chunk):
            c = sel[s:s + chunk]
            X = mesh.coords[mesh.cells[c]]
            cd = space.cell_dofs[c].astype(np.int64)
            Je = cf.jacobian(reg["kind"], X, x[cd], x1[cd], prm, part)
            ek = (cd[:, :, None] * n + cd[:, None, :]).ravel()
            pos = np.searchsorted(keys, ek)
            assert np.array_equal(keys[pos], ek), "cell block outside the pattern"
            np.add.at(vals, pos, Je.ravel())
    return vals


def assemble_vector(mesh, space, regions, common, x, x1, degree, chunk=50000):
    """b = assemble(-F), F = F_lin + F_nonlin."""
    b = np.zeros(space.num_dofs)
    cf = CellForms(mesh.gdim, space.biharmonic, degree)
    for reg in regions:
        prm = region_params(common, reg)
        sel = np.nonzero(mesh.markers == reg["marker"])[0]
        for s in range(0, len(sel), chunk):
            c = sel[s:s + chunk]
            X = mesh.coords[mesh.cells[c]]
            cd = space.cell_dofs[c].astype(np.int64)
            Rl, Rn = cf.residual(reg["kind"], X, x[cd], x1[cd], prm)
            np.add.at(b, cd.ravel(), -(Rl + Rn).ravel())
    return b


def ident_zeros(vals, pattern):
    rowptr, colidx = pattern
    n = len(rowptr) - 1
    rowsum = np.add.reduceat(np.abs(vals), rowptr[:-1])
    rowsum[np.diff(rowptr) == 0] = 0.0
    for r in np.nonzero(rowsum < DOLFIN_EPS)[0]:
        seg = colidx[rowptr[r]:rowptr[r + 1]]
        k = np.searchsorted(seg, r)
        assert seg[k] == r
        vals[rowptr[r] + k] = 1.0
    return vals


def apply_bc_matrix(vals, pattern, bc_dofs):
    rowptr, colidx = pattern
    for r in np.unique(bc_dofs):
        seg = slice(rowptr[r], rowptr[r + 1])
        vals[seg] = np.where(colidx[seg] == r, 1.0, 0.0)
    return vals


def apply_bc_rhs(b, x, bc_dofs, g):
    b[bc_dofs] = g - x[bc_dofs]
    return b


def apply_bc_vec(x, bc_dofs, g):
    x[bc_dofs] = g
    return x
