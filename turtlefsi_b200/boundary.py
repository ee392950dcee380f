"""Host-side boundary logic of the drop-in: facet markers, Dirichlet dof sets, surface
functionals and point evaluation (2-D and 3-D simplices, numpy).

Replaces what the problem files obtain from DOLFIN [dep]:
  AutoSubDomain(...).mark(boundaries, id), DomainBoundary()   turtleFSI/problems/TF_fsi.py:70-95
  DirichletBC(DVP.sub(i)[.sub(j)], value, boundaries, id)      TF_fsi.py:143-185 (topological)
  assemble((sigma(v,p,d,mu_f)*n)[k]*ds(id)), ... ("+")*dS(id)  TF_fsi.py:229-233
  Function.__call__(point) / peval                              TF_fsi.py:208-219
Marking rule of SubDomain.mark: a facet is marked when all its vertices AND its midpoint satisfy
inside(x, on_boundary). "+" side of an interior facet = its lowest-numbered incident cell.
"""
import itertools

import numpy as np

from .element import tabulate


class Facets:
    """codimension-1 entities with their incident cells."""

    def __init__(self, mesh):
        g = mesh.gdim
        self.mesh = mesh
        cells = mesh.cells.astype(np.int64)
        nc, nv = cells.shape[0], mesh.num_vertices
        # local facet k is opposite to local vertex k (UFC)
        loc = [tuple(v for v in range(g + 1) if v != k) for k in range(g + 1)]
        fv = np.stack([cells[:, l] for l in loc], axis=1)            # [nc, g+1, g] ascending already
        key = np.zeros((nc, g + 1), dtype=np.int64)
        for c in range(g):
            key = key * nv + fv[:, :, c]
        ukey, inv = np.unique(key.ravel(), return_inverse=True)
        self.num = len(ukey)
        verts = np.zeros((self.num, g), dtype=np.int64)
        rem = ukey.copy()
        for c in range(g - 1, -1, -1):
            verts[:, c] = rem % nv
            rem //= nv
        self.vertices = verts                                        # [nf, g]
        self.cell_facets = inv.reshape(nc, g + 1)
        # incident cells, ascending cell index
        order = np.argsort(inv, kind="stable")
        fsorted = inv[order]
        cell_of = order // (g + 1)
        localf = order % (g + 1)
        start = np.searchsorted(fsorted, np.arange(self.num))
        cnt = np.diff(np.append(start, len(fsorted)))
        self.cell0 = cell_of[start]
        self.local0 = localf[start]
        self.cell1 = np.where(cnt > 1, cell_of[np.minimum(start + 1, len(order) - 1)], -1)
        self.local1 = np.where(cnt > 1, localf[np.minimum(start + 1, len(order) - 1)], -1)
        self.on_boundary = cnt == 1
        X = mesh.coords
        self.midpoints = X[verts].mean(axis=1)
        self.markers = np.zeros(self.num, dtype=np.int64)

    def mark(self, inside, value):
        """inside(x [n, g], on_boundary [n]) -> bool [n]; DOLFIN SubDomain.mark semantics."""
        X = self.mesh.coords
        ok = inside(self.midpoints, self.on_boundary)
        for c in range(self.vertices.shape[1]):
            ok &= inside(X[self.vertices[:, c]], self.on_boundary)
        self.markers[ok] = value


def mark_cells(mesh, inside, value):
    X = mesh.coords
    ok = inside(X[mesh.cells].mean(axis=1))
    for c in range(mesh.cells.shape[1]):
        ok &= inside(X[mesh.cells[:, c]])
    mesh.markers[ok] = value


class DirichletBC:
    """DirichletBC(DVP.sub(field)[.sub(comp)], value, boundaries, marker), topological method.
    value: scalar, tuple (one entry per component) or callable(x [n, g]) -> [n, ncomp]."""

    def __init__(self, space, mesh, facets, field, value, marker, comp=None):
        self.field, self.comp, self.value = field, comp, value
        g = mesh.gdim
        fsel = np.nonzero(facets.markers == marker)[0]
        verts = np.unique(facets.vertices[fsel])
        if field == 2:
            self.dofs = [np.asarray(space.vertex_dof_p())[verts]]
            self.coords = mesh.coords[verts]
            self.comps = [0]
            return
        nodes = [space.vertex_node[verts]]
        if len(fsel):
            # edge midpoints lying on the marked facets
            fv = facets.vertices[fsel]
            pairs = set()
            for a, b in itertools.combinations(range(g), 2):
                for lo, hi in zip(np.minimum(fv[:, a], fv[:, b]), np.maximum(fv[:, a], fv[:, b])):
                    pairs.add((int(lo), int(hi)))
            ek = space.edge_lookup()
            eids = np.array(sorted(ek[p] for p in pairs), dtype=np.int64)
            nodes.append(space.edge_node[eids])
        nodes = np.concatenate(nodes)
        self.comps = list(range(g)) if comp is None else [comp]
        self.dofs = [space.node_dof(field, a)[nodes] for a in self.comps]
        self.coords = space.node_coords(mesh)[nodes]

    def values(self):
        n = len(self.coords)
        if callable(self.value):
            val = np.asarray(self.value(self.coords)).reshape(n, -1)
            return [val[:, a] for a in (self.comps if val.shape[1] > 1 else [0] * len(self.comps))]
        v = np.atleast_1d(np.asarray(self.value, dtype=float))
        if self.field == 2 or self.comp is not None:
            return [np.full(n, v[0])]
        return [np.full(n, v[a]) for a in self.comps]


class PointwiseBC:
    """DirichletBC(DVP.sub(field), value, inside, method="pointwise") [dep]: every dof of the sub-space
    whose coordinate satisfies inside(x [n, g]) -> bool [n] (problems/TaylorGreen2D.py:157-178: the
    pressure is pinned at one corner). Same dofs/values() protocol as DirichletBC."""

    def __init__(self, space, mesh, field, value, inside, comp=None):
        self.field, self.comp, self.value = field, comp, value
        g = mesh.gdim
        if field == 2:
            sel = np.nonzero(inside(mesh.coords))[0]
            self.dofs = [np.asarray(space.vertex_dof_p())[sel]]
            self.coords = mesh.coords[sel]
            self.comps = [0]
            return
        X = space.node_coords(mesh)
        sel = np.nonzero(inside(X))[0]
        self.comps = list(range(g)) if comp is None else [comp]
        self.dofs = [space.node_dof(field, a)[sel] for a in self.comps]
        self.coords = X[sel]

    values = DirichletBC.values


def interpolate(space, mesh, x, field, fn):
    """x[dofs of sub-space `field`] = fn(dof coordinates) — DOLFIN interpolate + assign of a Lagrange
    field (problems/TaylorGreen2D.py:186-198). fn(X [n, g]) -> [n, ncomp] (or [n] for p)."""
    if field == 2:
        x[np.asarray(space.vertex_dof_p())] = np.asarray(fn(mesh.coords)).reshape(-1)
        return x
    val = np.asarray(fn(space.node_coords(mesh))).reshape(space.num_nodes, -1)
    for a in range(mesh.gdim):
        x[space.node_dof(field, a)] = val[:, a]
    return x


def l2_norm(mesh, space, x, field):
    """sqrt(int |u_h|^2 dx) of the Lagrange field `field` of x, exactly (element mass matrices) — what
    DOLFIN's errornorm(ue, uh, "L2") returns when both arguments live in the same space
    (problems/TaylorGreen2D.py:216-221)."""
    from .quadrature import default_rule
    g = mesh.gdim
    P, W = default_rule(g, 4)
    phi2, _, phi1 = tabulate(g, P)
    X = mesh.coords[mesh.cells]
    detJ = np.abs(np.linalg.det(np.swapaxes(X[:, 1:] - X[:, :1], 1, 2)))
    n2 = space.n2
    if field == 2:
        M = np.einsum('q,qi,qj->ij', W, phi1, phi1)
        u = x[space.cell_dofs[:, 2 * g * n2:2 * g * n2 + g + 1]]
        return float(np.sqrt(np.einsum('c,ci,ij,cj->', detJ, u, M, u)))
    M = np.einsum('q,qi,qj->ij', W, phi2, phi2)
    off = {0: 0, 1: g * n2, 3: 2 * g * n2 + g + 1}[field]
    total = 0.0
    for a in range(g):
        u = x[space.cell_dofs[:, off + a * n2: off + (a + 1) * n2]]
        total += np.einsum('c,ci,ij,cj->', detJ, u, M, u)
    return float(np.sqrt(total))


class BCPlan:
    """index plan of a BC list, built once: sorted unique dofs and, for each, the position of the LAST BC
    entry that sets it (DOLFIN applies the list sequentially, later BCs win: newtonsolver.py:82,91,98).
    Per time step only values() of the BC objects is evaluated."""

    def __init__(self, bcs):
        self.bcs = list(bcs)
        cat = [np.asarray(d, dtype=np.int64) for bc in self.bcs for d in bc.dofs]
        self.sizes = [len(d) for d in cat]
        alld = np.concatenate(cat) if cat else np.zeros(0, dtype=np.int64)
        # last occurrence of every dof: first occurrence in the reversed list
        udofs, first_rev = np.unique(alld[::-1], return_index=True)
        self.dofs = udofs.astype(np.int32)
        self.pick = (len(alld) - 1 - first_rev).astype(np.int64)

    def matches(self, bcs):
        """True when `bcs` has the same objects-with-same-dof-arrays structure (cheap check per step)."""
        bcs = list(bcs)
        if len(bcs) != len(self.bcs):
            return False
        sizes = [len(d) for bc in bcs for d in bc.dofs]
        if sizes != self.sizes:
            return False
        return all(a is b or all(np.array_equal(x, y) for x, y in zip(a.dofs, b.dofs)) for a, b in zip(bcs, self.bcs))

    def values(self, bcs=None):
        bcs = self.bcs if bcs is None else bcs
        cat = [np.asarray(v, dtype=np.float64) for bc in bcs for v in bc.values()]
        allv = np.concatenate(cat) if cat else np.zeros(0)
        return allv[self.pick]


def flatten_bcs(bcs):
    """(dofs int32 [n] ascending, values [n]) in list order, later BCs win on shared dofs (DOLFIN applies
    the list sequentially: newtonsolver.py:82,91,98)."""
    plan = BCPlan(bcs)
    return plan.dofs, plan.values()


def _gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def surface_traction(mesh, space, facets, x, mu_f, facet_ids, degree=10):
    """sum over the given facets of  int sigma(v, p, d, mu_f) * n  dS  (vector [g]), evaluated on
    the "+" cell (cell0) with its outward normal; sigma = mu (grad v F^-1 + F^-T grad v^T) - p I
    (turtleFSI/modules/common.py:63-81; the drag/lift functionals of TF_fsi.py:229-233, TF_cfd.py:129-131).
    Facet rule: Gauss-Legendre on edges, the FIAT triangle rule on faces, exact to `degree`."""
    g = mesh.gdim
    bary, fw = _facet_rule(g, degree)
    total = np.zeros(g)
    ref_vertices = np.vstack([np.zeros(g), np.eye(g)])
    n2 = space.n2
    for f in facet_ids:
        c = facets.cell0[f]
        k = facets.local0[f]
        vloc, meas, nrm = _facet_frame(mesh, c, k)
        X = mesh.coords[mesh.cells[c]]
        K = np.linalg.inv((X[1:] - X[0]).T)
        phi2, dphi2, phi1 = tabulate(g, bary @ ref_vertices[vloc])
        u = x[space.cell_dofs[c]]
        dco = u[:g * n2].reshape(g, n2)
        vco = u[g * n2:2 * g * n2].reshape(g, n2)
        pco = u[2 * g * n2:2 * g * n2 + g + 1]
        gphys = np.einsum('qik,kb->qib', dphi2, K)
        Gd = np.einsum('ai,qib->qab', dco, gphys)
        Gv = np.einsum('ai,qib->qab', vco, gphys)
        p = phi1 @ pco
        F = np.eye(g) + Gd
        Fi = np.linalg.inv(F)
        H = Gv @ Fi
        sig = mu_f * (H + np.swapaxes(H, 1, 2)) - p[:, None, None] * np.eye(g)
        total += np.einsum('q,qab,b->a', fw * meas, sig, nrm)
    return total


def biharmonic_facet_operator(mesh, space, facets, bc_ids, alpha_u=0.01):
    """COO triplets (rows, cols, vals) of the exterior-facet terms of the biharmonic lifting with
    `extrapolation_sub_type="constrained_disp_vel"` (turtleFSI/modules/biharmonic.py:38-41):
        F_ext1 += alpha_u (grad(d) n) . beta  ds(i)        rows: w-test (beta), cols: d
        F_ext2 -= alpha_u (grad(w) n) . phi   ds(i)        rows: d-test (phi),  cols: w
    for i in bc_ids. Both are LINEAR, so they are a fixed sparse operator A_f: it is added to A_pre
    and b -= A_f x is added to the residual (tfsi_set_facet_operator). Facet quadrature:
    Gauss-Legendre (2-D) / the degree-4 triangle rule (3-D): exact for the P2 x P1 integrand."""
    from .quadrature import default_rule
    g = mesh.gdim
    n2 = space.n2
    ow = 2 * g * n2 + g + 1
    if g == 2:
        s, w = _gauss_legendre_01(3)
        fpts = np.stack([1.0 - s, s], axis=1)                 # barycentric on the facet
        fw = w
    else:
        P, W = default_rule(2, 4)
        fpts = np.concatenate([1.0 - P.sum(axis=1, keepdims=True), P], axis=1)
        fw = W * 2.0                                          # weights sum to 1 (fraction of the facet area)
    ref_vertices = np.vstack([np.zeros(g), np.eye(g)])
    rows, cols, vals = [], [], []
    sel = np.nonzero(np.isin(facets.markers, list(bc_ids)) & facets.on_boundary)[0]
    for f in sel:
        c, k = facets.cell0[f], facets.local0[f]
        vloc = [v for v in range(g + 1) if v != k]
        X = mesh.coords[mesh.cells[c]]
        K = np.linalg.inv((X[1:] - X[0]).T)
        if g == 2:
            t = X[vloc[1]] - X[vloc[0]]
            meas = np.linalg.norm(t)
            nrm = np.array([t[1], -t[0]]) / meas
        else:
            cr = np.cross(X[vloc[1]] - X[vloc[0]], X[vloc[2]] - X[vloc[0]])
            meas = 0.5 * np.linalg.norm(cr)
            nrm = cr / np.linalg.norm(cr)
        if np.dot(nrm, X[k] - X[vloc[0]]) > 0:
            nrm = -nrm
        pts = fpts @ ref_vertices[vloc]
        phi2, dphi2, _ = tabulate(g, pts)
        dn = np.einsum('qik,kb,b->qi', dphi2, K, nrm)          # d phi_j / d n
        M = alpha_u * meas * np.einsum('q,qi,qj->ij', fw, phi2, dn)   # M[i, j] = int phi_i dn(phi_j)
        dofs = space.cell_dofs[c]
        for a in range(g):
            wd = dofs[ow + a * n2: ow + (a + 1) * n2]
            dd = dofs[a * n2:(a + 1) * n2]
            rows += [np.repeat(wd, n2), np.repeat(dd, n2)]
            cols += [np.tile(dd, n2), np.tile(wd, n2)]
            vals += [M.ravel(), -M.ravel()]
    if not rows:
        z = np.zeros(0)
        return z.astype(np.int32), z.astype(np.int32), z
    return (np.concatenate(rows).astype(np.int32), np.concatenate(cols).astype(np.int32), np.concatenate(vals))


def _facet_rule(g, degree):
    """(barycentric points on the facet [nq, g], weights summing to 1) exact to `degree`."""
    from .quadrature import default_rule
    if g == 2:
        s, w = _gauss_legendre_01(degree // 2 + 1)
        return np.stack([1.0 - s, s], axis=1), w
    P, W = default_rule(2, degree)
    return np.concatenate([1.0 - P.sum(axis=1, keepdims=True), P], axis=1), W * 2.0


def _facet_frame(mesh, c, k):
    """(local vertex ids of the facet opposite to local vertex k of cell c, measure, outward unit normal)."""
    g = mesh.gdim
    vloc = [v for v in range(g + 1) if v != k]
    X = mesh.coords[mesh.cells[c]]
    if g == 2:
        t = X[vloc[1]] - X[vloc[0]]
        meas = np.linalg.norm(t)
        nrm = np.array([t[1], -t[0]]) / meas
    else:
        cr = np.cross(X[vloc[1]] - X[vloc[0]], X[vloc[2]] - X[vloc[0]])
        meas = 0.5 * np.linalg.norm(cr)
        nrm = cr / np.linalg.norm(cr)
    if np.dot(nrm, X[k] - X[vloc[0]]) > 0:
        nrm = -nrm
    return vloc, meas, nrm


def robin_facet_operators(mesh, space, facets, ds_s_ids, k_s, c_s, theta):
    """Robin condition of the solid (turtleFSI/modules/solid.py:76-86), for i over ds_s_ids:
        F_solid_linear += theta0 (k_s[i] d^n     + c_s[i] v^n    ) . psi ds(ds_s_ids[i])      :83
        F_solid_linear += theta1 (k_s[i] d^{n-1} + c_s[i] v^{n-1}) . psi ds(ds_s_ids[i])      :84
    Linear in the state: returns two COO operators (rows psi = v-test dofs; cols d and v),
    `now` (-> tfsi_set_facet_operator: A_pre += , b -= now x) and `prev` (-> tfsi_set_facet_operator_prev:
    b -= prev x1). The P2 x P2 facet mass matrix is integrated exactly (degree 4)."""
    g = mesh.gdim
    n2 = space.n2
    bary, fw = _facet_rule(g, 4)
    ref_vertices = np.vstack([np.zeros(g), np.eye(g)])
    th0, th1 = theta, 1.0 - theta
    out = {"now": ([], [], []), "prev": ([], [], [])}
    for sid, ks, cs in zip(ds_s_ids, k_s, c_s):
        for f in np.nonzero((facets.markers == sid) & facets.on_boundary)[0]:
            c, k = facets.cell0[f], facets.local0[f]
            vloc, meas, _ = _facet_frame(mesh, c, k)
            phi2, _, _ = tabulate(g, bary @ ref_vertices[vloc])
            M = meas * np.einsum('q,qi,qj->ij', fw, phi2, phi2)
            dofs = space.cell_dofs[c]
            for a in range(g):
                dd = dofs[a * n2:(a + 1) * n2]
                vd = dofs[(g + a) * n2:(g + a + 1) * n2]
                for key, th in (("now", th0), ("prev", th1)):
                    r, cc, v = out[key]
                    r += [np.repeat(vd, n2), np.repeat(vd, n2)]
                    cc += [np.tile(dd, n2), np.tile(vd, n2)]
                    v += [th * ks * M.ravel(), th * cs * M.ravel()]

    def pack(t):
        if not t[0]:
            z = np.zeros(0)
            return z.astype(np.int32), z.astype(np.int32), z
        return np.concatenate(t[0]).astype(np.int32), np.concatenate(t[1]).astype(np.int32), np.concatenate(t[2])
    return pack(out["now"]), pack(out["prev"])


def interior_facet_normal_load(mesh, space, facets, marker):
    """l_i = int n('+') . psi_i('+') dS(marker) over the INTERIOR facets with the given marker: the
    load a problem file adds as  p(t) * inner(n('+'), psi('+')) * dS(fsi_id)  to F_solid_linear
    (problems/Test_Cylinder/problem_aneu.py:212-214, problems/Stenosis_2D.py:172-174).
    "+" side [dep, DOLFIN Assembler]: the form carries cell sub-domains (dx_f, dx_s), so "+" is the
    incident cell with the HIGHER cell marker (the solid in the shipped problems); equal markers ->
    the lower cell index. Returns (dofs int32, vals) with duplicates merged."""
    g = mesh.gdim
    n2 = space.n2
    bary, fw = _facet_rule(g, 2)
    ref_vertices = np.vstack([np.zeros(g), np.eye(g)])
    acc = {}
    for f in np.nonzero((facets.markers == marker) & ~facets.on_boundary)[0]:
        c, k = facets.cell0[f], facets.local0[f]
        c1 = facets.cell1[f]
        if mesh.markers is not None and mesh.markers[c1] > mesh.markers[c]:
            c, k = c1, facets.local1[f]
        vloc, meas, nrm = _facet_frame(mesh, c, k)
        phi2, _, _ = tabulate(g, bary @ ref_vertices[vloc])
        w = meas * (fw @ phi2)
        dofs = space.cell_dofs[c]
        for a in range(g):
            for i in np.nonzero(np.abs(w) > 0)[0]:
                d = int(dofs[(g + a) * n2 + i])
                acc[d] = acc.get(d, 0.0) + nrm[a] * w[i]
    d = np.fromiter(acc.keys(), dtype=np.int64, count=len(acc))
    v = np.fromiter(acc.values(), dtype=np.float64, count=len(acc))
    order = np.argsort(d)
    return d[order].astype(np.int32), v[order]


def point_eval(mesh, space, x, point, field=0):
    """value of P2 vector field `field` at a point (first containing cell, as DOLFIN's bounding-box
    tree returns one of the colliding cells; the field is continuous)."""
    g = mesh.gdim
    X = mesh.coords[mesh.cells]
    Jg = np.swapaxes(X[:, 1:] - X[:, :1], 1, 2)
    xi = np.linalg.solve(Jg, (np.asarray(point)[None, :] - X[:, 0])[:, :, None])[:, :, 0]
    lam = np.concatenate([1 - xi.sum(axis=1, keepdims=True), xi], axis=1)
    inside = np.nonzero((lam > -1e-10).all(axis=1))[0]
    if len(inside) == 0:
        raise RuntimeError("point outside the mesh")
    c = inside[0]
    phi2, _, _ = tabulate(g, xi[c][None, :])
    n2 = space.n2
    off = {0: 0, 1: g * n2, 3: 2 * g * n2 + g + 1}[field]
    co = x[space.cell_dofs[c]][off:off + g * n2].reshape(g, n2)
    return co @ phi2[0]
