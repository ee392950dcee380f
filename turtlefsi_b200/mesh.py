"""Host-side mesh containers and generators (numpy only).

The reference obtains meshes from DOLFIN (`Mesh(xml.gz)`, `BoxMesh`, XDMF/HDF5 readers —
turtleFSI/problems/TF_fsi.py:68, problems/Test_Cylinder/problem_aneu.py:68-78). DOLFIN is not
available, so the hot path takes plain arrays; this module produces them:

* `read_dolfin_xml`  — the shipped `turtleFSI/mesh/TF_*.xml.gz` format
* `rectangle_mesh`   — DOLFIN `RectangleMesh(p0, p1, nx, ny, diagonal)` (problems/TaylorGreen2D.py:123)
* `box_mesh`         — N^3 cubes x 6 tets sharing the main diagonal (DOLFIN BoxMesh pattern),
                       jittered interior vertices (SURVEY 8d "Synthetic 3-D inputs")
* `tube_mesh`        — the same box mapped to a cylinder (R_o = 1, length 4), marker 1 (fluid)
                       for centroid radius < 0.8 else 2 (solid); mirrors problem_aneu.py:49-50

Cells always list their vertices in ascending global index (DOLFIN "ordered" mesh), which is
what fixes the local P2 edge numbering.
"""
import gzip
import re
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    coords: np.ndarray              # float64 [nv, g]
    cells: np.ndarray               # int32   [nc, g+1], ascending per row
    markers: np.ndarray = None      # int32   [nc] cell markers ("domains")
    meta: dict = field(default_factory=dict)

    @property
    def gdim(self):
        return self.coords.shape[1]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    def hmin(self):
        """DOLFIN Mesh.hmin(): min over cells of the greatest vertex-vertex distance."""
        X = self.coords[self.cells]
        g = self.gdim
        h = np.zeros(self.num_cells)
        for i in range(g + 1):
            for j in range(i + 1, g + 1):
                h = np.maximum(h, np.linalg.norm(X[:, i] - X[:, j], axis=1))
        return float(h.min())


def read_dolfin_xml(path):
    """DOLFIN XML (optionally .gz) with <vertex index x y [z]/> and <triangle|tetrahedron .../>."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        txt = f.read()
    dim = int(re.search(r'dim="(\d)"', txt).group(1))
    vpat = r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)"' + (r' z="([^"]+)"' if dim == 3 else "")
    verts = re.findall(vpat, txt)
    coords = np.zeros((len(verts), dim))
    for v in verts:
        coords[int(v[0])] = [float(s) for s in v[1:]]
    if dim == 2:
        cpat = r'<triangle index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)"'
    else:
        cpat = r'<tetrahedron index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)" v3="(\d+)"'
    cl = re.findall(cpat, txt)
    cells = np.zeros((len(cl), dim + 1), dtype=np.int32)
    for c in cl:
        cells[int(c[0])] = [int(s) for s in c[1:]]
    cells.sort(axis=1)
    return Mesh(coords, cells, np.zeros(len(cl), dtype=np.int32))


def read_turtle_demo(folder):
    """The default problem's mesh (turtleFSI/mesh/turtle_demo/{mc,mf}.h5, read by
    turtleFSI/problems/turtle_demo.py:58-87 through DOLFIN's XDMFFile [dep]). h5py is not available; the files are
    plain HDF5 (turtlefsi_b200.utils.hdf5_lite): mc.h5 = /data0 coordinates f8[3350,2], /data1 triangles i8[6578,3],
    /data2 cell markers i8[6578]; mf.h5 = /data1 marked facets i8[137,2], /data2 their markers i8[137].
    Returns (Mesh, facet_vertices [137, 2], facet_markers [137])."""
    import os
    from .utils.hdf5_lite import read_datasets
    mc = read_datasets(os.path.join(folder, "mc.h5"))
    mf = read_datasets(os.path.join(folder, "mf.h5"))
    coords = np.array(mc["/data0"], dtype=np.float64)
    cells = np.asarray(mc["/data1"]).astype(np.int32)
    markers = np.asarray(mc["/data2"]).astype(np.int32)
    fverts = np.asarray(mf["/data1"]).astype(np.int64)
    fmark = np.asarray(mf["/data2"]).astype(np.int64)
    if coords.ndim != 2 or cells.shape != (len(markers), coords.shape[1] + 1) or fverts.shape != (len(fmark), coords.shape[1]):
        raise RuntimeError("unexpected turtle_demo mesh files")
    if cells.min() != 0 or cells.max() != len(coords) - 1 or fverts.max() >= len(coords):
        raise RuntimeError("turtle_demo mesh files: vertex indices out of range")
    cells = np.sort(cells, axis=1)
    return Mesh(coords, cells, markers.copy(), meta=dict(kind="turtle_demo")), fverts, fmark


def refine(mesh):
    """Uniform refinement of a triangle mesh as DOLFIN's `refine(mesh)` does it (PlazaRefinementND,
    [dep] DOLFIN 2019.1, used by turtleFSI/problems/TF_cfd.py:49 and TF_csm.py:54): a triangle whose
    shortest/longest edge ratio is >= sqrt(2)/2 is split into 4 similar triangles ("red"); a more
    distorted one is bisected at the midpoint of its longest edge and each half again at the midpoint
    of its other edge. New vertices = edge midpoints. Children of cell c are cells 4c..4c+3."""
    if mesh.gdim != 2:
        raise NotImplementedError("refine: triangles only")
    cells = mesh.cells.astype(np.int64)
    nv = mesh.num_vertices
    X = mesh.coords
    # local edge k is opposite to local vertex k (UFC)
    loc = [(1, 2), (0, 2), (0, 1)]
    lo = np.stack([np.minimum(cells[:, a], cells[:, b]) for a, b in loc], axis=1)
    hi = np.stack([np.maximum(cells[:, a], cells[:, b]) for a, b in loc], axis=1)
    ukey, inv = np.unique((lo * nv + hi).ravel(), return_inverse=True)
    mid = nv + inv.reshape(-1, 3)                       # midpoint vertex of local edge k
    newX = np.vstack([X, 0.5 * (X[ukey // nv] + X[ukey % nv])])
    length = np.linalg.norm(X[hi] - X[lo], axis=2)     # [nc, 3]
    ok = length.min(axis=1) / length.max(axis=1) >= np.sqrt(2.0) / 2.0
    L = np.argmax(length, axis=1)                       # first longest edge
    out = np.zeros((len(cells), 4, 3), dtype=np.int64)
    v = cells
    e = mid
    # red refinement
    red = np.stack([np.stack([e[:, 0], e[:, 1], v[:, 2]], 1), np.stack([e[:, 1], e[:, 2], v[:, 0]], 1),
                    np.stack([e[:, 2], e[:, 0], v[:, 1]], 1), np.stack([e[:, 2], e[:, 1], e[:, 0]], 1)], axis=1)
    # bisection: v0, v1 at the ends of the longest edge (index L), v2 opposite
    i0, i1, i2 = (L + 1) % 3, (L + 2) % 3, L
    r = np.arange(len(cells))
    V0, V1, V2 = v[r, i0], v[r, i1], v[r, i2]
    E0, E1, E2 = e[r, i0], e[r, i1], e[r, i2]
    bis = np.stack([np.stack([E2, V2, E0], 1), np.stack([E2, E0, V1], 1),
                    np.stack([E2, V2, E1], 1), np.stack([E2, E1, V0], 1)], axis=1)
    out = np.where(ok[:, None, None], red, bis).reshape(-1, 3)
    out = np.sort(out, axis=1).astype(np.int32)
    markers = np.repeat(mesh.markers, 4) if mesh.markers is not None else None
    return Mesh(newX, out, markers, meta=dict(mesh.meta, refined=True))


def _morton3(ix, iy, iz):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return spread(ix) | (spread(iy) << np.uint64(1)) | (spread(iz) << np.uint64(2))


def box_mesh(n, jitter=0.1, seed=1234, locality=True):
    """[0,1]^3, n^3 cubes x 6 tets (6 n^3 cells), interior vertices moved by
    jitter*h*U(-1,1) (numpy default_rng(seed)). With locality=True vertices and cells are
    numbered along a Morton curve (a mesh generator's locality; lexicographic otherwise)."""
    nv1 = n + 1
    ii, jj, kk = np.meshgrid(np.arange(nv1), np.arange(nv1), np.arange(nv1), indexing="ij")
    ix, iy, iz = ii.ravel(), jj.ravel(), kk.ravel()
    h = 1.0 / n
    coords = np.stack([ix, iy, iz], axis=1).astype(np.float64) * h
    rng = np.random.default_rng(seed)
    noise = rng.uniform(-1.0, 1.0, size=coords.shape) * (jitter * h)
    interior = (ix > 0) & (ix < n) & (iy > 0) & (iy < n) & (iz > 0) & (iz < n)
    coords[interior] += noise[interior]

    def vid(a, b, c):
        return (a * nv1 + b) * nv1 + c

    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
    if locality:
        order = np.argsort(_morton3(ci, cj, ck), kind="stable")
        ci, cj, ck = ci[order], cj[order], ck[order]
    v0 = vid(ci, cj, ck)
    v1 = vid(ci + 1, cj, ck)
    v2 = vid(ci, cj + 1, ck)
    v3 = vid(ci + 1, cj + 1, ck)
    v4 = vid(ci, cj, ck + 1)
    v5 = vid(ci + 1, cj, ck + 1)
    v6 = vid(ci, cj + 1, ck + 1)
    v7 = vid(ci + 1, cj + 1, ck + 1)
    tets = np.stack([
        np.stack([v0, v1, v3, v7], axis=1),
        np.stack([v0, v1, v7, v5], axis=1),
        np.stack([v0, v5, v7, v4], axis=1),
        np.stack([v0, v3, v2, v7], axis=1),
        np.stack([v0, v6, v4, v7], axis=1),
        np.stack([v0, v2, v6, v7], axis=1),
    ], axis=1).reshape(-1, 4)
    if locality:
        vorder = np.argsort(_morton3(ix, iy, iz), kind="stable")
        newid = np.empty_like(vorder)
        newid[vorder] = np.arange(len(vorder))
        coords = coords[vorder]
        tets = newid[tets]
    cells = np.sort(tets, axis=1).astype(np.int32)
    return Mesh(coords, cells, np.ones(len(cells), dtype=np.int32), meta=dict(kind="box", n=n))


def tube_mesh(n, jitter=0.1, seed=1234, locality=True, r_fluid=0.8, length=4.0):
    """box_mesh(n) mapped to a cylinder of radius 1 and the given length (axis z);
    marker 1 (fluid) if the cell centroid radius < r_fluid else 2 (solid)."""
    m = box_mesh(n, jitter, seed, locality)
    a = 2.0 * m.coords[:, 0] - 1.0
    b = 2.0 * m.coords[:, 1] - 1.0
    x = a * np.sqrt(1.0 - 0.5 * b * b)
    y = b * np.sqrt(1.0 - 0.5 * a * a)
    coords = np.stack([x, y, length * m.coords[:, 2]], axis=1)
    cen = coords[m.cells].mean(axis=1)
    r = np.hypot(cen[:, 0], cen[:, 1])
    markers = np.where(r < r_fluid, 1, 2).astype(np.int32)
    return Mesh(coords, m.cells, markers, meta=dict(kind="tube", n=n))


def synthetic_state(mesh, space, seed, h=None):
    """SURVEY 8d "State distribution": d smooth + noise (keeps det F > 0), v ~ U(-1,1),
    p ~ 1e3 U(-1,1) (w ~ U(-1,1) if present). Returns x[N] in the space's global numbering."""
    g = mesh.gdim
    if h is None:
        h = mesh.hmin()
    rng = np.random.default_rng(seed)
    X = space.node_coords(mesh)                       # [nnodes, g]
    nn = X.shape[0]
    two_pi = 2.0 * np.pi
    s = [np.sin(two_pi * X[:, c]) for c in range(g)]
    c_ = [np.cos(two_pi * X[:, c]) for c in range(g)]
    if g == 3:
        smooth = np.stack([s[0] * c_[1] * s[2], c_[0] * s[1] * s[2], s[0] * s[1] * c_[2]], axis=1)
    else:
        smooth = np.stack([s[0] * c_[1], c_[0] * s[1]], axis=1)
    d = 1e-2 * h * smooth + 1e-3 * h * rng.uniform(-1, 1, size=(nn, g))
    v = rng.uniform(-1, 1, size=(nn, g))
    p = 1e3 * rng.uniform(-1, 1, size=mesh.num_vertices)
    x = np.zeros(space.num_dofs)
    for a in range(g):
        x[space.node_dof(0, a)] = d[:, a]
        x[space.node_dof(1, a)] = v[:, a]
    x[space.vertex_dof_p()] = p
    if space.biharmonic:
        w = rng.uniform(-1, 1, size=(nn, g))
        for a in range(g):
            x[space.node_dof(3, a)] = w[:, a]
    return x


def rectangle_mesh(p0, p1, nx, ny, diagonal="right"):
    """DOLFIN RectangleMesh [dep]: vertex iy*(nx+1)+ix at p0 + (ix*hx, iy*hy); every square
    (v0 v1 / v2 v3, v1 = v0+1, v2 = v0+nx+1) is cut into (v0,v1,v3),(v0,v2,v3) for "right" and into
    (v0,v1,v2),(v1,v2,v3) for "left". All cells get marker 1."""
    if diagonal not in ("right", "left"):
        raise ValueError("diagonal must be 'right' or 'left'")
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    if diagonal == "right":
        tri = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], axis=1)
    else:
        tri = np.stack([np.stack([v0, v1, v2], 1), np.stack([v1, v2, v3], 1)], axis=1)
    cells = np.sort(tri.reshape(-1, 3), axis=1).astype(np.int32)
    return Mesh(coords, cells, np.ones(len(cells), dtype=np.int32), dict(kind="rectangle", nx=nx, ny=ny))
