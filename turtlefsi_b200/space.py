"""Host-side mixed function space DVP = [P2^g (d), P2^g (v), P1 (p) (, P2^g (w))]: cell->dof
map and the fixed CSR sparsity pattern.

Reference: turtleFSI/monolithic.py:77-92 builds the MixedElement; DOLFIN's DofMapBuilder and
SparsityPatternBuilder ([dep], absent here) produce the dof numbering and the pattern that
`assemble(..., keep_diagonal=True)` uses (turtleFSI/modules/newtonsolver.py:26-27).  The C-ABI
takes `cell_dofs`, `rowptr`, `colidx` as INPUT so that DOLFIN's arrays can be passed unchanged;
this module is the stand-in used when DOLFIN is absent:

* local dof order  [d_0(n2)..d_{g-1}, v_0..v_{g-1}, p(n1) (, w_0..w_{g-1})], P2 local nodes =
  vertices then edge midpoints in UFC edge order (cells have ascending vertex ids);
* global numbering "interleaved" (default): P2 nodes are ordered by their lowest vertex
  (vertex node first, then the edges anchored at it) and all dofs living on one node are
  consecutive (d.., v.., p at vertices, w..) — a locality-preserving numbering of the kind
  DOLFIN's graph reordering produces; "blocked": one contiguous block per scalar component;
* pattern: union over ALL cells of the dense ndof x ndof block (+ diagonal, already included),
  columns ascending within a row (PETSc AIJ), explicit zeros kept.
"""
import numpy as np

EDGES = {
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


class MixedSpace:
    def __init__(self, mesh, biharmonic=False, numbering="interleaved"):
        g = mesh.gdim
        self.gdim = g
        self.biharmonic = bool(biharmonic)
        self.numbering = numbering
        self.n1 = g + 1
        self.n2 = g + 1 + len(EDGES[g])
        self.nfields2 = 3 if biharmonic else 2            # P2 vector fields: d, v (, w)
        self.ndof_cell = self.nfields2 * g * self.n2 + self.n1
        cells = mesh.cells.astype(np.int64)
        nc, nv = cells.shape[0], mesh.num_vertices
        self.num_cells, self.num_vertices = nc, nv

        # ---- edges -------------------------------------------------------------------
        lo = np.stack([cells[:, i] for (i, j) in EDGES[g]], axis=1)
        hi = np.stack([cells[:, j] for (i, j) in EDGES[g]], axis=1)
        key = lo * nv + hi                                 # cells ascending => lo < hi
        ukey, inv = np.unique(key.ravel(), return_inverse=True)
        ne = len(ukey)
        self.num_edges = ne
        self.edge_vertices = np.stack([ukey // nv, ukey % nv], axis=1)
        cell_edges = inv.reshape(nc, -1)
        self.cell_edges = cell_edges

        # ---- P2 nodes ordered by anchor vertex (vertex node, then its edges) -----------
        nn = nv + ne
        self.num_nodes = nn
        anchor = np.concatenate([np.arange(nv), self.edge_vertices[:, 0]])
        tieb = np.concatenate([np.zeros(nv, dtype=np.int64), 1 + self.edge_vertices[:, 1]])
        order = np.lexsort((tieb, anchor))                 # raw node id (v | nv+e) -> sorted
        node_of_raw = np.empty(nn, dtype=np.int64)
        node_of_raw[order] = np.arange(nn)
        self.vertex_node = node_of_raw[:nv]                # node id of vertex v
        self.edge_node = node_of_raw[nv:]                  # node id of edge e
        self.node_is_vertex = np.zeros(nn, dtype=bool)
        self.node_is_vertex[self.vertex_node] = True
        self.cell_nodes = np.concatenate(
            [self.vertex_node[cells], self.edge_node[cell_edges]], axis=1).astype(np.int32)

        # ---- global dof numbering ------------------------------------------------------
        nf2 = self.nfields2
        ncomp = np.where(self.node_is_vertex, nf2 * g + 1, nf2 * g).astype(np.int64)
        self.node_ncomp = ncomp
        if numbering == "interleaved":
            base = np.concatenate([[0], np.cumsum(ncomp)])
            self.node_base = base[:-1]
            self.num_dofs = int(base[-1])
        elif numbering == "blocked":
            self.node_base = None
            self.num_dofs = nf2 * g * nn + nv
        else:
            raise ValueError(numbering)
        self._vertex_rank = np.cumsum(self.node_is_vertex) - 1   # node -> index among vertices

        # local layout: d comps, v comps, p, w comps  (field ids: d=0, v=1, p=2, w=3)
        cd = []
        for f in (0, 1):
            for a in range(g):
                cd.append(self.node_dof(f, a)[self.cell_nodes])
        cd.append(self.vertex_dof_p()[cells])
        if biharmonic:
            for a in range(g):
                cd.append(self.node_dof(3, a)[self.cell_nodes])
        self.cell_dofs = np.concatenate(cd, axis=1).astype(np.int32)
        assert self.cell_dofs.shape[1] == self.ndof_cell

    # -------------------------------------------------------------------------------------
    def node_dof(self, field, comp):
        """global dof of component `comp` of P2 field (0=d, 1=v, 3=w) at every node [nn]."""
        g, nn = self.gdim, self.num_nodes
        if self.numbering == "interleaved":
            if field in (0, 1):
                slot = field * g + comp
                return self.node_base + slot
            slot = 2 * g + comp + self.node_is_vertex.astype(np.int64)   # after p on vertices
            return self.node_base + slot
        fidx = {0: 0, 1: 1, 3: 2}[field]
        blk = fidx * g + comp
        off = blk * nn if fidx < 2 else 2 * g * nn + self.num_vertices + comp * nn
        return off + np.arange(nn, dtype=np.int64)

    def vertex_dof_p(self):
        """global dof of p at every mesh vertex [nv]."""
        g = self.gdim
        if self.numbering == "interleaved":
            return self.node_base[self.vertex_node] + 2 * g
        return 2 * g * self.num_nodes + np.arange(self.num_vertices, dtype=np.int64)

    def edge_lookup(self):
        """{(lower vertex, higher vertex): edge id}, built once."""
        if getattr(self, "_edge_lookup", None) is None:
            self._edge_lookup = {(int(a), int(b)): i for i, (a, b) in enumerate(self.edge_vertices)}
        return self._edge_lookup

    def node_coords(self, mesh):
        X = np.zeros((self.num_nodes, self.gdim))
        X[self.vertex_node] = mesh.coords
        X[self.edge_node] = 0.5 * (mesh.coords[self.edge_vertices[:, 0]] +
                                   mesh.coords[self.edge_vertices[:, 1]])
        return X

    def sub_dofs(self, field, comp=None):
        """all global dofs of a sub-space: field 0=d, 1=v, 2=p, 3=w; comp None = all comps."""
        if field == 2:
            return np.asarray(self.vertex_dof_p(), dtype=np.int64)
        comps = range(self.gdim) if comp is None else [comp]
        return np.concatenate([self.node_dof(field, a) for a in comps])

    # -------------------------------------------------------------------------------------
    def node_graph(self):
        """CSR adjacency of P2 nodes (two nodes adjacent iff they share a cell), sorted."""
        nn = self.num_nodes
        cn = self.cell_nodes.astype(np.int64)
        keys = []
        step = max(1, 2_000_000 // (self.n2 * self.n2))
        for s in range(0, self.num_cells, step):
            blk = cn[s:s + step]
            k = (blk[:, :, None] * nn + blk[:, None, :]).ravel()
            keys.append(np.unique(k))
        ukey = np.unique(np.concatenate(keys))
        rows, cols = ukey // nn, ukey % nn
        nrowptr = np.zeros(nn + 1, dtype=np.int64)
        np.add.at(nrowptr, rows + 1, 1)
        return np.cumsum(nrowptr), cols

    def pattern(self):
        """(rowptr int64 [N+1], colidx int32 [nnz]) — DOLFIN-style dense-cell-block pattern."""
        if self.numbering != "interleaved":
            return pattern_from_cell_dofs(self.cell_dofs, self.num_dofs)
        nrp, ncol = self.node_graph()
        nn = self.num_nodes
        ncomp, base = self.node_ncomp, self.node_base
        # dof-length of every node row, then per dof row
        rowlen_node = np.add.reduceat(ncomp[ncol], nrp[:-1])
        rowlen = np.repeat(rowlen_node, ncomp)
        rowptr = np.concatenate([[0], np.cumsum(rowlen)]).astype(np.int64)
        # columns of one node row (expanded), then repeated for each dof of the node
        nnz_node = np.diff(nrp)
        cnt = ncomp[ncol]
        start = np.repeat(base[ncol], cnt)
        within = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        cols_node = (start + within)                       # concatenated over node rows
        noff = np.concatenate([[0], np.cumsum(rowlen_node)])
        # repeat every node-row segment ncomp[node] times
        seg_start = np.repeat(noff[:-1], ncomp)
        seg_len = rowlen
        idx = np.repeat(seg_start, seg_len) + (np.arange(seg_len.sum()) -
                                               np.repeat(rowptr[:-1], seg_len))
        colidx = cols_node[idx].astype(np.int32)
        del nnz_node
        return rowptr, colidx


def pattern_device(space, device):
    """(rowptr, colidx) as torch tensors on `device`, built there: node adjacency by a device
    sort/unique (torch = plumbing), dof-level expansion by the library's CUDA kernel. Needs the
    "interleaved" numbering. Same result as MixedSpace.pattern() (tests/test_gpu_parity.py)."""
    import ctypes as C

    import torch

    from . import lib as _lib
    assert space.numbering == "interleaved"
    L = _lib.load()
    nn = space.num_nodes
    cn = torch.as_tensor(space.cell_nodes.astype(np.int64)).to(device)
    keys = []
    step = max(1, 50_000_000 // (space.n2 * space.n2))
    for s in range(0, space.num_cells, step):
        blk = cn[s:s + step]
        keys.append(torch.unique((blk[:, :, None] * nn + blk[:, None, :]).reshape(-1)))
    ukey = torch.unique(torch.cat(keys))
    del keys, cn
    rows = ukey // nn
    ncol = (ukey % nn).to(torch.int32)
    del ukey
    nrowptr = torch.zeros(nn + 1, dtype=torch.int64, device=device)
    nrowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=nn), 0)
    del rows
    ncomp = torch.as_tensor(space.node_ncomp.astype(np.int32)).to(device)
    nbase = torch.as_tensor(space.node_base.astype(np.int64)).to(device)
    ncum = torch.zeros(ncol.numel() + 1, dtype=torch.int64, device=device)
    ncum[1:] = torch.cumsum(ncomp[ncol.long()].long(), 0)
    rowlen_node = ncum[nrowptr[1:]] - ncum[nrowptr[:-1]]
    rowptr = torch.zeros(space.num_dofs + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(torch.repeat_interleave(rowlen_node, ncomp.long()), 0)
    nnz = int(rowptr[-1].item())
    colidx = torch.empty(nnz, dtype=torch.int32, device=device)
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(device):
        rc = L.tfsi_expand_node_pattern(nn, ncol.numel(), p(nrowptr), p(ncol), p(ncum), p(nbase), p(ncomp),
                                        p(rowptr), p(colidx), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        if rc != 0:
            raise RuntimeError("tfsi_expand_node_pattern: %s" % L.tfsi_last_error(None).decode())
        torch.cuda.synchronize(device)
    return rowptr, colidx


def pattern_from_cell_dofs(cell_dofs, num_dofs):
    """generic (any numbering) pattern: union of dense cell blocks, columns ascending."""
    cd = cell_dofs.astype(np.int64)
    nd = cd.shape[1]
    step = max(1, 4_000_000 // (nd * nd))
    keys = []
    for s in range(0, cd.shape[0], step):
        blk = cd[s:s + step]
        keys.append(np.unique((blk[:, :, None] * num_dofs + blk[:, None, :]).ravel()))
    ukey = np.unique(np.concatenate(keys))
    rows, cols = ukey // num_dofs, ukey % num_dofs
    rowptr = np.zeros(num_dofs + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr), cols.astype(np.int32)
