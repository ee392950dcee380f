"""Cell partition of an ARBITRARY mesh across the GPUs of one box, the general counterpart
of the slab partition that bench.py uses for its synthetic tube (turtlefsi_b200/partition.py).

The reference runs SPMD under MPI: DOLFIN partitions the cells (SCOTCH) and keeps a ghost layer
(`parameters["ghost_mode"] = "shared_vertex"`, problems/Test_Cylinder/problem_aneu.py:14-17), PETSc owns
row blocks [dep]. Here, with no external partitioner:

* cells -> parts by recursive coordinate bisection of the centroids (deterministic, balanced to one cell);
* a mesh node (vertex or edge midpoint) is owned by ONE of the parts touching it, chosen by a deterministic
  balanced rule (`node_owners`; the plain "lowest part" rule makes part 0 assemble all the ghost layers)
  [DOLFIN votes at random: ownership need not match, only the global numbering does];
* part r assembles its own cells plus the ghost cells that touch a node it owns, so that every OWNED
  CSR row is complete without any reduction of matrix entries; rows of ghost dofs are not used;
* before an assembly the ghost dofs of the iterate are refreshed from their owners
  (tfsi_halo_pack -> NCCL send/recv through torch.distributed -> tfsi_halo_unpack); lists are ordered
  by global node id on both sides, so no index exchange is needed;
* norms: sums of squares over owned dofs + one all-reduce.

Every rank holds the GLOBAL mesh on the host (fine for the problem files' meshes; the 2 M-tet-per-GPU
bench builds its slabs rank-locally instead). Local vertex numbering is monotone in the global one, so
cells keep their ascending-vertex order and with it the orientation of every local edge.
"""
import numpy as np

from .mesh import Mesh
from .space import MixedSpace


def rcb_partition(mesh, nparts, metric="length"):
    """part id of every cell: recursive coordinate bisection of the cell centroids. The widest direction
    is cut so that the two sides get cell counts proportional to the parts they will hold. metric="cells" measures
    the width in CELLS (extent / median cell extent along the axis) instead of length: on anisotropic meshes the cut
    then crosses the fewest cells, i.e. the interface (and the ghost layer) is smallest."""
    X = mesh.coords[mesh.cells]
    cen = X.mean(axis=1)
    if metric == "cells":
        cen = cen / np.median(X.max(axis=1) - X.min(axis=1), axis=0)
    elif metric != "length":
        raise ValueError(metric)
    part = np.zeros(mesh.num_cells, dtype=np.int32)

    def split(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        left = np_ // 2
        c = cen[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = np.lexsort((idx, c[:, axis]))                 # ties broken by cell index: deterministic
        k = (len(idx) * left) // np_
        split(idx[order[:k]], p0, left)
        split(idx[order[k:]], p0 + left, np_ - left)
    split(np.arange(mesh.num_cells), 0, nparts)
    return part


def node_owners(cell_nodes, cell_part, world, num_nodes, balanced=True):
    """owner part of every mesh node (vertex or edge midpoint), a pure function of the SET of parts touching the node,
    so that every rank computes the same answer from the global cell partition.

    balanced=False: the lowest part touching the node (then part 0 owns all of its interfaces and assembles the most
    ghost cells, the last part none). balanced=True: every pair of neighbouring parts (p, q) gets an orientation
    "p owns the nodes shared with q", chosen greedily — largest interfaces first, towards the part that owns less so
    far — and a node shared by several parts goes to the part that wins most of the pairwise contests inside its set
    (ties: lowest part). The ghost layers (and with them the redundant cell work) are spread over the parts."""
    if world > 62:
        balanced = False
    cn = cell_nodes
    if not balanced:
        owner = np.full(num_nodes, world, dtype=np.int32)
        np.minimum.at(owner, cn.ravel(), np.repeat(cell_part, cn.shape[1]))
        return owner
    mask = np.zeros(num_nodes, dtype=np.int64)
    for p in range(world):                                    # parts touching each node, as a bit set
        nodes = np.unique(cn[cell_part == p])
        mask[nodes] |= np.int64(1) << np.int64(p)
    umask, inv, cnt = np.unique(mask, return_inverse=True, return_counts=True)
    sets = [[p for p in range(world) if (int(m) >> p) & 1] for m in umask]
    pair = {}
    for st, c in zip(sets, cnt):
        for i, p in enumerate(st):
            for q in st[i + 1:]:
                pair[(p, q)] = pair.get((p, q), 0) + int(c)
    load = np.zeros(world, dtype=np.int64)
    beats = np.zeros((world, world), dtype=bool)              # beats[p, q]: p owns what it shares with q
    ordered = sorted(pair.items(), key=lambda kv: (-kv[1], kv[0]))
    # the (up to 16) largest interfaces: exhaustive search for the orientation with the smallest maximum load
    # (then the smallest sum of squares); the small rest greedily
    top = ordered[:16]
    if top:
        k = len(top)
        P = np.array([pq[0] for pq, _ in top]); Q = np.array([pq[1] for pq, _ in top])
        Cw = np.array([c for _, c in top], dtype=np.int64)
        bits = (np.arange(1 << k, dtype=np.int64)[:, None] >> np.arange(k)) & 1          # 1: q owns
        loads = np.zeros((1 << k, world), dtype=np.int64)
        for j in range(k):
            np.add.at(loads, (np.arange(1 << k), np.where(bits[:, j] == 1, Q[j], P[j])), Cw[j])
        score = loads.max(axis=1).astype(np.float64) * 1e6 + (loads.astype(np.float64) ** 2).sum(axis=1) / max(float(Cw.sum()) ** 2, 1.0)
        best = int(np.argmin(score))                                                     # first minimum: deterministic
        for j in range(k):
            w, l = (Q[j], P[j]) if bits[best, j] else (P[j], Q[j])
            beats[w, l] = True
            load[w] += Cw[j]
    for (p, q), c in ordered[16:]:
        w, l = (p, q) if load[p] <= load[q] else (q, p)
        beats[w, l] = True
        load[w] += c
    owner_of_set = np.empty(len(umask), dtype=np.int32)
    for k, st in enumerate(sets):
        wins = [sum(beats[p, q] for q in st) for p in st]
        owner_of_set[k] = st[int(np.argmax(wins))] if st else world      # argmax: first maximum = lowest part
    return owner_of_set[inv]


class RankView:
    """what rank `rank` of `world` works on: local mesh (own + ghost cells), local space, maps to the
    global numbering and the halo lists per peer."""

    def __init__(self, mesh, rank, world, biharmonic=False, cell_part=None, gspace=None, balanced=True, node_owner=None):
        self.rank, self.world = rank, world
        self.cell_part = rcb_partition(mesh, world) if cell_part is None else cell_part
        gs = MixedSpace(mesh, biharmonic=biharmonic) if gspace is None else gspace
        self.gspace = gs
        cn = gs.cell_nodes                                                    # [nc, n2] global node ids
        owner = node_owners(cn, self.cell_part, world, gs.num_nodes, balanced) if node_owner is None else node_owner
        self.node_owner_global = owner
        own_cell = self.cell_part == rank
        ghost_cell = ~own_cell & (owner[cn] == rank).any(axis=1)
        self.cells_global = np.nonzero(own_cell | ghost_cell)[0]              # ascending: keeps the global order
        self.cell_owned = own_cell[self.cells_global]
        gcells = mesh.cells[self.cells_global]
        gverts = np.unique(gcells)                                            # ascending -> monotone renumbering
        lcells = np.searchsorted(gverts, gcells).astype(np.int32)
        self.vertex_global = gverts
        self.mesh = Mesh(mesh.coords[gverts], lcells,
                         None if mesh.markers is None else mesh.markers[self.cells_global].copy(),
                         dict(kind="rank-view", rank=rank, world=world))
        ls = MixedSpace(self.mesh, biharmonic=biharmonic)
        self.space = ls
        # local node -> global node (vertex nodes through the vertex map, edge nodes through their end points)
        node_global = np.empty(ls.num_nodes, dtype=np.int64)
        node_global[ls.vertex_node] = gs.vertex_node[gverts]
        ev = gverts[ls.edge_vertices]                                          # global end points, ascending pairs
        gkey = gs.edge_vertices[:, 0].astype(np.int64) * mesh.num_vertices + gs.edge_vertices[:, 1]
        order = np.argsort(gkey)
        pos = np.searchsorted(gkey[order], ev[:, 0].astype(np.int64) * mesh.num_vertices + ev[:, 1])
        node_global[ls.edge_node] = gs.edge_node[order[pos]]
        self.node_global = node_global
        self.node_owner = owner[node_global]
        self.node_owned = self.node_owner == rank
        # local dof -> global dof, owned-dof mask
        self.dof_global = np.empty(ls.num_dofs, dtype=np.int64)
        for field in (0, 1) + ((3,) if biharmonic else ()):
            for a in range(mesh.gdim):
                self.dof_global[ls.node_dof(field, a)] = gs.node_dof(field, a)[node_global]
        self.dof_global[ls.vertex_dof_p()] = np.asarray(gs.vertex_dof_p())[gverts]
        self._slot_table = None
        self.dof_owned = np.zeros(ls.num_dofs, dtype=bool)
        self.dof_owned[self._ordered_dofs(np.nonzero(self.node_owned)[0])] = True

    def halo_lists(self, views=None):
        """(send, recv): dicts peer -> int32 local dof lists. recv[p] = my ghost dofs owned by p; send[p] = my
        owned dofs that p holds as ghosts. Both ordered by (global node id, then dof slot), so they pair up.
        `views`: optional list of all RankView objects (tests); by default the peers' node sets are rebuilt
        from the global mesh, which every rank holds."""
        send, recv = {}, {}
        mine = self.node_global
        for p in range(self.world):
            if p == self.rank:
                continue
            ghosts = np.nonzero(self.node_owner == p)[0]
            if len(ghosts):
                ghosts = ghosts[np.argsort(self.node_global[ghosts])]
                recv[p] = self._ordered_dofs(ghosts)
            peer_nodes = views[p].node_global if views is not None else self._peer_nodes(p)
            want = np.intersect1d(peer_nodes, mine[self.node_owned])          # sorted by global id
            if len(want):
                lut = np.argsort(mine)
                loc = lut[np.searchsorted(mine[lut], want)]
                send[p] = self._ordered_dofs(loc)
        return send, recv

    def _ordered_dofs(self, nodes):
        """dofs of the nodes, node after node (d comps, v comps, p if vertex, w comps): same order on every rank."""
        if self._slot_table is None:
            ls = self.space
            g = self.mesh.gdim
            cols = [ls.node_dof(f, a) for f in (0, 1) for a in range(g)]
            pcol = np.full(ls.num_nodes, -1, dtype=np.int64)
            pcol[ls.vertex_node] = np.asarray(ls.vertex_dof_p())
            cols.append(pcol)
            if ls.biharmonic:
                cols += [ls.node_dof(3, a) for a in range(g)]
            self._slot_table = np.stack(cols, axis=1)
        flat = self._slot_table[np.asarray(nodes, dtype=np.int64)].ravel()
        return flat[flat >= 0].astype(np.int32)

    def _peer_nodes(self, p):
        """global node ids present on rank p (its own + ghost cells)."""
        cn = self.gspace.cell_nodes
        own = self.cell_part == p
        ghost = ~own & (self.node_owner_global[cn] == p).any(axis=1)
        return np.unique(cn[own | ghost])

    def scatter_global(self, xg):
        """local vector from a global one."""
        return np.asarray(xg)[self.dof_global]


class HaloExchangeRCB:
    """refreshes the ghost dofs of a device vector from their owners (any number of peers)."""

    def __init__(self, asm, send, recv, rank):
        import torch
        self.asm, self.rank = asm, rank
        dev = asm.device
        mk = lambda d: {p: torch.as_tensor(v).to(dev) for p, v in d.items()}
        self.send_idx, self.recv_idx = mk(send), mk(recv)
        buf = lambda d: {p: torch.empty(len(v), dtype=torch.float64, device=dev) for p, v in d.items()}
        self.send_buf, self.recv_buf = buf(send), buf(recv)

    def exchange(self, x, overlap=None):
        """overlap: optional callable run after the transfers are queued and before they are waited for (work that
        does not depend on the ghost values, e.g. the initialisation A := A_pre, then runs under the transfer)."""
        import ctypes as C

        import torch.distributed as dist
        asm = self.asm
        st = asm._stream()
        ptr = lambda t: C.c_void_p(t.data_ptr())
        ops = []
        for p, idx in sorted(self.send_idx.items()):
            asm._check(asm.lib.tfsi_halo_pack(asm.h, ptr(x), ptr(idx), idx.numel(), ptr(self.send_buf[p]), st))
            ops.append(dist.P2POp(dist.isend, self.send_buf[p], p))
        for p in sorted(self.recv_idx):
            ops.append(dist.P2POp(dist.irecv, self.recv_buf[p], p))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if overlap is not None:
            overlap()
        for req in reqs:
            req.wait()
        for p, idx in sorted(self.recv_idx.items()):
            asm._check(asm.lib.tfsi_halo_unpack(asm.h, ptr(x), ptr(idx), idx.numel(), ptr(self.recv_buf[p]), st))
