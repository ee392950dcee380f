"""Cell partition of the synthetic tube across the GPUs of one box + ghost-dof halo exchange.

The reference is SPMD over MPI ranks: DOLFIN partitions the cells, keeps a ghost layer
(`parameters["ghost_mode"] = "shared_vertex"`, problems/Test_Cylinder/problem_aneu.py:17) and
PETSc exchanges ghost values / off-process rows [dep]. Here:

* one process per GPU; rank r owns the cells of the z-slab [r n, (r+1) n) of an n x n x (world n)
  box mapped to the tube (weak scaling: the tube gets longer with the number of GPUs);
* a mesh node (vertex or edge midpoint) is owned by the LOWEST rank touching it; rank r assembles
  its own cells plus the ghost layer of cells (one layer of cubes above its slab) that touch nodes
  it owns, so every owned CSR row is complete without any reduction of matrix entries; rows of
  non-owned (ghost) dofs are simply not used;
* before each assembly the ghost dofs of the Newton iterate are refreshed: the bottom plane comes
  from rank r-1, the ghost layer above from rank r+1 (packed with tfsi_halo_pack, NCCL
  send/recv over NVLink via torch.distributed, tfsi_halo_unpack);
* norms are sums of squares over OWNED dofs + one all-reduce.

Each rank builds its local mesh directly (no global mesh in memory); coordinates, jitter and the
synthetic state are pure functions of GLOBAL indices, so ranks agree on shared nodes bit for bit.
"""
import time

import numpy as np

from .mesh import Mesh, _morton3


def _hash01(k, salt):
    """deterministic pseudo-random numbers in [0, 1) from integer keys (same on every rank)."""
    k = np.asarray(k, dtype=np.uint64)
    z = k + np.uint64((int(salt) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)      # wraps mod 2^64
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


class SlabMesh:
    """local mesh of one rank: own slab + ghost cube layer above."""

    def __init__(self, n, rank, world, jitter=0.1, r_fluid=0.8, length_per_rank=4.0, nz=None):
        nz = n if nz is None else nz          # cubes per rank along the tube axis
        self.n, self.nz, self.rank, self.world = n, nz, rank, world
        nzg = nz * world                      # cubes in z, globally
        z0 = rank * nz
        z1 = min((rank + 1) * nz + 1, nzg)    # exclusive; +1 = ghost layer above
        nzl = z1 - z0
        nv1 = n + 1
        ii, jj, kk = np.meshgrid(np.arange(nv1), np.arange(nv1), np.arange(z0, z1 + 1), indexing="ij")
        ix, iy, iz = ii.ravel(), jj.ravel(), kk.ravel()
        gvid = (ix.astype(np.int64) * nv1 + iy) * (nzg + 1) + iz          # global vertex id
        h = 1.0 / n
        ref = np.stack([ix * h, iy * h, iz * h], axis=1).astype(np.float64)   # z in units of the cube size
        interior = (ix > 0) & (ix < n) & (iy > 0) & (iy < n) & (iz > 0) & (iz < nzg)
        noise = np.stack([_hash01(gvid, s) for s in (1, 2, 3)], axis=1) * 2.0 - 1.0
        ref[interior] += jitter * h * noise[interior]
        # local numbering along a Morton curve
        order = np.argsort(_morton3(ix, iy, iz - z0), kind="stable")
        newid = np.empty_like(order)
        newid[order] = np.arange(len(order))

        def vid(a, b, c):
            return newid[(a * nv1 + b) * (nzl + 1) + (c - z0)]
        ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(z0, z1), indexing="ij")
        ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
        corder = np.argsort(_morton3(ci, cj, ck - z0), kind="stable")
        ci, cj, ck = ci[corder], cj[corder], ck[corder]
        v0, v1, v2, v3 = vid(ci, cj, ck), vid(ci + 1, cj, ck), vid(ci, cj + 1, ck), vid(ci + 1, cj + 1, ck)
        v4, v5, v6, v7 = vid(ci, cj, ck + 1), vid(ci + 1, cj, ck + 1), vid(ci, cj + 1, ck + 1), vid(ci + 1, cj + 1, ck + 1)
        tets = np.stack([
            np.stack([v0, v1, v3, v7], axis=1), np.stack([v0, v1, v7, v5], axis=1),
            np.stack([v0, v5, v7, v4], axis=1), np.stack([v0, v3, v2, v7], axis=1),
            np.stack([v0, v6, v4, v7], axis=1), np.stack([v0, v2, v6, v7], axis=1)], axis=1).reshape(-1, 4)
        self.cell_owned = np.repeat(ck < (rank + 1) * nz, 6)
        # NOTE: cells must list vertices in ascending GLOBAL order so that every rank sees the same
        # local edge numbering / orientation as a single-domain mesh would (DOLFIN ordered mesh)
        gv_of_local = np.empty(len(order), dtype=np.int64)
        gv_of_local[newid] = gvid
        srt = np.argsort(gv_of_local[tets], axis=1)
        cells = np.take_along_axis(tets, srt, axis=1)
        a = 2.0 * ref[:, 0] - 1.0
        b = 2.0 * ref[:, 1] - 1.0
        coords = np.stack([a * np.sqrt(1.0 - 0.5 * b * b), b * np.sqrt(1.0 - 0.5 * a * a),
                           length_per_rank * ref[:, 2]], axis=1)
        coords_l = np.empty_like(coords)
        coords_l[newid] = coords
        cen = coords_l[cells].mean(axis=1)
        markers = np.where(np.hypot(cen[:, 0], cen[:, 1]) < r_fluid, 1, 2).astype(np.int32)
        self.mesh = Mesh(coords_l, cells.astype(np.int32), markers, meta=dict(kind="tube-slab", n=n, rank=rank, world=world))
        self.global_vertex = gv_of_local
        izl = np.empty(len(order), dtype=np.int64)
        izl[newid] = iz
        self.vertex_iz = izl
        self.num_global_vertices = nv1 * nv1 * (nzg + 1)


class LocalSpaceGlobalKeys:
    """global keys / owners of the nodes of a rank-local MixedSpace."""

    def __init__(self, slab, space):
        n, rank, world = slab.nz, slab.rank, slab.world
        nvg = slab.num_global_vertices
        gv, viz = slab.global_vertex, slab.vertex_iz
        nn = space.num_nodes
        key = np.zeros(nn, dtype=np.int64)
        maxz = np.zeros(nn, dtype=np.int64)
        key[space.vertex_node] = gv                               # vertex key = global vertex id
        maxz[space.vertex_node] = viz
        ev = space.edge_vertices
        glo = np.minimum(gv[ev[:, 0]], gv[ev[:, 1]])
        ghi = np.maximum(gv[ev[:, 0]], gv[ev[:, 1]])
        key[space.edge_node] = nvg + glo * nvg + ghi              # edge key, disjoint from vertex keys
        maxz[space.edge_node] = np.maximum(viz[ev[:, 0]], viz[ev[:, 1]])
        self.key, self.maxz = key, maxz
        owner = np.clip((maxz - 1) // n, 0, world - 1)
        owner[maxz == 0] = 0
        self.owner = owner
        self.owned = owner == rank

    def node_dofs(self, space, nodes):
        """all dofs living on the given nodes, node-major (interleaved numbering: contiguous)."""
        out = []
        for nd in nodes:
            out.append(np.arange(space.node_base[nd], space.node_base[nd] + space.node_ncomp[nd]))
        return np.concatenate(out).astype(np.int32) if out else np.zeros(0, dtype=np.int32)


def halo_lists(slab, space, gk):
    """index lists (local dofs, ordered by global node key) for the two neighbours."""
    n, rank, world = slab.nz, slab.rank, slab.world
    top = (rank + 1) * n
    bot = rank * n

    def sorted_nodes(mask):
        nodes = np.nonzero(mask)[0]
        return nodes[np.argsort(gk.key[nodes], kind="stable")]

    def dofs(nodes):
        base, nc = space.node_base[nodes], space.node_ncomp[nodes]
        rep = np.repeat(base, nc)
        within = np.arange(nc.sum()) - np.repeat(np.cumsum(nc) - nc, nc)
        return (rep + within).astype(np.int32)
    lists = dict(send_up=None, send_down=None, recv_below=None, recv_above=None)
    if rank + 1 < world:
        lists["send_up"] = dofs(sorted_nodes(gk.owned & (gk.maxz == top) & _all_on_plane(slab, space, top)))
        lists["recv_above"] = dofs(sorted_nodes(gk.owner == rank + 1))
    if rank > 0:
        lists["recv_below"] = dofs(sorted_nodes(gk.owner == rank - 1))
        lists["send_down"] = dofs(sorted_nodes(gk.owned & (gk.maxz > bot) & (gk.maxz <= bot + 1)))
    return lists


def _all_on_plane(slab, space, k):
    """nodes lying entirely on the vertex plane z-index k."""
    viz = slab.vertex_iz
    on = np.zeros(space.num_nodes, dtype=bool)
    on[space.vertex_node] = viz == k
    ev = space.edge_vertices
    on[space.edge_node] = (viz[ev[:, 0]] == k) & (viz[ev[:, 1]] == k)
    return on


def synthetic_state_global(slab, space, gk, seed, h):
    """SURVEY 8d state distribution as a pure function of global node keys and coordinates."""
    mesh = slab.mesh
    X = space.node_coords(mesh)
    two_pi = 2.0 * np.pi
    s = [np.sin(two_pi * X[:, c]) for c in range(3)]
    c_ = [np.cos(two_pi * X[:, c]) for c in range(3)]
    smooth = np.stack([s[0] * c_[1] * s[2], c_[0] * s[1] * s[2], s[0] * s[1] * c_[2]], axis=1)
    u = lambda salt: _hash01(gk.key, seed * 16 + salt) * 2.0 - 1.0
    x = np.zeros(space.num_dofs)
    for a in range(3):
        x[space.node_dof(0, a)] = 1e-2 * h * smooth[:, a] + 1e-3 * h * u(a)
        x[space.node_dof(1, a)] = u(3 + a)
    x[space.vertex_dof_p()] = 1e3 * (_hash01(slab.global_vertex, seed * 16 + 7) * 2.0 - 1.0)
    return x


class HaloExchange:
    """refreshes the ghost dofs of a device vector from the owning ranks."""

    def __init__(self, asm, lists, rank, world, num_owned_cells, owned_dof_mask=None):
        import torch
        self.asm, self.rank, self.world = asm, rank, world
        self.num_owned_cells = int(num_owned_cells)
        dev = asm.device
        self.idx = {k: (torch.as_tensor(v).to(dev) if v is not None else None) for k, v in lists.items()}
        self.buf = {k: (torch.empty(len(v), dtype=torch.float64, device=dev) if v is not None else None)
                    for k, v in lists.items()}
        self.owned_mask = owned_dof_mask

    def exchange(self, x, overlap=None):
        """overlap: optional callable run after the transfers are queued and before they are waited for."""
        import ctypes as C

        import torch
        import torch.distributed as dist
        asm = self.asm
        st = asm._stream()
        p = lambda t: C.c_void_p(t.data_ptr())
        ops = []
        for key, peer in (("send_up", self.rank + 1), ("send_down", self.rank - 1)):
            if self.idx[key] is not None:
                asm._check(asm.lib.tfsi_halo_pack(asm.h, p(x), p(self.idx[key]), self.idx[key].numel(), p(self.buf[key]), st))
                ops.append(dist.P2POp(dist.isend, self.buf[key], peer))
        for key, peer in (("recv_below", self.rank - 1), ("recv_above", self.rank + 1)):
            if self.idx[key] is not None:
                ops.append(dist.P2POp(dist.irecv, self.buf[key], peer))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if overlap is not None:
            overlap()
        for req in reqs:
            req.wait()
        for key in ("recv_below", "recv_above"):
            if self.idx[key] is not None:
                asm._check(asm.lib.tfsi_halo_unpack(asm.h, p(x), p(self.idx[key]), self.idx[key].numel(), p(self.buf[key]), st))


def build_partitioned_workload(n, device, rank, world, regions, dt, theta, qdeg_j, qdeg_r):
    """the bench.py workload on `world` GPUs (see module docstring)."""
    from .assembler import FSIAssembler
    from .space import MixedSpace
    t0 = time.time()
    slab = SlabMesh(n, rank, world)
    mesh = slab.mesh
    space = MixedSpace(mesh)
    gk = LocalSpaceGlobalKeys(slab, space)
    asm = FSIAssembler(mesh, space, regions, dt, theta, extrapolation="laplace", extrapolation_sub_type="constant",
                       quad_degree_jacobian=qdeg_j, quad_degree_residual=qdeg_r, device=device, verify_pattern=0,
                       hmin=2.0 / n)
    x = synthetic_state_global(slab, space, gk, 2024, 2.0 / n)
    x1 = synthetic_state_global(slab, space, gk, 2025, 2.0 / n)
    X = space.node_coords(mesh)
    r = np.hypot(X[:, 0], X[:, 1])
    on_bnd = (X[:, 2] < 1e-9) | (X[:, 2] > 4.0 * world - 1e-9) | (r > 0.999)
    bc = np.concatenate([space.node_dof(f, a)[on_bnd] for f in (0, 1) for a in range(3)]).astype(np.int32)
    asm.set_bc_dofs(bc)
    halo = HaloExchange(asm, halo_lists(slab, space, gk), rank, world, int(slab.cell_owned.sum()))
    halo.owned_node_mask = gk.owned
    ghost_nodes = np.nonzero(~gk.owned)[0]
    base, nc = space.node_base[ghost_nodes], space.node_ncomp[ghost_nodes]
    halo.ghost_dofs = (np.repeat(base, nc) + (np.arange(nc.sum()) - np.repeat(np.cumsum(nc) - nc, nc))).astype(np.int64)
    return mesh, space, asm, x, x1, bc, halo, time.time() - t0
