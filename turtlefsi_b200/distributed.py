"""SPMD run of a problem file over several ranks — what `mpirun -np N turtleFSI ...` does in the reference
(tests/test_turtleFSI.py:40-44 runs TF_fsi with 1 and 2 processes against the same goldens).

`DistributedAssembler` wraps a rank-LOCAL assembler (turtlefsi_b200.assembler.FSIAssembler on a GPU; the
oracle backend in the CPU tests) built on a `partition_rcb.RankView`, and presents the assembler interface
that modules/newtonsolver.py and the problem files use, with the reference's distributed semantics:

* vectors live rank-locally (owned + ghost dofs); the owner's value is the authoritative one;
* before every assembly the ghost dofs of the iterate are refreshed from their owners (halo exchange);
* each rank assembles own + ghost cells: its OWNED rows of A and b are complete, no reduction of entries;
* norms: owned dofs + all-reduce;  Dirichlet rows / ident_zeros act on local rows (ghost rows are unused);
* the problem files keep seeing GLOBAL objects: `to_host` returns the gathered global vector (drag / lift /
  point evaluation hooks), `to_device` accepts a global vector (initial states) or the global BC value list.

The linear solve does not shard (BASELINE north_star: reported separately): `GatheredLU` sums the owned rows of
all ranks into the global CSR arrays, rank 0 factorises / solves on the host, the increment is broadcast.
"""
import numpy as np
import torch
import torch.distributed as dist


class DistributedAssembler:
    def __init__(self, view, local, gspace, halo=None):
        """view: RankView; local: assembler on (view.mesh, view.space); gspace: the global MixedSpace;
        halo: object with .exchange(local_vector) (HaloExchangeRCB on GPUs) or None = exchange through the
        host with torch.distributed (CPU tests)."""
        self.view, self.local, self.gspace = view, local, gspace
        self.rank, self.world = view.rank, view.world
        self.halo = halo
        self.num_dofs = gspace.num_dofs                      # GLOBAL sizes, as the problem files expect
        self.rowptr, self.colidx = gspace.pattern()
        self.nnz = len(self.colidx)
        self._send, self._recv = view.halo_lists()
        self._owned = np.nonzero(view.dof_owned)[0]
        self._owned_global = view.dof_global[self._owned]
        self._glob2loc = np.full(gspace.num_dofs, -1, dtype=np.int64)
        self._glob2loc[view.dof_global] = np.arange(len(view.dof_global))
        self._nbc_global = -1
        self._bc_sel = np.zeros(0, dtype=np.int64)
        self._entry_pos = None
        self.bc_dofs = None

    # ---- vectors -------------------------------------------------------------------------------------------
    def new_vector(self):
        return self.local.new_vector()

    def new_matrix(self):
        return self.local.new_matrix()

    def to_device(self, a):
        a = np.asarray(a, dtype=np.float64)
        if a.shape[0] == self.num_dofs:                      # global state vector -> local (owned + ghost)
            return self.local.to_device(a[self.view.dof_global])
        if a.shape[0] == self._nbc_global:                   # global BC value list -> values of the local BC dofs
            return self.local.to_device(a[self._bc_sel])
        raise ValueError("to_device expects a global vector or the global BC value list")

    def to_host(self, v):
        """gathered GLOBAL vector (every rank gets it): owners contribute their dofs."""
        loc = np.asarray(self.local.to_host(v))
        out = torch.zeros(self.num_dofs, dtype=torch.float64)
        out[torch.as_tensor(self._owned_global)] = torch.as_tensor(loc[self._owned])
        return self._allreduce(out).numpy()

    def copy(self, dst, src):
        self.local.copy(dst, src)

    def axpy(self, x, alpha, y):
        self.local.axpy(x, alpha, y)

    def _allreduce(self, t):
        if dist.get_backend() == "nccl":
            d = t.cuda()
            dist.all_reduce(d)
            return d.cpu()
        dist.all_reduce(t)
        return t

    def norm(self, v):
        loc = np.asarray(self.local.to_host(v))
        ss = torch.tensor([float((loc[self._owned] ** 2).sum())], dtype=torch.float64)
        return float(np.sqrt(self._allreduce(ss).item()))

    def refresh_ghosts(self, x):
        """ghost dofs of x <- their owners' values."""
        if self.halo is not None:
            self.halo.exchange(x)
            return
        xh = torch.as_tensor(np.array(self.local.to_host(x), dtype=np.float64))
        ops, rbuf = [], {}
        sbuf = {p: xh[torch.as_tensor(idx.astype(np.int64))].contiguous() for p, idx in self._send.items()}
        for p in sorted(sbuf):
            ops.append(dist.P2POp(dist.isend, sbuf[p], p))
        for p in sorted(self._recv):
            rbuf[p] = torch.empty(len(self._recv[p]), dtype=torch.float64)
            ops.append(dist.P2POp(dist.irecv, rbuf[p], p))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for p, idx in self._recv.items():
            xh[torch.as_tensor(idx.astype(np.int64))] = rbuf[p]
        self.local.copy(x, self.local.to_device(xh.numpy()))

    # ---- assembly ------------------------------------------------------------------------------------------
    def assemble_A_pre(self, x, x1, out=None):
        self.refresh_ghosts(x)
        self.refresh_ghosts(x1)
        return self.local.assemble_A_pre(x, x1) if out is None else self.local.assemble_A_pre(x, x1, out)

    def compress_A_pre(self, A_pre):
        if hasattr(self.local, "compress_A_pre"):
            return self.local.compress_A_pre(A_pre)

    def assemble_jacobian(self, x, x1, A_pre=None, out=None):
        self.refresh_ghosts(x)
        return self.local.assemble_jacobian(x, x1, A_pre, out)

    def assemble_residual(self, x, x1, out=None):
        self.refresh_ghosts(x)
        return self.local.assemble_residual(x, x1) if out is None else self.local.assemble_residual(x, x1, out)

    def ident_zeros(self, A):
        self.local.ident_zeros(A)

    # ---- boundary conditions, facet terms (problem files hand over GLOBAL dof numbers) -----------------------
    def set_bc_dofs(self, dofs):
        dofs = np.asarray(dofs, dtype=np.int64)
        self.bc_dofs = dofs.astype(np.int32)                 # the GLOBAL set (newtonsolver checks it every step)
        self._nbc_global = len(dofs)
        loc = self._glob2loc[dofs]
        self._bc_sel = np.nonzero(loc >= 0)[0]
        self.local.set_bc_dofs(loc[self._bc_sel].astype(np.int32))

    def apply_bc_matrix(self, A):
        self.local.apply_bc_matrix(A)

    def apply_bc_rhs(self, b, x, g):
        self.local.apply_bc_rhs(b, x, g)

    def apply_bc_vec(self, x, g):
        self.local.apply_bc_vec(x, g)

    def _local_coo(self, rows, cols, vals):
        """entries of a global COO operator whose ROW is owned here (their cells are local), in local numbering."""
        r = self._glob2loc[np.asarray(rows, dtype=np.int64)]
        keep = (r >= 0)
        keep[keep] = self.view.dof_owned[r[keep]]
        c = self._glob2loc[np.asarray(cols, dtype=np.int64)[keep]]
        assert (c >= 0).all(), "facet-operator column outside the local ghost layer"
        return r[keep].astype(np.int32), c.astype(np.int32), np.asarray(vals)[keep]

    def set_facet_operator(self, rows, cols, vals):
        self.local.set_facet_operator(*self._local_coo(rows, cols, vals))

    def set_facet_operator_prev(self, rows, cols, vals):
        self.local.set_facet_operator_prev(*self._local_coo(rows, cols, vals))

    def set_facet_load(self, dofs, vals, scale=1.0):
        d = self._glob2loc[np.asarray(dofs, dtype=np.int64)]
        keep = d >= 0
        keep[keep] = self.view.dof_owned[d[keep]]
        self.local.set_facet_load(d[keep].astype(np.int32), np.asarray(vals)[keep], scale)

    def set_facet_load_scale(self, scale):
        self.local.set_facet_load_scale(scale)

    # ---- gather for the (non-sharded) linear solve ----------------------------------------------------------
    def entry_positions(self):
        """for the local CSR entries of OWNED rows: (local entry index, position in the global CSR arrays)."""
        if self._entry_pos is None:
            lrp = np.asarray(self.local.rowptr.cpu() if hasattr(self.local.rowptr, "cpu") else self.local.rowptr).astype(np.int64)
            lci = np.asarray(self.local.colidx.cpu() if hasattr(self.local.colidx, "cpu") else self.local.colidx).astype(np.int64)
            cnt = np.diff(lrp)[self._owned]
            loc_idx = np.concatenate([np.arange(lrp[r], lrp[r + 1]) for r in self._owned]) if len(self._owned) else np.zeros(0, dtype=np.int64)
            grow = np.repeat(self._owned_global, cnt)
            gcol = self.view.dof_global[lci[loc_idx]]
            N = self.num_dofs
            gkeys = np.repeat(np.arange(N, dtype=np.int64), np.diff(self.rowptr)) * N + np.asarray(self.colidx, dtype=np.int64)
            pos = np.searchsorted(gkeys, grow * N + gcol)
            assert np.array_equal(gkeys[pos], grow * N + gcol)
            self._entry_pos = (loc_idx, pos)
        return self._entry_pos

    def gather_matrix(self, A):
        """global CSR values on every rank: owned rows of all ranks summed into place."""
        loc_idx, pos = self.entry_positions()
        vals = np.asarray(self.local.to_host(A))
        out = torch.zeros(self.nnz, dtype=torch.float64)
        out[torch.as_tensor(pos)] = torch.as_tensor(vals[loc_idx])
        return self._allreduce(out).numpy()


class GatheredLU:
    """up_sol of the distributed run: set_operator(A_local) / solve(b_local) -> dx_local. Rank 0 holds the host LU
    (same re-use protocol as turtlefsi_b200.linear_solver.HostLU); the increment is broadcast."""

    def __init__(self, dasm):
        from .linear_solver import HostLU
        self.d = dasm
        self.lu = HostLU(dasm.rowptr, dasm.colidx, dasm.num_dofs) if dasm.rank == 0 else None
        self.factorizations = 0

    def set_operator(self, A):
        vals = self.d.gather_matrix(A)
        if self.lu is not None:
            self.lu.set_operator(vals)

    def solve(self, b):
        bg = self.d.to_host(b)
        if self.lu is not None:
            dx = torch.as_tensor(np.asarray(self.lu.solve(bg), dtype=np.float64))
            self.factorizations = self.lu.factorizations
        else:
            dx = torch.zeros(self.d.num_dofs, dtype=torch.float64)
        if dist.get_backend() == "nccl":
            dd = dx.cuda()
            dist.broadcast(dd, 0)
            dx = dd.cpu()
        else:
            dist.broadcast(dx, 0)
        return self.d.to_device(dx.numpy())


def make_distributed(mesh, gspace, make_local, biharmonic=False, use_gpu_halo=False):
    """RankView of this process + the wrapped local assembler. make_local(view) -> local assembler."""
    from .partition_rcb import HaloExchangeRCB, RankView
    rank, world = dist.get_rank(), dist.get_world_size()
    view = RankView(mesh, rank, world, biharmonic=biharmonic, gspace=gspace)
    local = make_local(view)
    halo = HaloExchangeRCB(local, *view.halo_lists(), rank) if use_gpu_halo else None
    return DistributedAssembler(view, local, gspace, halo)
