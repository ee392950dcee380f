"""Host logic of the multi-GPU path (CPU): slab partition, node ownership, halo index lists, and a
world_size-2 gloo run of the halo exchange protocol."""
import os
import sys

import numpy as np
import pytest

from turtlefsi_b200.partition import (LocalSpaceGlobalKeys, SlabMesh, halo_lists, synthetic_state_global)
from turtlefsi_b200.space import MixedSpace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _local(n, r, w, nz=None):
    slab = SlabMesh(n, r, w, nz=nz)
    sp = MixedSpace(slab.mesh)
    return slab, sp, LocalSpaceGlobalKeys(slab, sp)


@pytest.mark.parametrize("world", [2, 3])
def test_partition_covers_global_mesh_and_owners_are_unique(world):
    n = 3
    glob = _local(n, 0, 1, nz=n * world)
    parts = [_local(n, r, world) for r in range(world)]
    assert sum(int(p[0].cell_owned.sum()) for p in parts) == glob[0].mesh.num_cells
    # every global node is owned by exactly one rank, and ranks agree on the owner
    owners = {}
    for r, (slab, sp, gk) in enumerate(parts):
        for k, o in zip(gk.key, gk.owner):
            assert owners.setdefault(int(k), int(o)) == int(o)
    assert set(owners) == set(int(k) for k in glob[2].key)
    for r, (slab, sp, gk) in enumerate(parts):
        mine = set(int(k) for k in gk.key[gk.owned])
        assert mine == {k for k, o in owners.items() if o == r}
        # every cell touching an owned node is present locally (owned rows complete without reduction)
        cells_here = {tuple(sorted(slab.global_vertex[c])) for c in slab.mesh.cells}
        gslab, gsp, ggk = glob
        touch = np.isin(ggk.key[gsp.cell_nodes], list(mine)).any(axis=1)
        for c in gslab.mesh.cells[touch]:
            assert tuple(sorted(gslab.global_vertex[c])) in cells_here


def test_halo_lists_pair_up_and_states_agree():
    world, n = 3, 3
    parts = [_local(n, r, world) for r in range(world)]
    lists = [halo_lists(*p) for p in parts]
    states = [synthetic_state_global(p[0], p[1], p[2], 2024, 0.3) for p in parts]

    def keys(p, dofs):          # (global node key, component) of local dofs
        slab, sp, gk = p
        node = np.searchsorted(sp.node_base, dofs, side="right") - 1
        return np.stack([gk.key[node], dofs - sp.node_base[node]], axis=1)
    for r in range(world - 1):
        up, below = lists[r]["send_up"], lists[r + 1]["recv_below"]
        down, above = lists[r + 1]["send_down"], lists[r]["recv_above"]
        assert np.array_equal(keys(parts[r], up), keys(parts[r + 1], below))
        assert np.array_equal(keys(parts[r + 1], down), keys(parts[r], above))
        # the synthetic state is a function of global keys: ghosts already agree with their owners
        assert np.array_equal(states[r][up], states[r + 1][below])
        assert np.array_equal(states[r + 1][down], states[r][above])
    # ghosts = exactly the non-owned dofs
    for r, p in enumerate(parts):
        slab, sp, gk = p
        ghost = np.concatenate([l for k, l in lists[r].items() if k.startswith("recv") and l is not None])
        node = np.searchsorted(sp.node_base, ghost, side="right") - 1
        assert set(node) == set(np.nonzero(~gk.owned)[0])


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from turtlefsi_b200.partition import LocalSpaceGlobalKeys, SlabMesh, halo_lists, synthetic_state_global
from turtlefsi_b200.space import MixedSpace
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
slab = SlabMesh(3, rank, world); sp = MixedSpace(slab.mesh); gk = LocalSpaceGlobalKeys(slab, sp)
lists = halo_lists(slab, sp, gk)
truth = synthetic_state_global(slab, sp, gk, 7, 0.3)
x = torch.as_tensor(truth.copy())
ghost = np.concatenate([l for k, l in lists.items() if k.startswith("recv") and l is not None])
x[torch.as_tensor(ghost.astype(np.int64))] = -1.0          # poison the ghosts
ops, bufs = [], {}
for key, peer in (("send_up", rank + 1), ("send_down", rank - 1)):
    if lists[key] is not None:
        bufs[key] = x[torch.as_tensor(lists[key].astype(np.int64))].contiguous()
        ops.append(dist.P2POp(dist.isend, bufs[key], peer))
for key, peer in (("recv_below", rank - 1), ("recv_above", rank + 1)):
    if lists[key] is not None:
        bufs[key] = torch.empty(len(lists[key]), dtype=torch.float64)
        ops.append(dist.P2POp(dist.irecv, bufs[key], peer))
for req in dist.batch_isend_irecv(ops):
    req.wait()
for key in ("recv_below", "recv_above"):
    if lists[key] is not None:
        x[torch.as_tensor(lists[key].astype(np.int64))] = bufs[key]
assert np.array_equal(x.numpy(), truth), "halo exchange did not restore the ghost values"
# distributed norm: owned dofs only + all-reduce == the norm of the global vector
owned_dofs = np.repeat(gk.owned, sp.node_ncomp)
ss = torch.tensor([float((truth[owned_dofs] ** 2).sum())], dtype=torch.float64)
dist.all_reduce(ss)
if rank == 0:
    g = SlabMesh(3, 0, 1, nz=3 * world); gsp = MixedSpace(g.mesh); ggk = LocalSpaceGlobalKeys(g, gsp)
    ref = float((synthetic_state_global(g, gsp, ggk, 7, 0.3) ** 2).sum())
    assert abs(ss.item() - ref) < 1e-12 * ref, (ss.item(), ref)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_halo_exchange_gloo_world2(tmp_path):
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


# ---- general meshes: recursive-coordinate-bisection partition (turtlefsi_b200/partition_rcb.py) ----------------

def _rcb_views(mesh, world, biharmonic=False):
    from turtlefsi_b200.partition_rcb import RankView, rcb_partition
    part = rcb_partition(mesh, world)
    gs = MixedSpace(mesh, biharmonic=biharmonic)
    return part, gs, [RankView(mesh, r, world, biharmonic=biharmonic, cell_part=part, gspace=gs) for r in range(world)]


@pytest.mark.parametrize("world", [2, 3, 5])
def test_rcb_partition_ownership_and_halo_lists(world):
    from turtlefsi_b200.mesh import tube_mesh
    mesh = tube_mesh(4)
    part, gs, views = _rcb_views(mesh, world)
    counts = np.bincount(part, minlength=world)
    assert counts.max() - counts.min() <= 1 and counts.sum() == mesh.num_cells       # balanced to one cell
    owned = np.zeros(gs.num_dofs, dtype=int)
    for v in views:
        owned[v.dof_global[v.dof_owned]] += 1
        # every cell touching an owned node is present locally, and cells keep their vertex order
        touch = (v.node_owner_global[gs.cell_nodes] == v.rank).any(axis=1)
        assert set(np.nonzero(touch)[0]) <= set(v.cells_global)
        assert np.array_equal(v.vertex_global[v.mesh.cells], mesh.cells[v.cells_global])
    assert (owned == 1).all()                                                          # every dof has ONE owner
    lists = [v.halo_lists() for v in views]
    for r, v in enumerate(views):
        send, recv = lists[r]
        for p, idx in send.items():
            assert np.array_equal(v.dof_global[idx], views[p].dof_global[lists[p][1][r]])   # pair up, same order
        ghost = np.concatenate(list(recv.values())) if recv else np.zeros(0, dtype=int)
        assert set(ghost) == set(np.nonzero(~v.dof_owned)[0])                          # ghosts = the non-owned dofs


def test_rcb_partitioned_oracle_assembly_equals_global_on_owned_rows():
    """owned rows of every rank's LOCAL assembly (own + ghost cells) == the rows of the single-domain assembly:
    no reduction of matrix entries is needed (2-D Turek mesh, 3 parts, biharmonic space; oracle C port)."""
    import scipy.sparse as sp
    from oracle.backend import OracleAssembler
    from tests.conftest import TUBE_PARAMS, TUBE_REGIONS
    from turtlefsi_b200.mesh import read_dolfin_xml, synthetic_state
    from turtlefsi_b200.boundary import mark_cells
    mesh = read_dolfin_xml(os.path.join(ROOT, "tests", "golden", "meshes", "TF_fsi.xml.gz"))
    mesh.markers[:] = 1
    mark_cells(mesh, lambda x: (x[:, 0] > 0.24) & (x[:, 0] < 0.61) & (np.abs(x[:, 1] - 0.2) < 0.011), 2)   # the flag
    assert (mesh.markers == 2).sum() > 50
    common = dict(TUBE_PARAMS, extrapolation="biharmonic", hmin=mesh.hmin())
    part, gs, views = _rcb_views(mesh, 3, biharmonic=True)
    xg, x1g = synthetic_state(mesh, gs, 11, h=0.01), synthetic_state(mesh, gs, 12, h=0.01)
    G = OracleAssembler(mesh, gs, TUBE_REGIONS, common, 5, 10, use_c_port=True)
    GA = sp.csr_matrix((G.assemble_jacobian(xg, x1g, G.assemble_A_pre(xg, x1g)), G.colidx, G.rowptr), shape=(gs.num_dofs,) * 2)
    gb = G.assemble_residual(xg, x1g)
    scale, covered = abs(GA).max(), 0
    for v in views:
        L = OracleAssembler(v.mesh, v.space, TUBE_REGIONS, common, 5, 10, use_c_port=True)
        x, x1 = v.scatter_global(xg), v.scatter_global(x1g)
        A = sp.csr_matrix((L.assemble_jacobian(x, x1, L.assemble_A_pre(x, x1)), L.colidx, L.rowptr), shape=(v.space.num_dofs,) * 2)
        b = L.assemble_residual(x, x1)
        own = np.nonzero(v.dof_owned)[0]
        Ao = A[own].tocoo()
        lifted = sp.csr_matrix((Ao.data, (v.dof_global[own][Ao.row], v.dof_global[Ao.col])), shape=GA.shape)
        rows = np.zeros(gs.num_dofs, dtype=bool)
        rows[v.dof_global[own]] = True
        D = (lifted - sp.diags(rows.astype(float)) @ GA)
        assert abs(D).max() < 1e-12 * scale
        assert np.abs(b[own] - gb[v.dof_global[own]]).max() < 1e-12 * np.abs(gb).max()
        covered += len(own)
    assert covered == gs.num_dofs


RCB_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from turtlefsi_b200.mesh import tube_mesh, synthetic_state
from turtlefsi_b200.partition_rcb import RankView
from turtlefsi_b200.space import MixedSpace
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mesh = tube_mesh(4)
view = RankView(mesh, rank, world)
truth = view.scatter_global(synthetic_state(mesh, view.gspace, 7, h=0.3))
send, recv = view.halo_lists()
x = torch.as_tensor(truth.copy())
x[torch.as_tensor(np.nonzero(~view.dof_owned)[0])] = -1.0                 # poison the ghosts
ops, rbuf = [], {}
sbuf = {p: x[torch.as_tensor(idx.astype(np.int64))].contiguous() for p, idx in send.items()}
for p in sorted(send):
    ops.append(dist.P2POp(dist.isend, sbuf[p], p))
for p in sorted(recv):
    rbuf[p] = torch.empty(len(recv[p]), dtype=torch.float64)
    ops.append(dist.P2POp(dist.irecv, rbuf[p], p))
for req in dist.batch_isend_irecv(ops):
    req.wait()
for p, idx in recv.items():
    x[torch.as_tensor(idx.astype(np.int64))] = rbuf[p]
assert np.array_equal(x.numpy(), truth), "halo exchange did not restore the ghost values"
ss = torch.tensor([float((truth[view.dof_owned] ** 2).sum())], dtype=torch.float64)
dist.all_reduce(ss)
ref = float((synthetic_state(mesh, view.gspace, 7, h=0.3) ** 2).sum())
assert abs(ss.item() - ref) < 1e-12 * ref
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_rcb_halo_exchange_gloo_world2(tmp_path):
    import subprocess
    script = tmp_path / "rcb_worker.py"
    script.write_text(RCB_WORKER % ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


@pytest.mark.parametrize("gdim,world", [(2, 3), (3, 4)])
def test_rcb_partition_on_unstructured_delaunay_meshes(gdim, world):
    """arbitrary (Delaunay) meshes: ownership, halo pairing, and owned rows of the partitioned oracle assembly
    equal to the single-domain rows."""
    import scipy.sparse as sp
    from scipy.spatial import Delaunay
    from oracle.backend import OracleAssembler
    from tests.conftest import TUBE_PARAMS, TUBE_REGIONS
    from turtlefsi_b200.mesh import Mesh, synthetic_state
    rng = np.random.default_rng(5 + gdim)
    pts = rng.uniform(0, 1, size=(70 if gdim == 2 else 45, gdim))
    tri = Delaunay(pts)
    cells = np.sort(tri.simplices, axis=1).astype(np.int32)
    X = pts[cells]
    vol = np.abs(np.linalg.det(X[:, 1:] - X[:, :1]))
    cells = cells[vol > 1e-6]                                                # drop degenerate slivers
    cen = pts[cells].mean(axis=1)
    mesh = Mesh(pts, cells, np.where(cen[:, 0] < 0.5, 1, 2).astype(np.int32))
    part, gs, views = _rcb_views(mesh, world)
    owned = np.zeros(gs.num_dofs, dtype=int)
    for v in views:
        owned[v.dof_global[v.dof_owned]] += 1
    assert (owned == 1).all()
    lists = [v.halo_lists() for v in views]
    for r, v in enumerate(views):
        for p, idx in lists[r][0].items():
            assert np.array_equal(v.dof_global[idx], views[p].dof_global[lists[p][1][r]])
    common = dict(TUBE_PARAMS, hmin=mesh.hmin())
    degJ, degR = (5, 10) if gdim == 2 else (4, 6)
    xg, x1g = synthetic_state(mesh, gs, 21, h=0.05), synthetic_state(mesh, gs, 22, h=0.05)
    G = OracleAssembler(mesh, gs, TUBE_REGIONS, common, degJ, degR, use_c_port=True)
    GA = sp.csr_matrix((G.assemble_jacobian(xg, x1g, G.assemble_A_pre(xg, x1g)), G.colidx, G.rowptr), shape=(gs.num_dofs,) * 2)
    gb = G.assemble_residual(xg, x1g)
    for v in views:
        L = OracleAssembler(v.mesh, v.space, TUBE_REGIONS, common, degJ, degR, use_c_port=True)
        x, x1 = v.scatter_global(xg), v.scatter_global(x1g)
        A = sp.csr_matrix((L.assemble_jacobian(x, x1, L.assemble_A_pre(x, x1)), L.colidx, L.rowptr), shape=(v.space.num_dofs,) * 2)
        own = np.nonzero(v.dof_owned)[0]
        Ao = A[own].tocoo()
        lifted = sp.csr_matrix((Ao.data, (v.dof_global[own][Ao.row], v.dof_global[Ao.col])), shape=GA.shape)
        rows = np.zeros(gs.num_dofs)
        rows[v.dof_global[own]] = 1.0
        assert abs(lifted - sp.diags(rows) @ GA).max() < 1e-12 * abs(GA).max()
        assert np.abs(L.assemble_residual(x, x1)[own] - gb[v.dof_global[own]]).max() < 1e-12 * np.abs(gb).max()
