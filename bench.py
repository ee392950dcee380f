#!/usr/bin/env python
"""bench.py — monolithic FSI Jacobian assembly throughput (cells/s) on B200.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...   the CPU restatement on host cores

One "step" = one pass of the hot path over the synthetic 3-D tube mesh: the call that replaces
`assemble(J_nonlinear, tensor=A); A.axpy(1.0, A_pre, True)` of turtleFSI's Newton solver
(turtleFSI/modules/newtonsolver.py:75-79), i.e. `tfsi_assemble_jacobian` (A := A_pre + J_nl(x)) into the solver's
persistent matrix (`A = Matrix(A_pre)`, newtonsolver.py:28 -> tfsi_bind_matrix).

Workload (BASELINE.json configs 4 / 5):
  N = 1      the 6*ncube^3-tet tube (ncube 69 -> 1 971 054 tets) on one GPU;
  N > 1      default `--scaling strong`: the SAME ONE tube cut into N cell parts (recursive coordinate bisection,
             turtlefsi_b200/partition_rcb.py), one part + its ghost cells per GPU, ghost dofs of the Newton iterate
             exchanged over NCCL before every assembly (inside the timed step);
             `--scaling weak`: 6*ncube^3 tets PER GPU (z-slabs of a tube that grows with N) — the r1 configuration,
             also measured (briefly) after the strong run and reported under "weak_scaling";
  --sweep    config 5: a list of per-GPU sizes (weak), one entry each under "sweep".
`value` = cells assembled per second with every input resident in HBM; `e2e` = the same metric through the
host-buffer C-ABI call (H2D of the Newton iterate, Jacobian + ident_zeros + BC rows, residual + BC, D2H of the
residual vector). Every run ends with an in-run PARITY check outside the timed region: ghost dofs are poisoned, the
real halo exchange is run, matrix and residual are assembled and a few OWNED rows per rank are recomputed by the
numpy complex-step oracle (oracle/spotcheck.py — the checker, never the thing measured); rc != 0 on failure.

The reference's own FEniCS/MPI assembly CANNOT be run (no DOLFIN in the image): the CPU arm is the plain-C
restatement under oracle/ ("port"), timed on the box's host cores on a bounded sample of the same mesh.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# parameters of turtleFSI/problems/Test_Cylinder/problem_aneu.py:23-47
DT, THETA = 1e-3, 0.501
REGIONS = [dict(marker=1, kind="fluid", rho=1025.0, mu=3.5e-3),
           dict(marker=2, kind="solid", rho=1000.0, mu=344827.5862068966, lambda_=3103448.2758620703)]
COMMON = dict(dt=DT, theta=THETA, extrapolation="laplace", extrapolation_sub_type="constant")
QDEG_J, QDEG_R = 4, 6

# Algorithmic FLOPs per cell of the Jacobian kernels (mul/add = 1, fma = 2; C' tensors counted once
# per block, no credit for the half-block recomputation) — derivation in DESIGN.md "FLOP model".
FLOPS_FLUID_PER_Q = 15598
FLOPS_SOLID_PER_Q = 4895
NQ_J = 14
METRIC = "monolithic Jacobian assembly cells/s"


def make_config(ncube, world, scaling):
    """`config` of the JSON line — a function of the command line only, so that both arms print the same one."""
    ncell = 6 * ncube ** 3
    if world == 1:
        par = "single GPU"
        mesh = "ONE tube mesh of 6*%d^3 = %d tets" % (ncube, ncell)
    elif scaling == "strong":
        par = ("cells of the ONE mesh partitioned across %d GPUs (recursive coordinate bisection, owned + ghost cells, "
               "owned CSR rows per GPU), ghost-dof halo exchange of the Newton iterate over NCCL inside every step" % world)
        mesh = "ONE tube mesh of 6*%d^3 = %d tets" % (ncube, ncell)
    else:
        par = "z-slabs of a tube that grows with the GPU count, ghost-dof halo exchange per step"
        mesh = "tube mesh of 6*%d^3 = %d tets PER GPU (%d GPUs)" % (ncube, ncell, world)
    return {
        "workload": "synthetic 3-D elastic-tube pressure-wave FSI (Test_Cylinder parameters): %s, P2-P2-P1 monolithic "
                    "(64 dofs per tet), laplace/constant lifting, Jacobian quadrature degree 4 (Keast 14 pts)" % mesh,
        "step": "A := A_pre + J_nl(x) into the fixed CSR pattern (newtonsolver.py:75-79)",
        "scaling_mode": "single" if world == 1 else scaling,
        "parallelism": par,
        "l2": "inputs larger than L2 (A alone is 24 GB at 2 M tets; the smallest part, 250 k tets, writes 3 GB per step)",
    }


def make_assembler(mesh, space, device, hmin, deterministic=False):
    from turtlefsi_b200.assembler import FSIAssembler
    return FSIAssembler(mesh, space, REGIONS, DT, THETA, extrapolation="laplace", extrapolation_sub_type="constant",
                        quad_degree_jacobian=QDEG_J, quad_degree_residual=QDEG_R, device=device, verify_pattern=0,
                        hmin=hmin, deterministic=deterministic)


def tube_bc(space, mesh, length):
    """synthetic Dirichlet set: d and v on the tube ends and the outer wall"""
    X = space.node_coords(mesh)
    r = np.hypot(X[:, 0], X[:, 1])
    on_bnd = (X[:, 2] < 1e-9) | (X[:, 2] > length - 1e-9) | (r > 0.999)
    return np.concatenate([space.node_dof(f, a)[on_bnd] for f in (0, 1) for a in range(3)]).astype(np.int32)


def build_workload(n, device, rank=0, world=1, verify=0, deterministic=False):
    """single-GPU workload (also used by tests/test_gpu_fullsize.py)."""
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    t0 = time.time()
    mesh = tube_mesh(n, seed=1234 + rank)
    space = MixedSpace(mesh)
    from turtlefsi_b200.assembler import FSIAssembler
    asm = FSIAssembler(mesh, space, REGIONS, DT, THETA, extrapolation="laplace",
                       extrapolation_sub_type="constant", quad_degree_jacobian=QDEG_J,
                       quad_degree_residual=QDEG_R, device=device, verify_pattern=verify, hmin=2.0 / n,
                       deterministic=deterministic)
    x = synthetic_state(mesh, space, 2024 + rank, h=2.0 / n)
    x1 = synthetic_state(mesh, space, 2025 + rank, h=2.0 / n)
    bc = tube_bc(space, mesh, 4.0)
    asm.set_bc_dofs(bc)
    return mesh, space, asm, x, x1, bc, time.time() - t0


class Workload:
    """everything one rank needs for the timed loop and the parity check."""

    def __init__(self, mesh, space, asm, x, x1, bc, halo, owned_cells, owned_nodes, ghost_dofs, setup_s, kind):
        self.mesh, self.space, self.asm, self.x, self.x1, self.bc, self.halo = mesh, space, asm, x, x1, bc, halo
        self.owned_cells, self.owned_nodes, self.ghost_dofs = int(owned_cells), owned_nodes, ghost_dofs
        self.setup_s, self.kind = setup_s, kind


def workload_single(n, device, deterministic=False):
    mesh, space, asm, x, x1, bc, t = build_workload(n, device, deterministic=deterministic)
    return Workload(mesh, space, asm, x, x1, bc, None, mesh.num_cells, np.ones(space.num_nodes, dtype=bool),
                    np.zeros(0, dtype=np.int64), t, "single")


def workload_strong(n, device, rank, world):
    """ONE 6 n^3 tube, cut into `world` parts by RCB; this rank's part + ghost cells."""
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.partition_rcb import HaloExchangeRCB, RankView, rcb_partition
    from turtlefsi_b200.space import MixedSpace
    t0 = time.time()
    gmesh = tube_mesh(n, seed=1234)
    gspace = MixedSpace(gmesh)
    part = rcb_partition(gmesh, world, metric="cells")
    view = RankView(gmesh, rank, world, cell_part=part, gspace=gspace, balanced=True)
    asm = make_assembler(view.mesh, view.space, device, 2.0 / n)
    x = view.scatter_global(synthetic_state(gmesh, gspace, 2024, h=2.0 / n))
    x1 = view.scatter_global(synthetic_state(gmesh, gspace, 2025, h=2.0 / n))
    bc = tube_bc(view.space, view.mesh, 4.0)
    asm.set_bc_dofs(bc)
    send, recv = view.halo_lists()
    halo = HaloExchangeRCB(asm, send, recv, rank)
    ghost = np.nonzero(~view.dof_owned)[0]
    w = Workload(view.mesh, view.space, asm, x, x1, bc, halo, view.cell_owned.sum(), view.node_owned, ghost,
                 time.time() - t0, "strong")
    w.ghost_cell_fraction = float(view.mesh.num_cells / max(view.cell_owned.sum(), 1) - 1.0)
    w.peers = sorted(set(send) | set(recv))
    w.halo_doubles = int(sum(len(v) for v in recv.values()))
    return w


def workload_weak(n, device, rank, world):
    from turtlefsi_b200.partition import build_partitioned_workload
    mesh, space, asm, x, x1, bc, halo, setup_s = build_partitioned_workload(n, device, rank, world, REGIONS, DT, THETA,
                                                                            QDEG_J, QDEG_R)
    from turtlefsi_b200.partition import LocalSpaceGlobalKeys  # noqa: F401  (ownership is recomputed below)
    owned_nodes = halo.owned_node_mask
    ghost = halo.ghost_dofs
    w = Workload(mesh, space, asm, x, x1, bc, halo, halo.num_owned_cells, owned_nodes, ghost, setup_s, "weak")
    w.ghost_cell_fraction = float(mesh.num_cells / max(halo.num_owned_cells, 1) - 1.0)
    return w


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in names.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def pin_to_gpu_numa(local_rank):
    """bind this rank's host threads (and with them its pinned buffers) to the CPUs NVML lists for its GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


# ------------------------------------------------------------------------------------------------------
# CPU arm
def cpu_sample(ncube, max_cells):
    """bounded sample of the bench mesh: the first `max_cells` cells (Morton order: one compact blob) of the SAME
    6*ncube^3 tube, as a mesh of their own (vertices renumbered monotonically)."""
    from turtlefsi_b200.mesh import Mesh, tube_mesh
    mesh = tube_mesh(ncube, seed=1234)
    if mesh.num_cells <= max_cells:
        return mesh, mesh.num_cells
    # a blob that holds fluid and solid cells in the mesh's own proportion: Morton blocks are compact, so take the
    # cells whose centroid lies in an axial slice of the tube
    cen = mesh.coords[mesh.cells].mean(axis=1)
    order = np.argsort(cen[:, 2], kind="stable")
    sel = np.sort(order[:max_cells])
    cells = mesh.cells[sel]
    verts = np.unique(cells)
    sub = Mesh(mesh.coords[verts], np.searchsorted(verts, cells).astype(np.int32), mesh.markers[sel].copy(),
               dict(kind="tube-sample", n=ncube))
    return sub, mesh.num_cells


def cpu_baseline(ncube, max_cells=200000, steps=1, threads=0, warm=True):
    """the C restatement (oracle 'port') on a bounded sample of the bench mesh. returns (cells, times, cores, text)."""
    from oracle import cport, fem_tables as ft
    from turtlefsi_b200.mesh import synthetic_state
    from turtlefsi_b200.space import MixedSpace
    cport.build_native()
    mesh, ntotal = cpu_sample(ncube, max_cells)
    space = MixedSpace(mesh)
    pattern = space.pattern()
    x = synthetic_state(mesh, space, 2024, h=2.0 / ncube)
    x1 = synthetic_state(mesh, space, 2025, h=2.0 / ncube)
    A = np.zeros(len(pattern[1]))
    rule = ft.quadrature(3, QDEG_J)
    # all host cores of the box (torchrun exports OMP_NUM_THREADS=1: do not rely on OpenMP's default)
    cores = threads or len(os.sched_getaffinity(0))
    if warm:
        cport.assemble_jacobian(mesh, space, REGIONS, DT, THETA, x, x1, rule, pattern, A, c1=min(2000, mesh.num_cells), nthreads=cores)
    times = []
    for _ in range(steps):
        A[:] = 0.0
        t0 = time.perf_counter()
        cport.assemble_jacobian(mesh, space, REGIONS, DT, THETA, x, x1, rule, pattern, A, nthreads=cores)
        times.append(time.perf_counter() - t0)
    sample = ("%d of the %d tets of the 6*%d^3 tube (an axial slice: %d fluid + %d solid), same parameters / state "
              "distribution / quadrature, J_nl scattered into the slice's own CSR pattern; plain-C port (FFC-style dense "
              "element tensor + binary-search insertion), OpenMP, -O3 -march=native" % (
                  mesh.num_cells, ntotal, ncube, int((mesh.markers == 1).sum()), int((mesh.markers == 2).sum())))
    return mesh.num_cells, times, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scaling = "strong" if args.scaling == "auto" else args.scaling
    nw = max(args.warmup, 0)
    ncells, times, cores, sample = cpu_baseline(args.ncube, args.cpu_cells, steps=nw + max(args.steps, 1))
    times = times[nw:]
    total = sum(times)
    value = ncells * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "cells/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong" if (args.gpus > 1 and scaling == "strong") else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.ncube, args.gpus, scaling),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's FEniCS/MPI assembly cannot be run here (no DOLFIN in the image); this is the plain-C "
                "restatement on all host cores; each step assembles a bounded sample of the config's mesh (cpu_baseline.sample)",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
def timed_steps(step, steps, warmup, device, world, dist):
    import torch
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps


def allreduce(val, op, device, world, dist):
    if world == 1:
        return float(val)
    import torch
    t = torch.tensor([float(val)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def prepare_matrices(w):
    """A_pre is assembled INTO the buffer that becomes A (no second dense copy), compressed, and A is bound as the
    solver's persistent matrix (A = Matrix(A_pre), newtonsolver.py:28)."""
    import torch
    asm = w.asm
    w.xd, w.x1d = torch.as_tensor(w.x).to(asm.device), torch.as_tensor(w.x1).to(asm.device)
    A = asm.new_matrix()
    asm.assemble_A_pre(w.xd, w.x1d, out=A)
    w.apre_nnz = asm.compress_A_pre(A)
    free, _ = torch.cuda.mem_get_info(asm.device)
    w.pair = PAIR and 8 * asm.nnz * 1.35 < free                # one more matrix + the fix-up list must fit (not at 6 M tets per GPU)
    if w.pair:
        # two matrices used alternately (the previous Jacobian is dead once the next one is assembled): the rows of the
        # idle one are re-initialised by the fluid kernel of the running assembly (TMA zero-fill between its tasks,
        # A_pre non-zeros added by REDs at the next assembly): no separate A := A_pre pass
        w.As = [A, asm.new_matrix()]
        w.bound_entries = asm.bind_matrix_pair(w.As[0], w.As[1], "compressed")
    else:
        w.As = [A]
        w.bound_entries = asm.bind_matrix(A)
    w.k = 0
    w.A = A
    w.b = asm.new_vector()
    torch.cuda.synchronize()


PAIR = True


def next_matrix(w):
    w.k += 1
    w.A = w.As[w.k % len(w.As)]
    return w.A


def make_step(w):
    asm, halo = w.asm, w.halo
    if halo is None:
        return lambda: asm.assemble_jacobian(w.xd, w.x1d, "compressed", next_matrix(w))

    def step():
        # ghost dofs of the Newton iterate over NCCL; the x-independent A := A_pre runs under the transfer
        A = next_matrix(w)
        halo.exchange(w.xd, overlap=lambda: asm.assemble_jacobian(None, None, "compressed", A, phase=1))
        asm.assemble_jacobian(w.xd, w.x1d, "compressed", A, phase=2)
    return step


def parity_check(w, device, world, dist, nodes_per_rank=4):
    """outside the timed region: poison the ghost dofs, run the REAL halo exchange, assemble, and compare a few rows
    this rank OWNS (matrix rows and residual entries) with the numpy complex-step oracle."""
    import torch
    from oracle.spotcheck import check_nodes
    asm, space, mesh = w.asm, w.space, w.mesh
    xt = w.xd.clone()
    halo_exact = True
    if w.halo is not None:
        gd = torch.as_tensor(w.ghost_dofs).to(device)
        xt[gd] = 1.0e30                                     # a missed ghost would wreck every row next to it
        w.halo.exchange(xt)
        torch.cuda.synchronize()
        halo_exact = bool(torch.equal(xt, w.xd))
    # a bound matrix keeps the unit rows that ident_zeros / bc.apply wrote into rows J_nl never touches (they are
    # idempotent in the Newton loop); for the comparison with the unconstrained oracle, A is released first: the
    # assembly below then is a full A := A_pre + J_nl
    asm.unbind_matrix()
    fused = 0
    if w.pair:
        # the pair is bound afresh (both matrices = A_pre in every row) and used A0, A1, A0: the second assembly zero-fills
        # A0 from inside its fluid kernel, the third adds the A_pre non-zeros back while it assembles — the matrix that
        # is checked went through the fused re-initialisation, with no unit rows written in between
        asm.bind_matrix_pair(w.As[0], w.As[1], "compressed")
        f0 = asm.fused_launches()
        for A in (w.As[0], w.As[1], w.As[0]):
            asm.assemble_jacobian(xt, w.x1d, "compressed", A)
        w.A = w.As[0]
        fused = asm.fused_launches() - f0
        if fused != 2:
            raise RuntimeError("parity check: the matrix pair did not take the fused path (%d fused launches)" % fused)
    else:
        asm.assemble_jacobian(xt, w.x1d, "compressed", w.A)
    asm.assemble_residual(xt, w.x1d, w.b)
    torch.cuda.synchronize()
    rng = np.random.default_rng(100 + (0 if world == 1 else dist.get_rank()))
    owned = np.nonzero(w.owned_nodes)[0]
    picks = list(rng.choice(owned, size=max(nodes_per_rank // 2, 1), replace=False))
    if len(w.ghost_dofs):
        # owned nodes that sit on the interface: some of the cells around them carry ghost dofs
        ghost_mask = np.zeros(space.num_dofs, dtype=bool)
        ghost_mask[w.ghost_dofs] = True
        cell_has_ghost = ghost_mask[space.cell_dofs].any(axis=1)
        iface = np.unique(space.cell_nodes[cell_has_ghost])
        iface = iface[w.owned_nodes[iface]]
        if len(iface):
            picks += list(rng.choice(iface, size=min(nodes_per_rank - len(picks), len(iface)), replace=False))
    rowptr, colidx = asm.rowptr, asm.colidx

    def get_row(rw):
        lo, hi = int(rowptr[rw].item()), int(rowptr[rw + 1].item())
        return colidx[lo:hi].cpu().numpy(), w.A[lo:hi].cpu().numpy()
    errA, errb, sb = check_nodes(mesh, space, REGIONS, COMMON, w.x, w.x1, picks, get_row, lambda rw: w.b[rw].item(),
                                 QDEG_J, QDEG_R)
    bscale = allreduce(float(w.b.abs().max().item()), None if world == 1 else dist.ReduceOp.MAX, device, world, dist)
    rel_b = errb / max(bscale, 1e-300)
    mx = None if world == 1 else dist.ReduceOp.MAX
    return {
        "max_rel_A": allreduce(errA, mx, device, world, dist),
        "max_rel_b": allreduce(rel_b, mx, device, world, dist),
        "ranks": world,
        "fused_reinitialisation_checked": bool(fused),
        "rows_checked_per_rank": int(sum(int(space.node_ncomp[n]) for n in picks)),
        "halo_exact": bool(allreduce(0.0 if halo_exact else 1.0, mx, device, world, dist) == 0.0),
        "tolerance": 1e-12,
        "how": "ghost dofs set to 1e30, real NCCL halo exchange, A and b assembled, owned rows (interior + interface "
               "nodes) recomputed by the numpy complex-step oracle; A relative to the row maximum, b to max|b|",
    }


def measure(w, args, device, world, dist, local_rank, steps, warmup, full=True):
    """times `steps` steps of the workload; returns a dict of measurements (rank-local parts reduced)."""
    import torch
    asm = w.asm
    prepare_matrices(w)
    step = make_step(w)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = asm.launch_count()
    asm.set_profiling(True)
    ms_step = timed_steps(step, steps, warmup, device, world, dist)
    launches = asm.launch_count() - l0
    kf_ms, kf_n = asm.kernel_time(0)
    ks_ms, ks_n = asm.kernel_time(1)
    asm.set_profiling(False)
    clocks = sampler.stop()
    nrun = steps + warmup
    SUM = None if world == 1 else dist.ReduceOp.SUM
    total_cells = allreduce(w.owned_cells, SUM, device, world, dist)
    out = dict(ms_step=ms_step, value=total_cells / (ms_step * 1e-3), total_cells=total_cells,
               launches=int(round(launches * steps / nrun)), clocks=clocks,
               kf_ms=kf_ms / nrun, ks_ms=ks_ms / nrun, kf_n=kf_n // nrun, ks_n=ks_n // nrun)
    if not full:
        return out

    def timed(f, k=5):
        f()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            f()
        z.record()
        torch.cuda.synchronize()
        return a.elapsed_time(z) / k
    gd = torch.zeros(len(w.bc), dtype=torch.float64, device=device)
    out["ms_res"] = timed(lambda: asm.assemble_residual(w.xd, w.x1d, w.b))
    # the same step with the masked re-initialisation in the FOREGROUND (what a single bound matrix costs)
    os.environ["TFSI_NO_FUSED_INIT"] = "1"
    asm.assemble_jacobian(w.xd, w.x1d, "compressed", w.A)
    asm.set_profiling(True)
    t_both = timed(lambda: asm.assemble_jacobian(w.xd, w.x1d, "compressed", w.A))          # foreground init + kernels
    kf1_ms, kf1_n = asm.kernel_time(0)
    ks1_ms, ks1_n = asm.kernel_time(1)
    asm.set_profiling(False)
    del os.environ["TFSI_NO_FUSED_INIT"]
    out["ms_step_foreground_init"] = t_both
    # the cell kernels WITHOUT the pair's extra work (zero-fill of the idle matrix, in-kernel mass blocks)
    out["kf_alone_ms"] = kf1_ms / max(kf1_n, 1) * max(out["kf_n"], 1)
    out["ks_alone_ms"] = ks1_ms / max(ks1_n, 1) * max(out["ks_n"], 1)
    out["ms_init"] = t_both - out["kf_alone_ms"] - out["ks_alone_ms"]
    out["ms_ident"] = timed(lambda: asm.ident_zeros(w.A))
    xs = w.xd.clone()                                   # bc.apply(x) writes the iterate: time it on a copy
    out["ms_bc"] = timed(lambda: (asm.apply_bc_matrix(w.A), asm.apply_bc_rhs(w.b, xs, gd), asm.apply_bc_vec(xs, gd)))
    del xs
    out["ms_norm"] = timed(lambda: asm.sumsq(w.b))
    if w.halo is not None:
        out["ms_halo"] = timed(lambda: w.halo.exchange(w.xd))

    # ---- e2e: host buffers through the C-ABI. A Newton solve re-uses the previous time level: x1 is uploaded once
    # per time step (every 4th call here), x and the BC values at every call --------------------------------------
    N = w.space.num_dofs
    xh = torch.as_tensor(w.x).pin_memory()
    x1h = torch.as_tensor(w.x1).pin_memory()
    gh = torch.zeros(len(w.bc), dtype=torch.float64).pin_memory()
    bh = torch.empty(N, dtype=torch.float64).pin_memory()
    for _ in range(2):
        asm.newton_assemble_host(xh, x1h, gh, "compressed", w.A, bh, recompute=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = 8
    t0 = time.perf_counter()
    for it in range(e2e_steps):
        asm.newton_assemble_host(xh, x1h if it % 4 == 0 else None, gh, "compressed", next_matrix(w), bh, recompute=True)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    MX = None if world == 1 else dist.ReduceOp.MAX
    e2e_s = allreduce(e2e_s, MX, device, world, dist)
    out["e2e"] = {"value": total_cells / e2e_s, "unit": "cells/s",
                  "h2d_bytes_per_step": int(N * 8 + N * 8 // 4 + len(w.bc) * 8), "d2h_bytes_per_step": int(N * 8 + 8),
                  "ms_per_step": e2e_s * 1e3,
                  "what": "tfsi_newton_assemble_host on every rank, every copy inside the timed region: H2D x + BC values "
                          "(pinned) at every call, H2D x1 at every 4th call (the previous time level changes once per "
                          "time step, not per Newton iteration); A := A_pre runs under the H2D copy -> residual + BC -> "
                          "Jacobian kernels + ident_zeros + BC rows (D2H of b and ||b||^2 run under them); A stays in "
                          "HBM for the device LU. bytes are per rank"}
    return out


def roofline_of(w, m, steps):
    from turtlefsi_b200.assembler import probe_fp64_peak
    peak_tf, _ = probe_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_fluid = int((w.mesh.markers == 1).sum())
    n_solid = w.mesh.num_cells - n_fluid
    flops_fluid, flops_solid = FLOPS_FLUID_PER_Q * NQ_J, FLOPS_SOLID_PER_Q * NQ_J
    achieved_tf = flops_fluid * n_fluid / (m["kf_ms"] * 1e-3) / 1e12
    asm = w.asm
    # per step: (re)written entries of A + non-zeros of A_pre read (8 B each + 1 bit per entry) + dofs, ranks, state, coords
    alg_bytes_step = 8 * w.bound_entries + 8 * w.apre_nnz * (w.bound_entries / asm.nnz) + asm.nnz // 8 + \
        w.mesh.num_cells * (64 * 4 + 10 * 64 * 2 + 2 * 8.4 * 8 + 100)
    traffic = None
    tnote = "no ncu capture of this launch is committed"
    # the capture that matches the mode of this run: r3 = matrix pair (the fluid launch also zero-fills the idle matrix),
    # r2 = single matrix
    tp = os.path.join(ROOT, "profiles", "r4_fluid_traffic.json" if w.pair else "r2_fluid_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        traffic = t["dram_bytes_per_cell"] * n_fluid
        tnote = t["note"]
    step_flops = flops_fluid * n_fluid + flops_solid * n_solid
    return {
        "bound": "fp64", "kernel": "jac_fluid_kernel<3,14>", "achieved": achieved_tf, "peak": peak_tf,
        "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": "measured: register-resident DFMA probe (tfsi_probe_fp64_peak) in this run; "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "flops_per_cell": flops_fluid, "cells_per_launch": n_fluid, "launches_per_step": m["kf_n"],
        "ms_per_launch": m["kf_ms"] / max(m["kf_n"], 1),
        "traffic": traffic, "traffic_note": tnote,
        # matrix pair: the launch above also zero-fills the idle matrix; the same kernel on a single matrix (what the
        # separate masked A := A_pre pass of that mode then costs is breakdown_ms.init_A_pre_rows_if_foreground)
        "fluid_kernel_single_matrix": {"ms_per_launch": m.get("kf_alone_ms", 0.0) / max(m["kf_n"], 1),
                                       "achieved": flops_fluid * n_fluid / (m["kf_alone_ms"] * 1e-3) / 1e12,
                                       "frac": flops_fluid * n_fluid / (m["kf_alone_ms"] * 1e-3) / 1e12 / peak_tf}
        if m.get("kf_alone_ms") else None,
        # the same launch against SURVEY.md 8(d)'s counting model (dense sum factorisation, 23.8 kFLOP per point = 0.333
        # MFLOP per fluid tet at 14 points): the kernels integrate the same forms with 34 % fewer operations (sparse P2
        # reference gradients), so `frac` above is quoted on the SMALLER count they actually need
        "survey_model": {"flops_per_cell": 23800 * NQ_J, "achieved": 23800 * NQ_J * n_fluid / (m["kf_ms"] * 1e-3) / 1e12,
                         "frac": 23800 * NQ_J * n_fluid / (m["kf_ms"] * 1e-3) / 1e12 / peak_tf},
        "step": {"achieved": step_flops / (m["ms_step"] * 1e-3) / 1e12, "frac": step_flops / (m["ms_step"] * 1e-3) / 1e12 / peak_tf,
                 "what": "(fluid + solid algorithmic FLOPs of this rank) / whole step time incl. the re-initialisation of A (fused into the cell kernels with the matrix pair, a separate masked pass without)"},
        "hbm": {"achieved": alg_bytes_step / (m["ms_step"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes_step / (m["ms_step"] * 1e-3) / 1e9 / hbm_peak,
                "algorithmic_bytes_per_cell": alg_bytes_step / w.mesh.num_cells,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
        "solid_kernel": {"kernel": "jac_solid_kernel<3,14>", "ms_per_step": m["ks_ms"], "flops_per_cell": flops_solid,
                         "achieved": flops_solid * n_solid / (m["ks_ms"] * 1e-3) / 1e12 if m["ks_n"] else None,
                         "frac": flops_solid * n_solid / (m["ks_ms"] * 1e-3) / 1e12 / peak_tf if m["ks_n"] else None},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"],
                    help="N > 1: strong = ONE 6*ncube^3 mesh over all GPUs (default), weak = 6*ncube^3 per GPU")
    ap.add_argument("--ncube", type=int, default=int(os.environ.get("TFSI_BENCH_NCUBE", "69")),
                    help="tube mesh of 6*ncube^3 tets (69 -> 1,971,054)")
    ap.add_argument("--cpu-cells", type=int, default=200000, help="cells of the CPU arm's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1, strong: skip the additional weak-scaling measurement")
    ap.add_argument("--sweep", default="", help="comma-separated ncube list: config 5 size sweep (per-GPU sizes, weak)")
    ap.add_argument("--deterministic", action="store_true", help="colour-ordered scatter (bitwise reproducible)")
    ap.add_argument("--no-pair", action="store_true",
                    help="ONE persistent matrix re-initialised by a separate masked A := A_pre pass before every assembly, "
                         "instead of the default matrix pair (tfsi_bind_matrix_pair: the idle matrix is re-initialised by "
                         "the fluid kernel of the running assembly); the pair needs a second matrix in HBM and is "
                         "dropped automatically when it does not fit")
    ap.add_argument("--newton", type=int, default=int(os.environ.get("TFSI_BENCH_NEWTON_NCUBE", "6")),
                    help="ncube of the 3-D tube on which one full Newton step (with the device LU) is timed; 0 = skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    global PAIR
    PAIR = not args.no_pair and not args.deterministic

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: turtlefsi_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncpus_pinned = pin_to_gpu_numa(local_rank) if world > 1 else 0
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    scaling = "strong" if args.scaling == "auto" else args.scaling

    if world == 1:
        w = workload_single(args.ncube, device, args.deterministic)
    elif scaling == "strong":
        w = workload_strong(args.ncube, device, rank, world)
    else:
        w = workload_weak(args.ncube, device, rank, world)
    m = measure(w, args, device, world, dist, local_rank, args.steps, args.warmup)
    parity = parity_check(w, device, world, dist)
    MX = None if world == 1 else dist.ReduceOp.MAX
    ghost_max = allreduce(getattr(w, "ghost_cell_fraction", 0.0), MX, device, world, dist)
    setup_max = allreduce(w.setup_s, MX, device, world, dist)
    roof = roofline_of(w, m, args.steps) if rank == 0 else None
    mesh_cells, n_fluid = w.mesh.num_cells, int((w.mesh.markers == 1).sum())
    ndofs, nnz, apre_nnz, bound_entries = w.space.num_dofs, w.asm.nnz, w.apre_nnz, w.bound_entries
    ncol = w.asm.num_colours
    pair_mode = bool(w.pair)

    # ---- N > 1, strong: the weak-scaling configuration alongside (fewer steps) ------------------------------
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        del w
        torch.cuda.empty_cache()
        ww = workload_weak(args.ncube, device, rank, world)
        mw = measure(ww, args, device, world, dist, local_rank, max(3, min(args.steps, 5)), 3, full=False)
        pw = parity_check(ww, device, world, dist)
        weak = {"value": mw["value"], "unit": "cells/s", "ms_per_step": mw["ms_step"], "cells_total": mw["total_cells"],
                "cells_per_gpu": mw["total_cells"] / world, "parity": {k: pw[k] for k in ("max_rel_A", "max_rel_b", "halo_exact")},
                "what": "6*%d^3 tets PER GPU (z-slabs of a tube that grows with N), halo exchange inside the step" % args.ncube}
        del ww
        torch.cuda.empty_cache()

    # ---- config 5: size sweep (per-GPU sizes) ---------------------------------------------------------------
    sweep = None
    if args.sweep:
        sweep = []
        try:
            del w
        except NameError:
            pass
        torch.cuda.empty_cache()
        for nc in [int(s) for s in args.sweep.split(",") if s]:
            ws = workload_single(nc, device) if world == 1 else workload_weak(nc, device, rank, world)
            ms_ = measure(ws, args, device, world, dist, local_rank, 5, 3, full=False)
            ps = parity_check(ws, device, world, dist, nodes_per_rank=2)
            free, tot = torch.cuda.mem_get_info(device)
            sweep.append({"ncube": nc, "cells_total": ms_["total_cells"], "cells_per_gpu": ms_["total_cells"] / world,
                          "value": ms_["value"], "ms_per_step": ms_["ms_step"], "hbm_used_gb": (tot - free) / 1e9,
                          "max_rel_A": ps["max_rel_A"], "max_rel_b": ps["max_rel_b"]})
            del ws
            torch.cuda.empty_cache()

    # ---- metric part (ii): one full Newton step on a 3-D tube with the device LU (reported separately) -------
    newton = None
    if rank == 0 and args.newton > 0:
        try:
            newton = newton_step_breakdown(args.newton, device)
        except Exception as e:      # the linear solve is outside the hot path: report, do not fail the bench
            newton = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": m["value"], "unit": "cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_step"],
        "higher_is_better": True, "scaling": "strong" if (world > 1 and scaling == "strong") else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.ncube, world, scaling),
        "roofline": roof,
        "e2e": m["e2e"],
        "gpu_launches": m["launches"],
        "clocks": m["clocks"],
        "parity": parity,
        "problem_size": {"cells_total": m["total_cells"], "rank0_cells_incl_ghost": mesh_cells, "rank0_fluid_cells": n_fluid,
                         "rank0_dofs": ndofs, "rank0_nnz": nnz, "nnz_per_cell": nnz / mesh_cells,
                         "A_pre_nonzero_fraction": apre_nnz / nnz, "entries_rewritten_per_step": bound_entries,
                         "max_ghost_cell_fraction": ghost_max, "colour_classes": ncol, "matrix_pair": pair_mode,
                         "host_cpus_pinned_per_rank": ncpus_pinned},
        "breakdown_ms": {"jacobian_step": m["ms_step"], "jac_fluid_kernel": m["kf_ms"], "jac_solid_kernel": m["ks_ms"],
                         "init_A_pre_rows_if_foreground": m["ms_init"],
                         "jacobian_step_foreground_init": m["ms_step_foreground_init"], "halo_exchange_alone": m.get("ms_halo"),
                         "residual": m["ms_res"], "ident_zeros": m["ms_ident"], "bc_apply": m["ms_bc"], "norm": m["ms_norm"],
                         "linear_solve": (newton or {}).get("linear_solve_ms")},
        "residual_cells_per_s": mesh_cells / (m["ms_res"] * 1e-3),
        "newton_step": newton,
        "setup_s": setup_max,
    }
    if weak is not None:
        line["weak_scaling"] = weak
    if sweep is not None:
        line["sweep"] = sweep
    if not args.no_cpu_baseline:
        ncells, times, cores, sample = cpu_baseline(args.ncube, args.cpu_cells, steps=2)
        line["cpu_baseline"] = {"value": ncells / min(times), "unit": "cells/s", "cores": cores, "kind": "port",
                                "sample": sample}
    print(json.dumps(line))
    ok = parity["max_rel_A"] <= 1e-12 and parity["max_rel_b"] <= 1e-12 and parity["halo_exact"]
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def newton_step_breakdown(ncube, device):
    """BASELINE metric part (ii): wall time of ONE Newton iteration (newtonsolver.py:52-113) on a 3-D tube small enough
    for the one-off host analysis of the device LU — assemble-J / assemble-R / BC + ident / factor / solve / update,
    each device-synchronised. The linear solve is reported separately and not optimised (north_star)."""
    import torch
    from turtlefsi_b200.linear_solver import make_device_solver
    w = workload_single(ncube, device)
    asm = w.asm
    prepare_matrices(w)
    g = torch.zeros(len(w.bc), dtype=torch.float64, device=device)
    solver, solver_name, note = make_device_solver(asm)

    def t(f):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r
    x = w.xd.clone()
    out = {"mesh": "6*%d^3 = %d tets, %d dofs, nnz %d" % (ncube, w.mesh.num_cells, w.space.num_dofs, asm.nnz),
           "solver": solver_name, "solver_note": note}
    # iteration 0: Jacobian re-assembled -> numeric factorisation; iteration 1: Jacobian re-used -> solve only
    for it, recompute in ((0, True), (1, False)):
        rec = {}
        if recompute:
            rec["assemble_J"], _ = t(lambda: asm.assemble_jacobian(x, w.x1d, "compressed", w.A))
            rec["ident_bc_matrix"], _ = t(lambda: (asm.ident_zeros(w.A), asm.apply_bc_matrix(w.A)))
            solver.set_operator(w.A)
        rec["assemble_R"], _ = t(lambda: asm.assemble_residual(x, w.x1d, w.b))
        rec["bc_rhs"], _ = t(lambda: asm.apply_bc_rhs(w.b, x, g))
        s0 = dict(solver.seconds)
        rec["linear_solve_total"], dx = t(lambda: solver.solve(w.b))
        for k in solver.seconds:
            rec["lu_" + k] = (solver.seconds[k] - s0[k]) * 1e3
        rec["update_bc_norms"], _ = t(lambda: (x.add_(dx), asm.apply_bc_vec(x, g), asm.norm(w.b), asm.norm(dx)))
        out["iteration_%d_%s" % (it, "reassembled" if recompute else "reused")] = rec
    r0, r1 = out["iteration_0_reassembled"], out["iteration_1_reused"]
    out["linear_solve_ms"] = {"first_iteration_incl_one_off_host_analysis": r0["linear_solve_total"],
                              "refactor": r0.get("lu_refactor"), "solve_reused": r1["linear_solve_total"]}
    out["newton_step_wall_ms"] = {"reassembled": sum(v for k, v in r0.items() if not k.startswith("lu_")),
                                  "reused": sum(v for k, v in r1.items() if not k.startswith("lu_"))}
    return out


if __name__ == "__main__":
    main()
