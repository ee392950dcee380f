#!/usr/bin/env python
"""bench.py — monolithic FSI Jacobian assembly throughput (cells/s) on B200.

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...   the CPU restatement on host cores

One "step" = one pass of the hot path over the synthetic 3-D tube mesh: the call that replaces
`assemble(J_nonlinear, tensor=A); A.axpy(1.0, A_pre, True)` of turtleFSI's Newton solver
(turtleFSI/modules/newtonsolver.py:75-79), i.e. `tfsi_assemble_jacobian` (A := A_pre + J_nl(x)).
`value` = cells assembled per second with every input resident in HBM; `e2e` = the same metric
through the host-buffer C-ABI call (H2D of the Newton iterate, Jacobian + ident_zeros + BC rows,
residual + BC, D2H of the residual vector). See DESIGN.md "Measurement".

The reference's own FEniCS/MPI assembly CANNOT be run (no DOLFIN in the image): the CPU arm is
the plain-C restatement under oracle/ ("port"), timed on the box's host cores.
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
QDEG_J, QDEG_R = 4, 6

# Algorithmic FLOPs per cell of the Jacobian kernels (mul/add = 1, fma = 2; C' tensors counted once
# per block, no credit for the half-block recomputation) — derivation in DESIGN.md "FLOP model".
FLOPS_FLUID_PER_Q = 15598
FLOPS_SOLID_PER_Q = 4895
NQ_J = 14


def build_workload(n, device, rank=0, world=1, verify=0):
    from turtlefsi_b200.assembler import FSIAssembler
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    t0 = time.time()
    mesh = tube_mesh(n, seed=1234 + rank)
    space = MixedSpace(mesh)
    asm = FSIAssembler(mesh, space, REGIONS, DT, THETA, extrapolation="laplace",
                       extrapolation_sub_type="constant", quad_degree_jacobian=QDEG_J,
                       quad_degree_residual=QDEG_R, device=device, verify_pattern=verify, hmin=2.0 / n)
    x = synthetic_state(mesh, space, 2024 + rank, h=2.0 / n)
    x1 = synthetic_state(mesh, space, 2025 + rank, h=2.0 / n)
    # synthetic Dirichlet set: d and v on the tube ends and the outer wall
    X = space.node_coords(mesh)
    r = np.hypot(X[:, 0], X[:, 1])
    on_bnd = (X[:, 2] < 1e-9) | (X[:, 2] > 4.0 - 1e-9) | (r > 0.999)
    bc = np.concatenate([space.node_dof(f, a)[on_bnd] for f in (0, 1) for a in range(3)]).astype(np.int32)
    asm.set_bc_dofs(bc)
    return mesh, space, asm, x, x1, bc, time.time() - t0


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
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_baseline(n_cpu=20, steps=1, threads=0):
    """the C restatement (oracle 'port') on a bounded sample: a tube mesh of 6*n_cpu^3 cells with
    the same parameters and state distribution. returns (cells/s, cores, description)."""
    from oracle import cport, fem_tables as ft
    from turtlefsi_b200.mesh import synthetic_state, tube_mesh
    from turtlefsi_b200.space import MixedSpace
    cport.build()
    mesh = tube_mesh(n_cpu)
    space = MixedSpace(mesh)
    pattern = space.pattern()
    x = synthetic_state(mesh, space, 2024, h=2.0 / n_cpu)
    x1 = synthetic_state(mesh, space, 2025, h=2.0 / n_cpu)
    A = np.zeros(len(pattern[1]))
    rule = ft.quadrature(3, QDEG_J)
    # all host cores of the box (torchrun exports OMP_NUM_THREADS=1: do not rely on OpenMP's default)
    cores = threads or len(os.sched_getaffinity(0))
    cport.assemble_jacobian(mesh, space, REGIONS, DT, THETA, x, x1, rule, pattern, A, c1=2000, nthreads=cores)  # warm
    times = []
    for _ in range(steps):
        A[:] = 0.0
        t0 = time.perf_counter()
        cport.assemble_jacobian(mesh, space, REGIONS, DT, THETA, x, x1, rule, pattern, A, nthreads=cores)
        times.append(time.perf_counter() - t0)
    sample = "tube mesh 6*%d^3 = %d tets (%d fluid), Jacobian J_nl into the CSR pattern, plain-C port, OpenMP" % (
        n_cpu, mesh.num_cells, int((mesh.markers == 1).sum()))
    return mesh.num_cells, times, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        pass  # the port has no warm state beyond the first-touch done inside cpu_baseline()
    ncells, times, cores, sample = cpu_baseline(args.cpu_ncube, steps=max(args.steps, 1))
    total = sum(times)
    value = ncells * len(times) / total
    line = {
        "impl": "reference", "metric": "monolithic Jacobian assembly cells/s", "value": value,
        "unit": "cells/s", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 3-D elastic-tube FSI (Test_Cylinder parameters), laplace/constant lifting, "
                               "quadrature degree 4; CPU arm runs a bounded sample: " + sample},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's FEniCS/MPI assembly cannot be run here (no DOLFIN); this is the plain-C restatement",
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncube", type=int, default=int(os.environ.get("TFSI_BENCH_NCUBE", "69")),
                    help="tube mesh of 6*ncube^3 tets per GPU (69 -> 1,971,054)")
    ap.add_argument("--cpu-ncube", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

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
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    from turtlefsi_b200.assembler import probe_fp64_peak
    if world > 1:
        from turtlefsi_b200.partition import build_partitioned_workload
        wl = build_partitioned_workload(args.ncube, device, rank, world, REGIONS, DT, THETA, QDEG_J, QDEG_R)
        mesh, space, asm, x, x1, bc, halo, setup_s = wl
        n_owned_cells = halo.num_owned_cells
    else:
        mesh, space, asm, x, x1, bc, setup_s = build_workload(args.ncube, device)
        halo, n_owned_cells = None, mesh.num_cells
    n_fluid = int((mesh.markers == 1).sum())
    n_solid = mesh.num_cells - n_fluid
    xd, x1d = torch.as_tensor(x).to(device), torch.as_tensor(x1).to(device)
    gd = torch.zeros(len(bc), dtype=torch.float64, device=device)
    A_pre = asm.assemble_A_pre(xd, x1d)
    apre_nnz = asm.compress_A_pre(A_pre)
    A = asm.new_matrix()
    b = asm.new_vector()
    torch.cuda.synchronize()

    def step():
        if halo is not None:
            halo.exchange(xd)           # ghost dofs of the Newton iterate, NCCL send/recv
        asm.assemble_jacobian(xd, x1d, A_pre, A)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = asm.launch_count()
    asm.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = asm.launch_count() - l0       # cell kernels + D2D initialisation copies
    kf_ms, kf_n = asm.kernel_time(0)
    ks_ms, ks_n = asm.kernel_time(1)
    asm.set_profiling(False)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        c = torch.tensor([float(n_owned_cells)], dtype=torch.float64, device=device)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_cells = float(c.item())
    else:
        total_cells = float(n_owned_cells)
    ms_step = ms_total / args.steps
    value = total_cells / (ms_step * 1e-3)

    # ---- residual alongside, and the Newton-assembly breakdown (device timed) ---------------------
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
    ms_res = timed(lambda: asm.assemble_residual(xd, x1d, b))
    ms_ident = timed(lambda: asm.ident_zeros(A))
    ms_bc = timed(lambda: (asm.apply_bc_matrix(A), asm.apply_bc_rhs(b, xd, gd), asm.apply_bc_vec(xd, gd)))
    ms_norm = timed(lambda: asm.sumsq(b))

    # ---- e2e: host buffers through the C-ABI --------------------------------------------------------
    xh = torch.as_tensor(x).pin_memory()
    x1h = torch.as_tensor(x1).pin_memory()
    gh = torch.zeros(len(bc), dtype=torch.float64).pin_memory()
    bh = torch.empty(space.num_dofs, dtype=torch.float64).pin_memory()
    for _ in range(2):
        asm.newton_assemble_host(xh, x1h, gh, A_pre, A, bh, recompute=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        asm.newton_assemble_host(xh, x1h, gh, A_pre, A, bh, recompute=True)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_cells / e2e_s
    h2d = 2 * space.num_dofs * 8 + len(bc) * 8
    d2h = space.num_dofs * 8 + 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (Jacobian over fluid cells) --------------------------------
    peak_tf, _ = probe_fp64_peak()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    flops_fluid = FLOPS_FLUID_PER_Q * NQ_J
    flops_solid = FLOPS_SOLID_PER_Q * NQ_J
    kf_avg_ms = kf_ms / args.steps           # all fluid-cell launches of one step (the list is cut into segments)
    ks_avg_ms = ks_ms / args.steps
    achieved_tf = flops_fluid * n_fluid / (kf_avg_ms * 1e-3) / 1e12
    nnz_per_cell = asm.nnz / mesh.num_cells
    # A written once + non-zeros of A_pre read (8 B each + 1 bit per entry) + dofs, ranks, state, coords per cell
    alg_bytes_step = 8 * asm.nnz + 8 * apre_nnz + asm.nnz // 8 + mesh.num_cells * (64 * 4 + 10 * 64 * 2 + 2 * 8.4 * 8 + 100)
    roofline = {
        "bound": "fp64", "kernel": "jac_fluid_kernel<3,14>", "achieved": achieved_tf, "peak": peak_tf,
        "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "peak_source": "measured: register-resident DFMA probe (tfsi_probe_fp64_peak) in this run; "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "flops_per_cell": flops_fluid, "cells_per_step": n_fluid, "launches_per_step": kf_n // max(args.steps, 1),
        "ms_per_step": kf_avg_ms,
        "traffic": 15209.0 * n_fluid,
        "traffic_note": "bytes per launch = 15.21 KB/cell x fluid cells; per-cell figure from ncu --set full of the same "
                        "kernel at 6*32^3 (profiles/r1_jac_kernels_raw.csv: dram read 969.4 MB + write 707.6 MB for 110263 "
                        "fluid cells); 218 kFLOP/cell -> 14 FLOP/B, above the 5.6 FLOP/B ridge: FP64-bound",
        "hbm": {"achieved": alg_bytes_step / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes_step / (ms_step * 1e-3) / 1e9 / hbm_peak,
                "algorithmic_bytes_per_cell": alg_bytes_step / mesh.num_cells,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
        "solid_kernel": {"kernel": "jac_solid_kernel<3,14>", "ms_per_step": ks_avg_ms, "flops_per_cell": flops_solid,
                         "achieved": flops_solid * n_solid / (ks_avg_ms * 1e-3) / 1e12 if ks_n else None},
    }
    line = {
        "metric": "monolithic Jacobian assembly cells/s", "value": value, "unit": "cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": "synthetic 3-D elastic-tube FSI, 6*%d^3 = %d tets per GPU (%d fluid + %d solid), %d dofs, "
                        "nnz %d (%.0f per cell), P2-P2-P1 monolithic, laplace/constant lifting, Jacobian quadrature "
                        "degree 4 (14 pts), Test_Cylinder parameters" % (
                            args.ncube, mesh.num_cells, n_fluid, n_solid, space.num_dofs, asm.nnz, nnz_per_cell),
            "step": "tfsi_assemble_jacobian: A := A_pre + J_nl(x) (A streamed once from the bitmap-compressed A_pre, "
                    "%.0f %% of the pattern non-zero; then one kernel per cell region: fluid, solid)" % (100.0 * apre_nnz / asm.nnz),
            "l2": "inputs larger than L2 (A and A_pre are %.1f GB each)" % (asm.nnz * 8 / 1e9),
            "parallelism": "cells partitioned in z-slabs, ghost-dof halo exchange per step" if world > 1 else "single GPU",
        },
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3,
                "what": "tfsi_newton_assemble_host, every copy inside the timed region: H2D x,x1,g (pinned; A := A_pre "
                        "runs under it) -> residual+BC -> Jacobian kernels+ident_zeros+BC rows (D2H b + ||b||^2 run "
                        "under them); A stays in HBM for the device LU"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "breakdown_ms": {"jacobian_step": ms_step, "jac_fluid_kernel": kf_avg_ms, "jac_solid_kernel": ks_avg_ms,
                         "residual": ms_res, "ident_zeros": ms_ident, "bc_apply": ms_bc, "norm": ms_norm,
                         "linear_solve": None},
        "residual_cells_per_s": total_cells / world / (ms_res * 1e-3),
        "setup_s": setup_s,
    }
    if not args.no_cpu_baseline:
        ncells, times, cores, sample = cpu_baseline(args.cpu_ncube, steps=1)
        line["cpu_baseline"] = {"value": ncells / times[0], "unit": "cells/s", "cores": cores, "kind": "port",
                                "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
