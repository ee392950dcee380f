"""turtleFSI/problems/__init__.py counterpart: `default_variables` (same keys and values:
reference :41-117), the default problem-file hooks (reference :465-523) and checkpoint / restart
(reference :183-219, :361-381) with numpy files instead of XDMF/HDF5, and the visualisation output
(reference :222-323) as XDMF time series with raw binary heavy data (turtlefsi_b200.utils.xdmf; h5py is not in
this image), and the kill / pause files (reference :326-358)."""
import json
import os

import numpy as np

_compiler_parameters = dict({"quadrature_degree": 4, "optimize": True})   # problems/__init__.py:16-17

default_variables = dict(
    dt=0.001, theta=0.501, T=1, t=0, counter=0,
    v_deg=2, p_deg=1, d_deg=2,
    dx_f_id=1, dx_s_id=2, ds_s_id=None,
    fluid_properties=[], rho_f=1.0E3, mu_f=1.0,
    solid_properties=[], material_model="StVenantKirchoff",
    rho_s=1.0E3, mu_s=5.0E4, nu_s=0.45, lambda_s=4.5E5, k_s=0.0, c_s=0.0,
    gravity=None, Um=0.8,
    fluid="fluid", solid="solid", robin_bc=False,
    extrapolation="laplace", extrapolation_sub_type="constant", bc_ids=[],
    linear_solver="mumps", solver="newtonsolver",
    atol=1e-7, rtol=1e-7, max_it=50, lmbda=1.0, recompute=5, recompute_tstep=1,
    compiler_parameters=_compiler_parameters,
    loglevel=20, verbose=True, save_step=10, save_deg=1, checkpoint_step=500,
    folder="results", sub_folder=None, restart_folder=None, killtime=None,
)


def set_problem_parameters(**namespace):
    """Set values to the problem variables (reference :465-476)."""
    return {}


def get_mesh_domain_and_boundaries(**namespace):
    raise NotImplementedError("You need to define the mesh, domains and boundaries")


def initiate(**namespace):
    return {}


def create_bcs(**namespace):
    return dict(bcs=[])


def pre_solve(**namespace):
    pass


def post_solve(**namespace):
    pass


def finished(**namespace):
    pass


def create_folders(folder, sub_folder, restart_folder, **namespace):
    """reference :118-180: <cwd>/<folder>/<sub_folder or next free number>/{Checkpoint, Visualization}; a restart
    continues in the folder it restarts from."""
    from pathlib import Path
    if restart_folder is None:
        path = Path.cwd() / folder
        if sub_folder is not None:
            path = path.joinpath(str(sub_folder))
        else:
            numbers = [int(i.name) for i in path.glob("*") if i.name.isdigit()]
            path = path.joinpath(str(max(numbers) + 1 if numbers else 1))
    else:
        path = Path(restart_folder)
    if "Checkpoint" in str(path):
        path = path.parent
    if not path.joinpath("Checkpoint").exists() and restart_folder is not None:
        raise NotADirectoryError(("The restart folder: {} does not have a sub folder 'Checkpoint' where we can"
                                  " restart the simulation from.").format(restart_folder))
    checkpoint_folder = path.joinpath("Checkpoint")
    visualization_folder = path.joinpath("Visualization")
    run_number = keep_previous_visualization(visualization_folder)
    if int(os.environ.get("RANK", "0")) == 0:
        checkpoint_folder.mkdir(parents=True, exist_ok=True)
        visualization_folder.mkdir(parents=True, exist_ok=True)
    return dict(checkpoint_folder=str(checkpoint_folder), visualization_folder=str(visualization_folder),
                results_folder=str(path), run_number=run_number)


def keep_previous_visualization(visualization_folder):
    """reference :147-170: series left by an earlier run (a restart continues in its folder) are renamed
    `<name>_run_<N>.xdmf` (+ their heavy-data file, here `.bin` instead of `.h5`, and the link inside the XML), so the
    new run does not overwrite them; `merge_visualization_files` joins them again at the end. Returns the run number."""
    from pathlib import Path
    folder = Path(visualization_folder)
    if not folder.is_dir() or not list(folder.glob("*")):
        return 0
    done = [int(f.stem.split("_")[-1]) for f in folder.glob("velocity_run_*.bin")]
    run_number = max(done) + 1 if done else 1
    if int(os.environ.get("RANK", "0")) == 0:
        for name in ("displacement", "velocity", "pressure"):
            if not folder.joinpath(name + ".bin").exists():
                continue
            new = name + "_run_" + str(run_number)
            for suffix in (".bin", ".xdmf"):
                folder.joinpath(name + suffix).rename(folder.joinpath(new + suffix))
            xml = folder.joinpath(new + ".xdmf")
            xml.write_text(xml.read_text().replace(">" + name + ".bin<", ">" + new + ".bin<"))
    return run_number


def merge_visualization_files(visualization_folder, **namespace):
    """reference :414-461 (called after a restarted run, monolithic.py:224-227): the steps of the renamed earlier series
    and of this run in ONE `<name>.xdmf`, ordered in time, steps the restart recomputed taken from the later run; the
    heavy-data files stay where they are (every DataItem names its own)."""
    import re
    import xml.etree.ElementTree as ET
    from pathlib import Path
    folder = Path(visualization_folder)
    for name in ("displacement", "velocity", "pressure"):
        files = sorted(folder.glob(name + "*.xdmf"))
        if len(files) < 2 or not folder.joinpath(name + ".xdmf").exists():
            continue
        runs = []
        for f in files:
            root = ET.fromstring(re.sub(r"<!DOCTYPE[^>]*>", "", f.read_text()))
            grids = [g for g in root.iter("Grid") if g.get("GridType") == "Uniform"]
            if grids:
                runs.append((float(grids[0].find("Time").get("Value")), f, root, grids))
        runs.sort(key=lambda r: r[0])
        merged = []
        for k, (_, f, root, grids) in enumerate(runs):
            if k + 1 < len(runs):
                cut = runs[k + 1][0]                                         # the next run starts here again
                grids = [g for g in grids if float(g.find("Time").get("Value")) < cut - 1e-12 * max(1.0, abs(cut))]
            merged += grids
        base = runs[0][2]
        collection = next(g for g in base.iter("Grid") if g.get("GridType") == "Collection")
        for g in list(collection):
            collection.remove(g)
        for k, g in enumerate(merged):
            g.set("Name", "step_%d" % k)
            collection.append(g)
        out = folder.joinpath(name + ".xdmf")
        tmp = folder.joinpath(name + ".xdmf.tmp")
        tmp.write_text('<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n' + ET.tostring(base, encoding="unicode") + "\n")
        for _, f, _, _ in runs:
            if f != out:
                f.unlink()
        os.replace(tmp, out)


def checkpoint(dvp_, default_variables, checkpoint_folder, assembler, **namespace):
    """reference :183-219: store the current parameters (only keys of default_variables, e.g. t and counter) and
    the fields of the LAST time level (d1, v1, p1, (w1) = dvp_["n-1"] after the state rotation). Written to a
    temporary file first, then moved, so that an existing checkpoint is never left corrupted."""
    default_variables.update((k, namespace[k]) for k in (default_variables.keys() & namespace.keys()))
    os.makedirs(checkpoint_folder, exist_ok=True)
    x1 = assembler.to_host(dvp_["n-1"])                      # gathered global vector (collective under torchrun)
    if getattr(assembler, "rank", 0) != 0:
        return
    keep = {k: v for k, v in default_variables.items() if isinstance(v, (int, float, str, bool, list, type(None)))}
    keep["restart_folder"] = None if keep.get("restart_folder") is None else str(keep["restart_folder"])
    with open(os.path.join(checkpoint_folder, "default_variables.json"), "w") as f:
        json.dump(keep, f)
    tmp = os.path.join(checkpoint_folder, "tmp_fields.npy")
    np.save(tmp, np.asarray(x1, dtype=np.float64))
    os.replace(tmp, os.path.join(checkpoint_folder, "checkpoint_fields.npy"))


def start_from_checkpoint(dvp_, restart_folder, assembler, **namespace):
    """reference :361-381: both time levels n and n-1 start from the stored fields."""
    x1 = np.load(os.path.join(restart_folder, "checkpoint_fields.npy"))
    if np.isnan(x1).any():
        raise ValueError("Nan values detected in the checkpoint fields. Stopping simulation.")
    xd = assembler.to_device(x1)
    assembler.copy(dvp_["n-1"], xd)
    assembler.copy(dvp_["n"], xd)


def save_files_visualization(visualization_folder, dvp_, t, save_deg, v_deg, p_deg, mesh, DVP, assembler, **namespace):
    """reference :222-305: displacement.xdmf, velocity.xdmf and pressure.xdmf in the Visualization folder, one stored
    step per call. save_deg == 1 stores the values at the mesh vertices; save_deg == 2 stores the P2 fields d and v as
    P1 fields on the uniformly refined mesh (the reference interpolates with a transfer matrix onto `refine(mesh)`: for
    nodal P2 fields those are the dof values at the vertices and edge mid-points themselves) and the pressure, whose
    degree is below save_deg, on the original mesh (reference :293-298). Returns the file handles on the first call
    (the driver merges them into its namespace, monolithic.py:193-194), {} afterwards."""
    from ..utils.xdmf import XDMFSeries, p2_refined_cells
    if save_deg > v_deg:
        raise ValueError("save_deg must be less than or equal to v_deg")
    x = np.asarray(assembler.to_host(dvp_["n"]), dtype=np.float64)        # gathered global vector (collective)
    if getattr(assembler, "rank", 0) != 0:
        return {}
    g = mesh.gdim
    refined = save_deg > 1 and v_deg >= save_deg
    return_dict = {}
    if "d_file" not in namespace:
        os.makedirs(str(visualization_folder), exist_ok=True)
        if refined:
            coords_dv, cells_dv = DVP.node_coords(mesh), p2_refined_cells(DVP.cell_nodes, g)
        else:
            coords_dv, cells_dv = mesh.coords, mesh.cells
        names = dict(d_file="displacement.xdmf", v_file="velocity.xdmf", p_file="pressure.xdmf")
        return_dict = {k: XDMFSeries(os.path.join(str(visualization_folder), f),
                                     *((mesh.coords, mesh.cells) if k == "p_file" else (coords_dv, cells_dv)))
                       for k, f in names.items()}
        namespace.update(return_dict)
    nodes = np.arange(DVP.num_nodes) if refined else DVP.vertex_node
    d = np.stack([x[DVP.node_dof(0, a)[nodes]] for a in range(g)], axis=1)
    v = np.stack([x[DVP.node_dof(1, a)[nodes]] for a in range(g)], axis=1)
    p = x[DVP.vertex_dof_p()]
    write_solution(d, v, p, namespace["d_file"], namespace["v_file"], namespace["p_file"], t)
    return return_dict


def write_solution(d, v, p, d_file, v_file, p_file, t):
    """reference :308-323: the names ParaView shows."""
    d_file.write("Displacement", d, t)
    v_file.write("Velocity", v, t)
    p_file.write("Pressure", p, t)


def _any_rank(found):
    """MPI.sum(MPI.comm_world, found) > 0 of the reference (:331,:352): every rank takes the same decision."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            import torch
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([int(bool(found))], device=dev)
            dist.all_reduce(t)
            return int(t.item()) > 0
    except ImportError:
        pass
    return bool(found)


def check_if_kill(folder, killtime, elapsed_time):
    """reference :326-345: a file named `killturtle` in the results folder (removed by rank 0), or `killtime` seconds
    of wall time, stop the run cleanly (the driver writes a checkpoint first, monolithic.py:209-212).
    `elapsed_time`: seconds since the start of the run (the reference passes its DOLFIN Timer)."""
    from pathlib import Path
    kill = Path(folder) / "killturtle"
    if _any_rank(kill.exists()):
        if int(os.environ.get("RANK", "0")) == 0 and kill.exists():
            kill.unlink()
            print("killturtle Found! Stopping simulations cleanly...")
        return True
    if _any_rank(killtime is not None and killtime <= float(elapsed_time)):
        if int(os.environ.get("RANK", "0")) == 0:
            print("Given killtime reached! Stopping simulations cleanly...")
        return True
    return False


def check_if_pause(folder):
    """reference :348-358: a file named `pauseturtle` in the results folder pauses the run until it is removed."""
    from pathlib import Path
    pause = Path(folder) / "pauseturtle"
    if _any_rank(pause.exists()):
        if int(os.environ.get("RANK", "0")) == 0:
            print("pauseturtle Found! Simulations paused. Remove " + str(pause) + " to resume simulations...")
        return True
    return False


def print_information(counter, t, T, dt, elapsed_time, verbose, **namespace):
    """reference :400-411: one line per time step (verbose) or a progress bar."""
    if verbose:
        print("Solved for timestep {:d}, t = {:2.08f} in {:3.1f} s".format(counter, t, elapsed_time))
    else:
        j = counter / int(T / dt + 1)
        print("Progress: [{:<20s}] {:2.1f}%, last solve took {:3.1f} s".format("=" * int(20 * j - 1) + ">", 100 * j, elapsed_time),
              end="\r")
