"""turtleFSI/monolithic.py counterpart: the driver of a problem file, with the reference's order
of operations (monolithic.py:20-230): parameters -> mesh -> mixed space -> domain properties ->
physics modules (selected BY NAME with the same --fluid/--solid/--extrapolation/--solver values)
-> initiate -> create_bcs -> solver_setup -> time loop (pre_solve, newtonsolver, state rotation,
post_solve) -> finished. Every hook is called as f(**namespace) and its returned dict is merged
back, as the reference does with vars().update(...)."""
import copy
import time
import importlib

from .problems import default_variables as _defaults
from .space import MixedSpace


def assign_domain_properties(dx_f_id, rho_f, mu_f, fluid_properties, dx_s_id, material_model, rho_s, mu_s,
                             lambda_s, solid_properties, ds_s_id=None, robin_bc=False, **namespace):
    """turtleFSI/modules/domain.py:16-121: scalar-or-list inputs -> per-region property lists."""
    as_list = lambda v: v if isinstance(v, list) else [v]
    dx_f_id_list, dx_s_id_list = as_list(dx_f_id), as_list(dx_s_id)
    if len(fluid_properties) == 0:
        fluid_properties = [{"dx_f_id": i, "rho_f": as_list(rho_f)[k] if isinstance(rho_f, list) else rho_f,
                             "mu_f": as_list(mu_f)[k] if isinstance(mu_f, list) else mu_f}
                            for k, i in enumerate(dx_f_id_list)]
    elif isinstance(fluid_properties, dict):
        fluid_properties = [fluid_properties]
    if len(solid_properties) == 0:
        pick = lambda v, k: v[k] if isinstance(v, list) else v
        solid_properties = [{"dx_s_id": i, "material_model": pick(material_model, k), "rho_s": pick(rho_s, k),
                             "mu_s": pick(mu_s, k), "lambda_s": pick(lambda_s, k)} for k, i in enumerate(dx_s_id_list)]
    elif isinstance(solid_properties, dict):
        solid_properties = [solid_properties]
    from .modules.solid import material_constants
    regions = [dict(marker=fp["dx_f_id"], kind="fluid", rho=fp["rho_f"], mu=fp["mu_f"]) for fp in fluid_properties]
    regions += [dict(marker=sp["dx_s_id"], kind="solid", rho=sp["rho_s"], mu=sp["mu_s"], lambda_=sp["lambda_s"],
                     **material_constants(sp)) for sp in solid_properties]
    ds_s_ext_id_list = (ds_s_id if isinstance(ds_s_id, list) else [ds_s_id]) if robin_bc else None   # domain.py:106-119
    return dict(dx_f_id_list=dx_f_id_list, dx_s_id_list=dx_s_id_list, fluid_properties=fluid_properties,
                solid_properties=solid_properties, regions=regions, ds_s_ext_id_list=ds_s_ext_id_list)


def _set_facet_terms(assembler, ns):
    """the LINEAR facet integrals of the selected modules (biharmonic.py:38-41 ds(i) terms, solid.py:76-86
    Robin ds_s terms) as fixed COO operators, handed to the assembler once."""
    import numpy as np
    terms = {"now": [], "prev": []}
    for key in ("extrapolation", "solid"):
        m = _physics_module(ns[key])
        t = m.facet_terms(**ns) if hasattr(m, "facet_terms") else None
        for lvl in terms:
            if t and t.get(lvl) is not None and len(t[lvl][2]):
                terms[lvl].append(t[lvl])
    for lvl, setter in (("now", "set_facet_operator"), ("prev", "set_facet_operator_prev")):
        if terms[lvl]:
            getattr(assembler, setter)(*[np.concatenate([t[i] for t in terms[lvl]]) for i in range(3)])


def _physics_module(name):
    """`exec("from turtleFSI.modules.{} import fluid_setup".format(fluid))` of monolithic.py:139-144: the module is
    looked up BY NAME under turtlefsi_b200.modules; a name that is not there is imported as the user's own module
    (it must return turtlefsi_b200.ufl_lite Forms, see modules/fluid.py)."""
    try:
        return importlib.import_module("turtlefsi_b200.modules." + name)
    except ModuleNotFoundError as e:
        if e.name != "turtlefsi_b200.modules." + name:
            raise
        return importlib.import_module(name)


def load_problem(problem):
    """problem: a module object (or any object with the seven hooks), the name of a module under
    turtlefsi_b200.problems, or the path of a problem file. A file that is written against DOLFIN (`from dolfin import
    *`, e.g. the reference's own turtleFSI/problems/TF_fsi.py, unmodified) is run through turtlefsi_b200.compat."""
    import os
    if not isinstance(problem, str):
        return problem
    path = problem if problem.endswith(".py") else problem + ".py"
    if os.path.isfile(path):                                                # monolithic.py:24-25: a local file first
        from .compat import load_problem_file
        return load_problem_file(path)
    return importlib.import_module("turtlefsi_b200.problems." + problem)


def run(problem, make_assembler, make_solver, overrides=None, max_steps=None, create_results_folder=False, progress=False):
    """problem: see load_problem. make_assembler(mesh, space, regions, ns) -> assembler;
    make_solver(assembler) -> up_sol. Returns the final namespace (what the reference keeps in its module globals)."""
    mod = load_problem(problem)
    ns = {}
    default_variables = copy.deepcopy(_defaults)
    import argparse
    ns["args"] = argparse.Namespace(**(overrides or {}))                                             # monolithic.py:21
    default_variables.update(mod.set_problem_parameters(default_variables=default_variables, args=ns["args"]) or {})   # :35
    default_variables.update(overrides or {})                                                        # CLI: :38-43
    if default_variables.get("restart_folder") is not None:                                          # :44-50
        import json
        import os
        rf = str(default_variables["restart_folder"])
        rf = rf if "Checkpoint" in rf else os.path.join(rf, "Checkpoint")
        with open(os.path.join(rf, "default_variables.json")) as f:
            default_variables.update(json.load(f))                    # previous settings, incl. t and counter
        default_variables.update({k: v for k, v in (overrides or {}).items() if k in ("T", "verbose", "checkpoint_step")})
        default_variables["restart_folder"] = rf
    ns.update(default_variables)
    ns["default_variables"] = default_variables
    ns.setdefault("results_folder", None)
    if create_results_folder:                                                                         # :59
        from .problems import create_folders
        ns.update(create_folders(**ns))
    mesh, domains, boundaries = mod.get_mesh_domain_and_boundaries(**ns)                              # :63
    ns.update(mesh=mesh, domains=domains, boundaries=boundaries)
    DVP = MixedSpace(mesh, biharmonic=(ns["extrapolation"] == "biharmonic"))                          # :77-92
    ns["DVP"] = DVP
    ns.update(assign_domain_properties(**ns))                                                         # :124-125
    for key, fn in (("fluid", "fluid_setup"), ("solid", "solid_setup"), ("extrapolation", "extrapolate_setup")):
        ns.update(getattr(_physics_module(ns[key]), fn)(**ns))                                        # :139-144
    # what the modules RETURNED decides what is assembled: region table and switches from the kernel terms of the forms
    from .ufl_lite import regions_of
    regions, switches = regions_of(ns["F_fluid_linear"] + ns["F_fluid_nonlinear"] + ns["F_solid_linear"] + ns["F_solid_nonlinear"])
    if (switches["extrapolation"] == "biharmonic") != DVP.biharmonic:
        raise RuntimeError("the mixed space has %s w field but the forms %s the biharmonic lifting"
                           % ("a" if DVP.biharmonic else "no", "lack" if DVP.biharmonic else "carry"))
    ns["regions"] = regions
    assembler = make_assembler(mesh, DVP, regions, dict(ns, **switches))
    ns["assembler"] = assembler
    _set_facet_terms(assembler, ns)
    dvp_ = {k: assembler.new_vector() for k in ("n-2", "n-1", "n")}                                   # :95-111
    ns["dvp_"] = dvp_
    ns["up_sol"] = make_solver(assembler)                                                             # :136
    ns.update(mod.initiate(**ns) or {})                                                               # :147
    ns.update(mod.create_bcs(**ns) or {})                                                             # :150
    solver = _physics_module(ns["solver"])                                                            # :153
    ns.update(solver.solver_setup(**ns))
    ns["dvp_res"] = assembler.new_vector()                                                            # :157
    # restart data are loaded AFTER solver_setup (:160-162): A_pre stays frozen at the state before them
    from .problems import (check_if_kill, check_if_pause, checkpoint, merge_visualization_files, print_information,
                           save_files_visualization, start_from_checkpoint)
    if ns.get("restart_folder") is not None:
        start_from_checkpoint(**ns)
    if ns.get("results_folder") is not None:
        import os
        ns.setdefault("checkpoint_folder", os.path.join(str(ns["results_folder"]), "Checkpoint"))
    ns.setdefault("residual_history", [])
    t, T, dt = ns["t"], ns["T"], ns["dt"]
    counter = ns["counter"]
    ns["first_step_num"] = counter                                                                    # :168
    stop = False
    t_start = t_step = time.time()
    while t <= T + dt / 10 and not stop:                                                              # :169
        t += dt
        ns.update(t=t, counter=counter)
        tmp = mod.pre_solve(**ns)
        if tmp is not None:
            ns.update(tmp)
        ns.update(solver.newtonsolver(**ns))                                                          # :177
        assembler.copy(dvp_["n-2"], dvp_["n-1"])                                                      # :179-182
        assembler.copy(dvp_["n-1"], dvp_["n"])
        tmp = mod.post_solve(**ns)
        if tmp is not None:
            ns.update(tmp)
        if ns.get("checkpoint_folder") is not None and counter % ns["checkpoint_step"] == 0:               # :193-195
            ns.update(t=t, counter=counter)
            checkpoint(**ns)
        if ns.get("visualization_folder") is not None and counter % ns["save_step"] == 0:                 # :192-194
            ns.update(t=t, counter=counter)
            ns.update(save_files_visualization(**ns))
        counter += 1
        if progress and getattr(assembler, "rank", 0) == 0:                                            # :199-201
            now = time.time()
            print_information(counter=counter, t=t, T=T, dt=dt, elapsed_time=now - t_step, verbose=ns["verbose"])
            t_step = now
        if ns.get("results_folder") is not None:                                                      # :203-212
            while check_if_pause(ns["results_folder"]):
                time.sleep(ns.get("pause_poll_seconds", 5))
            if check_if_kill(ns["results_folder"], ns.get("killtime"), time.time() - t_start):
                ns.update(t=t, counter=counter)
                checkpoint(**ns)
                stop = True
        if max_steps is not None and counter >= max_steps:
            stop = True
    ns.update(t=t, counter=counter)
    if ns.get("restart_folder") is not None and ns.get("visualization_folder") is not None \
            and getattr(assembler, "rank", 0) == 0:                                                   # :224-227
        merge_visualization_files(**ns)
    mod.finished(**ns)
    return ns
