"""turtleFSI/modules/newtonsolver.py counterpart — same functions, same control flow, same
returned keys; the four DOLFIN calls that carry the cost are replaced by C-ABI calls into the
sm_100a kernels (through an assembler object: turtlefsi_b200.assembler.FSIAssembler on the GPU).

  reference line (modules/newtonsolver.py)                 here
  :26-27  A_pre = assemble(J_linear, keep_diagonal=True)   asm.assemble_A_pre(x, x1)
  :28     A = Matrix(A_pre)                                asm.new_matrix() (same fixed pattern)
  :75-79  assemble(J_nonlinear, tensor=A); A.axpy(A_pre)   asm.assemble_jacobian(x, x1, A_pre, A)
  :81     A.ident_zeros()                                  asm.ident_zeros(A)
  :82     [bc.apply(A) for bc in bcs]                      asm.apply_bc_matrix(A)
  :88     assemble(-F, tensor=b)                           asm.assemble_residual(x, x1, b)
  :91     [bc.apply(b, x) ...]                             asm.apply_bc_rhs(b, x, g)
  :93     up_sol.solve(dvp_res.vector(), b)                up_sol.solve(b)   (not optimised here)
  :95,98  x.axpy(lmbda, dx); [bc.apply(x) ...]             asm.axpy / asm.apply_bc_vec(x, g)
  :105-106 b.norm('l2'), norm(dvp_res, 'l2')               asm.norm
"""
from math import isnan

from ..boundary import BCPlan
from ..ufl_lite import regions_of


def solver_setup(F_fluid_linear, F_fluid_nonlinear, F_solid_linear, F_solid_nonlinear, assembler, dvp_, up_sol, bcs,
                 regions=None, **namespace):
    """Pre-assemble the linear part of the Jacobian (frozen at the state present now, as in the reference: SURVEY 8a
    "frozen-A_pre"). The forms are summed as the reference sums them (newtonsolver.py:18-24); their cell integrals
    are the kernels' (the assembler was configured from them), facet terms a problem file has added since
    (problem_aneu.py:214) are collected as loads."""
    F_lin = F_fluid_linear + F_solid_linear
    F_nonlin = F_solid_nonlinear + F_fluid_nonlinear
    F = F_lin + F_nonlin
    if regions is not None:
        now, _ = regions_of(F)
        if sorted(map(repr, map(sorted, (r.items() for r in now)))) != sorted(map(repr, map(sorted, (r.items() for r in regions)))):
            raise RuntimeError("the cell integrals of the forms changed after the physics modules were set up; "
                               "only facet terms can be added in initiate / create_bcs")
    loads = F.expression_terms()
    for it in loads:
        if it.test_function() is None:
            raise ValueError("a term added to the residual form has no test function")
        if it.depends_on_state():
            raise NotImplementedError("a facet / cell term that depends on the unknown was added to the residual form: "
                                      "only loads (coefficients, expressions, normals) are evaluated on this path")
    bc_plan = BCPlan(bcs)                     # dof index maps built once; only values() runs per step
    assembler.set_bc_dofs(bc_plan.dofs)
    x = lambda f: getattr(f, "vec", f)        # dvp_ holds assembler vectors (or dolfin-style Functions around them)
    A_pre = assembler.assemble_A_pre(x(dvp_["n"]), x(dvp_["n-1"]))
    if hasattr(assembler, "compress_A_pre"):
        assembler.compress_A_pre(A_pre)           # A_pre is final here (problem files edit the forms before)
    A = assembler.new_matrix()
    assembler.copy(A, A_pre)                  # A = Matrix(A_pre)
    A_pair = None
    if hasattr(assembler, "bind_matrix_pair") and assembler.pair_fits():
        # two persistent matrices used alternately by the re-assemblies (the previous Jacobian is dead once the next one
        # is assembled): the idle one is re-initialised from inside the fluid kernel of the running assembly
        A_pair = (A, assembler.new_matrix())
        assembler.bind_matrix_pair(A_pair[0], A_pair[1], "compressed")
    elif hasattr(assembler, "bind_matrix"):
        assembler.bind_matrix(A)              # persistent solver matrix: re-assemblies rewrite only the rows J_nl touches
    b = None
    return dict(F=F, J_nonlinear="derivative(F_nonlin, dvp_n): hand-derived in the kernels", A_pre=A_pre, A=A, b=b,
                up_sol=up_sol, bc_plan=bc_plan, facet_loads=loads, A_pair=A_pair)


def _update_facet_loads(assembler, facet_loads, mesh, DVP, boundaries, compiler_parameters):
    """loads added by the problem file, p(t) * inner(n('+'), psi('+')) * dS(id) and the like: re-evaluated with the
    CURRENT values of their coefficients (pre_solve has just updated them) and handed to the device as a sparse
    vector (tfsi_set_facet_load: b -= l)."""
    import numpy as np
    from ..ufl_lite import assemble_load
    # assemble(-F) carries no form_compiler_parameters in the reference (newtonsolver.py:88): estimated degree
    l = assemble_load(facet_loads, mesh, DVP, boundaries)
    dofs = np.nonzero(l)[0].astype(np.int32)
    assembler.set_facet_load(dofs, l[dofs], 1.0)


def newtonsolver(F, J_nonlinear, A_pre, A, b, bcs, lmbda, recompute, recompute_tstep, compiler_parameters,
                 dvp_, up_sol, dvp_res, rtol, atol, max_it, counter, first_step_num, verbose, assembler,
                 residual_history=None, bc_plan=None, facet_loads=None, A_pair=None, **namespace):
    """Newton scheme with the reference's Jacobian-reuse predicates (newtonsolver.py:44-113)."""
    # Initial values
    iter = 0
    residual = 10**8
    rel_res = 10**8

    # Capture if residual increases from last iteration
    last_rel_res = residual
    last_residual = rel_res
    x, x1 = dvp_["n"], dvp_["n-1"]
    if facet_loads:
        _update_facet_loads(assembler, facet_loads, namespace["mesh"], namespace["DVP"], namespace["boundaries"],
                            compiler_parameters)
    # The dof set uploaded by solver_setup is normally fixed and only the values are time dependent; a pre_solve
    # hook may nevertheless hand back another list (the reference re-reads `bcs` at every call): then the device
    # set is replaced, so that g always lines up with the dofs the kernels index.
    if bc_plan is None or not bc_plan.matches(bcs):
        bc_plan = BCPlan(bcs)
    if getattr(assembler, "bc_dofs", None) is None or len(assembler.bc_dofs) != len(bc_plan.dofs) \
            or not (assembler.bc_dofs == bc_plan.dofs).all():
        assembler.set_bc_dofs(bc_plan.dofs)
    g = assembler.to_device(bc_plan.values(bcs))
    while rel_res > rtol and residual > atol and iter < max_it:
        # Check if recompute Jacobian from 'recompute_tstep' (time step)
        recompute_for_timestep = iter == 0 and (counter % recompute_tstep == 0)

        # Check if recompute Jacobian from 'recompute' (iteration)
        recompute_frequency = iter > 0 and iter % recompute == 0

        # Recompute Jacobian due to increased residual
        recompute_residual = iter > 0 and last_residual < residual

        # Recompute Jacobian on first step of simulation (important if restart is used)
        recompute_initialize = iter == 0 and counter == first_step_num

        if recompute_for_timestep or recompute_frequency or recompute_residual or recompute_initialize:
            if verbose:
                print("Compute Jacobian matrix")
            if A_pair is not None:                            # the other matrix of the pair: the one in use is dead from here on
                A = A_pair[1] if A is A_pair[0] else A_pair[0]
            assembler.assemble_jacobian(x, x1, A_pre, A)      # J_nonlinear + A_pre
            assembler.ident_zeros(A)
            assembler.apply_bc_matrix(A)
            up_sol.set_operator(A)                            # new values -> re-factorise

        # Assemble right hand side vector
        if b is None:
            b = assembler.assemble_residual(x, x1)
        else:
            assembler.assemble_residual(x, x1, b)

        # Apply boundary conditions before solve
        assembler.apply_bc_rhs(b, x, g)
        # Solve the linear system A * x = b
        dx = up_sol.solve(b)
        assembler.copy(dvp_res, dx)
        # Update solution using the Newton increment, then re-apply the boundary conditions
        assembler.axpy(x, lmbda, dvp_res)
        assembler.apply_bc_vec(x, g)

        # Reset residuals
        last_residual = residual
        last_rel_res = rel_res

        # Check residual
        residual = assembler.norm(b)
        rel_res = assembler.norm(dvp_res)
        if rel_res > 1E20 or residual > 1E20 or isnan(rel_res) or isnan(residual):
            raise RuntimeError("Error: The simulation has diverged during the Newton solve with residual = %.3e and relative residual = %.3e" % (residual, rel_res))

        if verbose:
            print("Newton iteration %d: r (atol) = %.3e (tol = %.3e), r (rel) = %.3e (tol = %.3e) "
                  % (iter, residual, atol, rel_res, rtol))
        if residual_history is not None:
            residual_history.append((counter, iter, residual, rel_res))
        iter += 1

    return dict(up_sol=up_sol, A=A, b=b, bc_plan=bc_plan)
