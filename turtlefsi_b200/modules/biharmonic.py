"""turtleFSI/modules/biharmonic.py counterpart: the mesh-lifting term added to F_fluid_linear
(laplace.py:33-46 / biharmonic.py:31-36 cell terms / no_extrapolation.py)."""


def extrapolate_setup(F_fluid_linear, extrapolation_sub_type="constant", **namespace):
    term = dict(F_fluid_linear)
    term.update(extrapolation="biharmonic", extrapolation_sub_type=extrapolation_sub_type)
    return dict(F_fluid_linear=term)


def facet_terms(assembler, mesh, DVP, boundaries, bc_ids, extrapolation_sub_type, **namespace):
    """biharmonic.py:38-41: with "constrained_disp_vel" the lifting gets the exterior-facet terms
    alpha_u (grad(d) n).beta ds(i) - alpha_u (grad(w) n).phi ds(i), i in bc_ids. They are linear in
    (d, w): a fixed sparse operator (COO) that the driver hands to the assembler once
    (A_pre += A_f, b -= A_f x)."""
    if extrapolation_sub_type != "constrained_disp_vel" or not bc_ids:
        return None
    from ..boundary import biharmonic_facet_operator
    return dict(now=biharmonic_facet_operator(mesh, DVP, boundaries, bc_ids, alpha_u=0.01))
