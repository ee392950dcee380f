"""turtleFSI/modules/biharmonic.py counterpart: alpha_u (w . beta - grad(d) : grad(beta)) * dx_f +
alpha_u grad(w) : grad(phi) * dx_f added to F_fluid_linear (biharmonic.py:31-36) as a named cell integral of the
kernels; the `constrained_disp_vel` exterior-facet terms (:38-41) are a fixed sparse operator (`facet_terms`)."""
from ..ufl_lite import KernelIntegral

SUB_TYPES = ("constrained_disp", "constrained_disp_vel")


def extrapolate_setup(F_fluid_linear, extrapolation_sub_type, dx_f_id_list, **namespace):
    for fluid_region in range(len(dx_f_id_list)):
        F_fluid_linear += KernelIntegral("biharmonic", "linear", dx_f_id_list[fluid_region],
                                         extrapolation_sub_type=extrapolation_sub_type)
    return dict(F_fluid_linear=F_fluid_linear)


def facet_terms(assembler, mesh, DVP, boundaries, bc_ids, extrapolation_sub_type, **namespace):
    """biharmonic.py:38-41: with "constrained_disp_vel" the lifting gets the exterior-facet terms
    alpha_u (grad(d) n).beta ds(i) - alpha_u (grad(w) n).phi ds(i), i in bc_ids. They are linear in
    (d, w): a fixed sparse operator (COO) that the driver hands to the assembler once
    (A_pre += A_f, b -= A_f x)."""
    if extrapolation_sub_type != "constrained_disp_vel" or not bc_ids:
        return None
    from ..boundary import biharmonic_facet_operator
    return dict(now=biharmonic_facet_operator(mesh, DVP, boundaries, bc_ids, alpha_u=0.01))
