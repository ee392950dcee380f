"""turtleFSI/modules/laplace.py counterpart: alfa * inner(grad(d), grad(phi)) * dx_f added to F_fluid_linear for every
fluid region (laplace.py:33-46), alfa chosen by `extrapolation_sub_type` — as a named cell integral of the kernels."""
from ..ufl_lite import KernelIntegral

SUB_TYPES = ("constant", "small_constant", "volume", "volume_change")


def extrapolate_setup(F_fluid_linear, extrapolation_sub_type, dx_f_id_list, **namespace):
    if extrapolation_sub_type not in SUB_TYPES:
        raise RuntimeError("Could not find extrapolation method {}".format(extrapolation_sub_type))      # laplace.py:41-42
    for fluid_region in range(len(dx_f_id_list)):
        F_fluid_linear += KernelIntegral("laplace", "linear", dx_f_id_list[fluid_region],
                                         extrapolation_sub_type=extrapolation_sub_type)
    return dict(F_fluid_linear=F_fluid_linear)
