"""turtleFSI/modules/elastic.py counterpart: the linear-elasticity lifting J S_linear(d) F^-T : grad(phi) * dx_f
(elastic.py:18-27; E = 1 / CellVolume, nu = 0.25) added to F_fluid_linear — a named cell integral of the kernels
(non-linear in d, frozen into A_pre at the set-up state as the reference does)."""
from ..ufl_lite import KernelIntegral


def extrapolate_setup(F_fluid_linear, dx_f_id_list, **namespace):
    for fluid_region in range(len(dx_f_id_list)):
        F_fluid_linear += KernelIntegral("elastic", "linear", dx_f_id_list[fluid_region])
    return dict(F_fluid_linear=F_fluid_linear)
