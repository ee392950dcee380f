"""turtleFSI/modules/fluid.py counterpart: the ALE Navier-Stokes residual of every fluid region, split into a linear
and a non-linear part exactly as the reference splits it (fluid.py:32-56: N1..N6 non-linear, L1, L2, L4 linear) — but
as NAMED cell integrals the sm_100a kernels implement (`jac_fluid_kernel`, `residual_kernel`, `lin_jacobian_kernel`)
instead of UFL. What is returned here is what gets assembled: the driver builds the assembler's region table from the
terms present in the forms (ufl_lite.regions_of)."""
from ..ufl_lite import Form, KernelIntegral


def fluid_setup(dx_f_id_list, fluid_properties, **namespace):
    F_fluid_linear = Form()
    F_fluid_nonlinear = Form()
    for fluid_region in range(len(dx_f_id_list)):
        const = dict(rho=fluid_properties[fluid_region]["rho_f"], mu=fluid_properties[fluid_region]["mu_f"])
        F_fluid_nonlinear += KernelIntegral("fluid", "nonlinear", dx_f_id_list[fluid_region], **const)
        F_fluid_linear += KernelIntegral("fluid", "linear", dx_f_id_list[fluid_region], **const)
    return dict(F_fluid_linear=F_fluid_linear, F_fluid_nonlinear=F_fluid_nonlinear)
