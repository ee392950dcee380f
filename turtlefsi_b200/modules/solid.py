"""turtleFSI/modules/solid.py counterpart: inertia + kinematic constraint (linear, solid.py:41-43), the stress of the
region's material model (non-linear, :62-63; common.py:104-125), viscoelasticity (:46-59) and gravity (:65-68) of every solid region as named cell integrals of the sm_100a kernels
(`jac_solid_kernel`, `residual_kernel`, `lin_jacobian_kernel`); the Robin condition (:76-86) is a pair of fixed sparse
facet operators (`facet_terms`)."""
from ..ufl_lite import Form, KernelIntegral


MATERIAL_KEYS = {"StVenantKirchoff": (), "StVenantKirchoffEnergy": (), "NeoHookean": (),
                 "MooneyRivlin": ("C01", "C10", "C11"), "Gent": ("Jm",)}


def material_constants(sp):
    """the constants of one solid region beyond rho_s / mu_s / lambda_s, under the reference's own keys: material model
    (common.py:104-125; the extra constants are read there with sp["C01"] ... so a missing one is a KeyError, as in the
    reference) and viscoelasticity (solid.py:46-59, Svisc_D of common.py:129-139; an entry other than Form1 / Form2 means
    none, solid.py:56-61)."""
    model = sp["material_model"]
    if model not in MATERIAL_KEYS:
        raise RuntimeError('Invalid entry for material_model, choose from ["StVenantKirchoff",'
                           '"StVenantKirchoffEnergy","NeoHookean","MooneyRivlin","Gent"]')
    const = dict(material_model=model)
    for key in MATERIAL_KEYS[model]:
        const[key] = sp[key]
    if sp.get("viscoelasticity") in ("Form1", "Form2"):
        const.update(viscoelasticity=sp["viscoelasticity"], mu_visc_s=sp["mu_visc_s"], lambda_visc_s=sp["lambda_visc_s"])
    return const


def solid_setup(dx_s_id_list, solid_properties, **namespace):
    F_solid_linear = Form()
    F_solid_nonlinear = Form()
    for solid_region in range(len(dx_s_id_list)):
        sp = solid_properties[solid_region]
        const = dict(rho=sp["rho_s"], mu=sp["mu_s"], lambda_=sp["lambda_s"])
        const.update(material_constants(sp))
        F_solid_linear += KernelIntegral("solid", "linear", dx_s_id_list[solid_region], **const)
        F_solid_nonlinear += KernelIntegral("solid", "nonlinear", dx_s_id_list[solid_region], **const)
    return dict(F_solid_linear=F_solid_linear, F_solid_nonlinear=F_solid_nonlinear)


def facet_terms(mesh, DVP, boundaries, theta, robin_bc=False, ds_s_ext_id_list=None, k_s=0.0, c_s=0.0, **namespace):
    """solid.py:76-86: Robin condition theta0 (k_s d + c_s v).psi ds_s + theta1 (k_s d^{n-1} + c_s v^{n-1}).psi ds_s.
    Linear in the state -> two fixed COO operators: `now` (A_pre += , b -= now x) and `prev` (b -= prev x1).
    Same argument checks as the reference (:79-81)."""
    if not robin_bc:
        return None
    from collections.abc import Iterable
    assert isinstance(k_s, Iterable), "k_s should be an iterable (e.g., list, tuple, etc.)."
    assert isinstance(c_s, Iterable), "c_s should be an iterable (e.g., list, tuple, etc.)."
    assert len(k_s) == len(c_s) == len(ds_s_ext_id_list), "k_s, c_s, and ds_s_ext_id_list should have the same length."
    from ..boundary import robin_facet_operators
    now, prev = robin_facet_operators(mesh, DVP, boundaries, ds_s_ext_id_list, k_s, c_s, theta)
    return dict(now=now, prev=prev)
