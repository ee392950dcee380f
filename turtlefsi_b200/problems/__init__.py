"""turtleFSI/problems/__init__.py counterpart: `default_variables` (same keys and values:
reference :41-117) and the default problem-file hooks (reference :465-523). I/O (XDMF,
checkpoints, kill/pause files) is out of scope of this repository."""

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
