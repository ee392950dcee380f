"""Robin-BC validation — turtleFSI/problems/RobinBC_validation.py re-hosted without DOLFIN: a solid
cylinder under constant gravity (y), held only by the Robin condition (spring k_s, damper c_s) on its
outer wall, compared with a mass-spring-damper ODE (:96-108, :133-158). The reference's mesh is an
external download (:13-14); here the cylinder is `tube_mesh` with every cell solid (the reference sets
all domains to the solid id as well, :85-86)."""
import numpy as np

from turtlefsi_b200.boundary import Facets, point_eval
from turtlefsi_b200.mesh import tube_mesh
from turtlefsi_b200.problems import *  # noqa: F401,F403


def set_problem_parameters(default_variables, **namespace):          # RobinBC_validation.py:25-72
    E_s_val, nu_s_val = 1E6, 0.45
    mu_s_val = E_s_val / (2 * (1 + nu_s_val))
    lambda_s_val = nu_s_val * 2. * mu_s_val / (1. - 2. * nu_s_val)
    default_variables.update(dict(
        T=0.1, dt=0.0005, theta=0.501, atol=1e-10, rtol=1e-10,
        outer_wall_id=1033, ds_s_id=[1033],
        rho_f=1.025E3, mu_f=3.5E-3, rho_s=1.0E3, mu_s=mu_s_val, nu_s=nu_s_val, lambda_s=lambda_s_val,
        robin_bc=True, k_s=[1.0E5], c_s=[1.0E2], dx_f_id=1, dx_s_id=1002,
        extrapolation="laplace", extrapolation_sub_type="constant",
        recompute=30, recompute_tstep=100, folder="robinbc_validation", gravity=2.0, fluid="no_fluid",
        ncube=4, length=4.0, residual_quadrature_degree=6))
    return default_variables


def get_mesh_domain_and_boundaries(ncube, length, outer_wall_id, dx_s_id, **namespace):   # :75-88
    mesh = tube_mesh(ncube, length=length)
    mesh.markers[:] = dx_s_id                                                    # domains.set_all(1002)
    boundaries = Facets(mesh)
    boundaries.mark(lambda x, on: on, outer_wall_id)                             # lateral surface ...
    boundaries.mark(lambda x, on: on & (np.abs(x[:, 2]) < 1e-12), 1011)          # ... minus the two end planes
    boundaries.mark(lambda x, on: on & (np.abs(x[:, 2] - length) < 1e-12), 1011)
    return mesh, mesh.markers, boundaries


def create_bcs(**namespace):                                         # :90-95: Robin only, no Dirichlet data
    return dict(bcs=[])


def _mass_spring_damper_system_ode(x, t, params_dict):               # :96-108
    F, A, c, k, m = (params_dict[n] for n in "FAckm")
    return [x[1], (F - c * x[1] * A - k * x[0] * A) / m]


def initiate(mesh, **namespace):                                     # :110-124
    middle_point = 0.5 * (mesh.coords.max(axis=0) + mesh.coords.min(axis=0))
    return dict(d_list=[], middle_point=middle_point)


def post_solve(dvp_, d_list, middle_point, mesh, DVP, assembler, **namespace):    # :127-132
    d_list.append(point_eval(mesh, DVP, assembler.to_host(dvp_["n"]), middle_point, field=0)[1])
    return dict(d_list=d_list)


def finished(T, dt, mesh, rho_s, k_s, c_s, boundaries, gravity, d_list, outer_wall_id, **namespace):   # :135-158
    from scipy.integrate import odeint
    X = mesh.coords[mesh.cells]
    volume = float(np.abs(np.linalg.det(X[:, 1:] - X[:, :1])).sum() / 6.0)       # assemble(1*dx)
    fs = np.nonzero((boundaries.markers == outer_wall_id) & boundaries.on_boundary)[0]
    P = mesh.coords[boundaries.vertices[fs]]
    area = float(0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1).sum())   # assemble(1*ds_robin)
    params = dict(m=volume * rho_s, k=k_s[0], c=c_s[0], A=area, F=-gravity * volume * rho_s)
    t_analytical = dt * np.arange(len(d_list) + 1)
    sol = odeint(_mass_spring_damper_system_ode, [0, 0], t_analytical, args=(params,))
    namespace["default_variables"].update(analytical_displacement=sol[1:, 0], ode_params=params)   # the reference plots them
