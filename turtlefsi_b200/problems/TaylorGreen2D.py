"""Taylor-Green vortex on a fixed domain — turtleFSI/problems/TaylorGreen2D.py re-hosted without
DOLFIN (structured-mesh branch, :121-123). A DOLFIN-free KNOWN ANSWER for the fluid forms: the
analytic velocity/pressure (:134-155) are imposed on the boundary and as initial state, and the
L2 distance of the discrete solution to their interpolants is accumulated per step (:211-231)."""
import numpy as np

from turtlefsi_b200.boundary import DirichletBC, Facets, PointwiseBC, interpolate, l2_norm
from turtlefsi_b200.mesh import rectangle_mesh
from turtlefsi_b200.problems import *  # noqa: F401,F403

DOLFIN_EPS = 3.0e-16


def set_problem_parameters(default_variables, **namespace):          # TaylorGreen2D.py:19-46
    default_variables.update(dict(
        mu_f=0.01, T=1, dt=0.01, theta=0.5, rho_f=1,
        folder="tg2d_results", solid="no_solid", extrapolation="no_extrapolation",
        save_step=500, checkpoint_step=500, L=2., v_deg=2, p_deg=1, atol=1e-10, rtol=1e-10,
        total_error_v=0, total_error_p=0, mesh_size=0.25, mesh_type="structured", external_mesh=False, N=40,
        recompute=100, recompute_tstep=100))
    return default_variables


def get_mesh_domain_and_boundaries(N, mesh_type, external_mesh, **namespace):   # TaylorGreen2D.py:91-133
    if external_mesh or mesh_type != "structured":
        raise RuntimeError("only the structured RectangleMesh branch is re-hosted")
    mesh = rectangle_mesh((-1.0, -1.0), (1.0, 1.0), N, N, "right")
    boundaries = Facets(mesh)
    boundaries.mark(lambda x, on: on, 1)                                          # Wall: the whole boundary
    mesh.markers[:] = 1
    return mesh, mesh.markers, boundaries


class analytical_velocity:                                           # TaylorGreen2D.py:135-146
    def __init__(self):
        self.t, self.nu = 0, 0.01

    def __call__(self, x):
        e = np.exp(-2. * np.pi * np.pi * self.nu * self.t)
        return np.stack([-np.sin(np.pi * x[:, 1]) * np.cos(np.pi * x[:, 0]) * e,
                         np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 1]) * e], axis=1)


class analytical_pressure:                                           # TaylorGreen2D.py:148-158
    def __init__(self):
        self.t, self.nu = 0, 0.01

    def __call__(self, x):
        return -(np.cos(2 * np.pi * x[:, 0]) + np.cos(2 * np.pi * x[:, 1])) * np.exp(-4. * np.pi * np.pi * self.nu * self.t) / 4.


def top_right_point(x):                                              # TaylorGreen2D.py:160-166
    return (np.abs(x[:, 0] - 1.0) < DOLFIN_EPS) & (np.abs(x[:, 1] - 1.0) < DOLFIN_EPS)


def create_bcs(DVP, mesh, boundaries, **namespace):                  # TaylorGreen2D.py:169-183
    velocity = analytical_velocity()
    p_bc_val = analytical_pressure()
    u_bc = DirichletBC(DVP, mesh, boundaries, 1, velocity, 1)
    p_bc = PointwiseBC(DVP, mesh, 2, p_bc_val, top_right_point)
    return dict(bcs=[u_bc, p_bc], velocity=velocity, p_bc_val=p_bc_val)


def initiate(dvp_, DVP, mesh, assembler, **namespace):               # TaylorGreen2D.py:185-200
    x = np.zeros(DVP.num_dofs)
    interpolate(DVP, mesh, x, 1, analytical_velocity())
    interpolate(DVP, mesh, x, 2, analytical_pressure())
    xd = assembler.to_device(x)
    assembler.copy(dvp_["n"], xd)
    assembler.copy(dvp_["n-1"], xd)
    return dict(dvp_=dvp_, error_v_list=[], error_p_list=[])


def pre_solve(t, velocity, p_bc_val, **namespace):                   # TaylorGreen2D.py:202-209
    velocity.t = t
    p_bc_val.t = t
    return dict(velocity=velocity, p_bc_val=p_bc_val)


def post_solve(DVP, mesh, dt, dvp_, total_error_v, total_error_p, velocity, p_bc_val, assembler, verbose,
               error_v_list, error_p_list, **namespace):            # TaylorGreen2D.py:211-231
    x = assembler.to_host(dvp_["n"])
    xe = np.zeros_like(x)
    interpolate(DVP, mesh, xe, 1, velocity)
    interpolate(DVP, mesh, xe, 2, p_bc_val)
    E_v = l2_norm(mesh, DVP, xe - x, 1)
    E_p = l2_norm(mesh, DVP, xe - x, 2)
    total_error_v += E_v * dt
    total_error_p += E_p * dt
    error_v_list.append(E_v)
    error_p_list.append(E_p)
    if verbose:
        print("velocity error:", E_v)
        print("pressure error:", E_p)
    return dict(total_error_v=total_error_v, total_error_p=total_error_p)


def finished(total_error_v, total_error_p, verbose, **namespace):   # TaylorGreen2D.py:233-245 (pickle dump: I/O, out of scope)
    if verbose:
        print(f"total error for the velocity: {total_error_v:2.6e}")
        print(f"total error for the pressure: {total_error_p:2.6e}")
