"""Synthetic 3-D elastic-tube FSI (BASELINE.json config 4 at test scale): the tube mesh of
`turtlefsi_b200.mesh.tube_mesh` with the material and solver parameters of
turtleFSI/problems/Test_Cylinder/problem_aneu.py:21-64 (E_s = 1e6, nu_s = 0.45, rho_f = 1025,
mu_f = 3.5e-3, dt = 1e-3, theta = 0.501, atol = rtol = 1e-5, recompute = 3000, laplace/constant,
quadrature degree 6 for the residual — problem_aneu.py:12). Boundary conditions: ramped parabolic
inflow, fixed solid ends, p = 0 at the outlet; with `inner_pressure=True` also the interior-facet load
p(t) n('+').psi('+') dS(fsi_id) of problem_aneu.py:185-196,212-214 (InnerP ramp) on the fluid-solid
interface — off by default so that the BASELINE config stays Dirichlet-only."""
import numpy as np

from turtlefsi_b200.boundary import DirichletBC, Facets, interior_facet_normal_load, point_eval
from turtlefsi_b200.mesh import tube_mesh
from turtlefsi_b200.problems import *  # noqa: F401,F403


def set_problem_parameters(default_variables, **namespace):
    E_s_val, nu_s_val = 1E6, 0.45
    mu_s_val = E_s_val / (2 * (1 + nu_s_val))
    lambda_s_val = nu_s_val * 2. * mu_s_val / (1. - 2. * nu_s_val)
    default_variables.update(dict(
        T=0.02, dt=0.001, theta=0.501, atol=1e-5, rtol=1e-5,
        rho_f=1.025E3, mu_f=3.5E-3, rho_s=1.0E3, mu_s=mu_s_val, nu_s=nu_s_val, lambda_s=lambda_s_val,
        dx_f_id=1, dx_s_id=2, extrapolation="laplace", extrapolation_sub_type="constant",
        recompute=3000, recompute_tstep=1, folder="tube3d_results",
        ncube=5, r_fluid=0.8, length=4.0, v_max=0.2, t_ramp=0.01, residual_quadrature_degree=6,
        fsi_id=22, inner_pressure=False))
    return default_variables


def get_mesh_domain_and_boundaries(ncube, r_fluid, length, fsi_id, **namespace):
    mesh = tube_mesh(ncube, r_fluid=r_fluid, length=length)
    boundaries = Facets(mesh)
    rad = lambda x: np.hypot(x[:, 0], x[:, 1])
    boundaries.mark(lambda x, on: on, 33)                                            # outer wall
    boundaries.mark(lambda x, on: on & (np.abs(x[:, 2]) < 1e-12), 11)                # inlet plane
    boundaries.mark(lambda x, on: on & (np.abs(x[:, 2] - length) < 1e-12), 12)       # outlet plane
    # fluid part of the end planes: facets of fluid cells
    fluid_cell = mesh.markers == 1
    end_in = (boundaries.markers == 11) & fluid_cell[boundaries.cell0]
    end_out = (boundaries.markers == 12) & fluid_cell[boundaries.cell0]
    boundaries.markers[end_in] = 2                                                   # inlet_id
    boundaries.markers[end_out] = 3                                                  # outlet_id1
    interface = ~boundaries.on_boundary & (mesh.markers[boundaries.cell0] != mesh.markers[np.maximum(boundaries.cell1, 0)])
    boundaries.markers[interface] = fsi_id                                           # fluid-solid interface
    return mesh, mesh.markers, boundaries


def initiate(length, **namespace):
    return dict(probe=[0.0, 0.0, 0.5 * length], probe_v=[], probe_d=[], time_list=[])


class Inflow:
    def __init__(self, v_max, r_fluid, t_ramp):
        self.v_max, self.r, self.t_ramp, self.factor = v_max, r_fluid, t_ramp, 0.0

    def update(self, t):
        self.factor = self.v_max * (0.5 * (1 - np.cos(np.pi * min(t, self.t_ramp) / self.t_ramp)))

    def __call__(self, x):
        prof = np.maximum(0.0, 1.0 - (x[:, 0] ** 2 + x[:, 1] ** 2) / self.r ** 2)
        return np.stack([np.zeros(len(x)), np.zeros(len(x)), self.factor * prof], axis=1)


class InnerP:                                                        # problem_aneu.py:185-196
    def __init__(self, t):
        self.t = t

    def value(self):
        return self.t * (10932.4 / 0.1) if self.t < 0.1 else self.t * 10023.84667 + 9930.01533


def create_bcs(DVP, mesh, boundaries, v_max, r_fluid, t_ramp, assembler, inner_pressure, fsi_id, **namespace):
    inflow = Inflow(v_max, r_fluid, t_ramp)
    p_out_bc_val = InnerP(t=0.0)
    if inner_pressure:                                               # problem_aneu.py:210-214
        # F_solid_linear += p_out_bc_val * inner(n('+'), psi('+')) * dSS(fsi_id)
        assembler.set_facet_load(*interior_facet_normal_load(mesh, DVP, boundaries, fsi_id), scale=p_out_bc_val.value())
    zero = (0.0, 0.0, 0.0)

    def bc(field, value, marker):
        return DirichletBC(DVP, mesh, boundaries, field, value, marker)
    bcs = [bc(1, inflow, 2), bc(0, zero, 2),                  # u_inlet, d_inlet        (problem_aneu.py:226-232)
           bc(1, zero, 11), bc(0, zero, 11),                  # solid inlet end fixed
           bc(1, zero, 12), bc(0, zero, 12),                  # solid outlet end fixed
           bc(0, zero, 3), bc(2, 0.0, 3)]                     # mesh fixed and p = 0 at the outlet
    return dict(bcs=bcs, inflow=inflow, p_out_bc_val=p_out_bc_val)


def pre_solve(t, inflow, p_out_bc_val, assembler, inner_pressure, **namespace):     # problem_aneu.py:243-248
    inflow.update(t)
    p_out_bc_val.t = t
    if inner_pressure:
        assembler.set_facet_load_scale(p_out_bc_val.value())


def post_solve(t, dvp_, probe, probe_v, probe_d, time_list, mesh, DVP, assembler, **namespace):
    x = assembler.to_host(dvp_["n"])
    probe_v.append(point_eval(mesh, DVP, x, probe, field=1))
    probe_d.append(point_eval(mesh, DVP, x, [0.85, 0.0, probe[2]], field=0))
    time_list.append(t)
