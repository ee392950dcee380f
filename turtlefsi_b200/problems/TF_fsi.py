"""Turek FSI benchmark — turtleFSI/problems/TF_fsi.py re-hosted without DOLFIN: the same seven
hooks with the same names, argument meaning and returned keys; geometry markers, boundary
conditions and post-processing follow the reference line by line (cited below)."""
import os

import numpy as np

from turtlefsi_b200.boundary import DirichletBC, Facets, mark_cells, point_eval, surface_traction
from turtlefsi_b200.mesh import read_dolfin_xml
from turtlefsi_b200.problems import *  # noqa: F401,F403

DOLFIN_EPS = 3.0e-16


def near(x, x0, eps=DOLFIN_EPS):
    return np.abs(x - x0) < eps


def _mesh_file(name):
    here = os.path.dirname(os.path.abspath(__file__))
    for base in (os.environ.get("TURTLEFSI_MESH_DIR", ""),
                 os.path.join(here, "..", "..", "tests", "golden", "meshes"),
                 "/root/reference/turtleFSI/mesh"):
        p = os.path.join(base, name)
        if base and os.path.exists(p):
            return p
    raise FileNotFoundError(name)


def set_problem_parameters(default_variables, **namespace):          # TF_fsi.py:26-63
    default_variables.update(dict(
        T=30, dt=0.01, theta=0.51,
        Um=2.0, rho_f=1.0e3, mu_f=1.0, rho_s=1.0e3, nu_s=0.4, mu_s=2.0e6, lambda_s=4e6,
        folder="TF_fsi_results", extrapolation="biharmonic",
        extrapolation_sub_type="constrained_disp_vel", bc_ids=[2, 3, 4, 6],
        recompute=1,
        R=0.05, H=0.41, L=2.5, f_L=0.35, f_H=0.02, c_x=0.2, c_y=0.2))
    default_variables["compiler_parameters"] = dict(default_variables["compiler_parameters"], quadrature_degree=5)
    return default_variables


def get_mesh_domain_and_boundaries(R, H, L, f_L, f_H, c_x, c_y, **namespace):   # TF_fsi.py:66-104
    mesh = read_dolfin_xml(_mesh_file("TF_fsi.xml.gz"))
    boundaries = Facets(mesh)

    def circle(x):
        return (x[:, 0] - c_x) ** 2 + (x[:, 1] - c_y) ** 2 < R ** 2 + DOLFIN_EPS * 1e5

    boundaries.mark(lambda x, on: on, 1)                                          # DomainBoundary
    boundaries.mark(lambda x, on: near(x[:, 1], H) | near(x[:, 1], 0), 2)         # Wall
    boundaries.mark(lambda x, on: near(x[:, 0], 0), 3)                            # Inlet
    boundaries.mark(lambda x, on: near(x[:, 0], L), 4)                            # Outlet
    boundaries.mark(lambda x, on: near(x[:, 1], c_y + f_H / 2) | near(x[:, 1], c_y - f_H / 2) |
                    near(x[:, 0], c_x + R + f_L), 5)                              # Bar
    boundaries.mark(lambda x, on: circle(x), 6)                                   # Circle
    boundaries.mark(lambda x, on: circle(x) & (c_y - f_H / 2 <= x[:, 1]) & (x[:, 1] <= c_y + f_H / 2) &
                    (x[:, 0] > c_x), 7)                                           # Barwall
    mesh.markers[:] = 1
    mark_cells(mesh, lambda x: (c_y - f_H / 2 <= x[:, 1]) & (x[:, 1] <= c_y + f_H / 2) &
               (c_x <= x[:, 0]) & (x[:, 0] <= c_x + R + f_L), 2)                  # Bar_area
    return mesh, mesh.markers, boundaries


def initiate(c_x, c_y, R, f_L, **namespace):                         # TF_fsi.py:107-117
    return dict(displacement_x_list=[], displacement_y_list=[], drag_list=[], lift_list=[], time_list=[],
                coord=[c_x + R + f_L, c_y])


class Inlet:                                                           # TF_fsi.py:120-140
    def __init__(self, Um, H, **kwargs):
        self.Um = 1.5 * Um
        self.H = H
        self.factor = 0

    def update(self, t):
        if t < 2:
            self.factor = 0.5 * (1 - np.cos(t * np.pi / 2)) * self.Um
        else:
            self.factor = self.Um

    def __call__(self, x):
        return np.stack([self.factor * x[:, 1] * (self.H - x[:, 1]) / (self.H / 2.0) ** 2,
                         np.zeros(len(x))], axis=1)


def create_bcs(DVP, mesh, v_deg, Um, H, boundaries, extrapolation_sub_type, **namespace):   # TF_fsi.py:143-185
    inlet = Inlet(Um, H, degree=v_deg)

    def bc(field, value, marker):
        return DirichletBC(DVP, mesh, boundaries, field, value, marker)
    u_inlet = bc(1, inlet, 3)
    u_wall = bc(1, (0.0, 0.0), 2)
    u_circ = bc(1, (0.0, 0.0), 6)
    u_barwall = bc(1, (0.0, 0.0), 7)
    p_out = bc(2, 0, 4)
    bcs = [u_wall, u_inlet, u_circ, u_barwall, p_out]
    if extrapolation_sub_type != "constrained_disp_vel":
        for i in [bc(0, (0.0, 0.0), m) for m in (2, 3, 4, 6, 7)]:
            bcs.append(i)
    else:
        for i in [bc(3, (0.0, 0.0), m) for m in (2, 3, 4, 6, 7)] + [bc(0, (0.0, 0.0), m) for m in (2, 3, 4, 6, 7)]:
            bcs.append(i)
    return dict(bcs=bcs, inlet=inlet)


def pre_solve(t, inlet, **namespace):                                  # TF_fsi.py:188-190
    inlet.update(t)


def post_solve(t, dvp_, coord, displacement_x_list, displacement_y_list, drag_list, lift_list, mu_f,
               verbose, time_list, mesh, DVP, boundaries, assembler, **namespace):   # TF_fsi.py:223-250
    x = assembler.to_host(dvp_["n"])
    f6 = np.nonzero(boundaries.markers == 6)[0]
    f5 = np.nonzero((boundaries.markers == 5) & ~boundaries.on_boundary)[0]      # dS(5): interior facets
    f6 = f6[boundaries.on_boundary[f6]]                                            # ds(6): exterior facets
    tr = surface_traction(mesh, DVP, boundaries, x, mu_f, f6) + surface_traction(mesh, DVP, boundaries, x, mu_f, f5)
    Dr, Li = -tr[0], -tr[1]
    drag_list.append(Dr)
    lift_list.append(Li)
    time_list.append(t)
    d_eval = point_eval(mesh, DVP, x, coord, field=0)
    displacement_x_list.append(d_eval[0])
    displacement_y_list.append(d_eval[1])
    if verbose:
        print("Distance x: {:e}".format(displacement_x_list[-1]))
        print("Distance y: {:e}".format(displacement_y_list[-1]))
        print("Drag: {:e}".format(drag_list[-1]))
        print("Lift: {:e}".format(lift_list[-1]))


def finished(results_folder, displacement_x_list, displacement_y_list, drag_list, lift_list, time_list, **namespace):
    if results_folder is None:
        return
    os.makedirs(results_folder, exist_ok=True)
    np.savetxt(os.path.join(results_folder, 'Lift.txt'), lift_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'Drag.txt'), drag_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'Time.txt'), time_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'dis_x.txt'), displacement_x_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'dis_y.txt'), displacement_y_list, delimiter=',')
