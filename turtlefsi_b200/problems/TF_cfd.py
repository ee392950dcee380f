"""Turek CFD benchmark — turtleFSI/problems/TF_cfd.py re-hosted without DOLFIN (same hooks, same
parameters; the reference lines are cited). Fluid only: `solid="no_solid"`,
`extrapolation="no_extrapolation"`, mesh refined once (TF_cfd.py:49)."""
import os

import numpy as np

from turtlefsi_b200.boundary import DirichletBC, Facets, surface_traction
from turtlefsi_b200.mesh import read_dolfin_xml, refine
from turtlefsi_b200.problems import *  # noqa: F401,F403
from turtlefsi_b200.problems.TF_fsi import DOLFIN_EPS, Inlet as _FsiInlet, _mesh_file, near


def set_problem_parameters(default_variables, **namespace):          # TF_cfd.py:21-43
    default_variables.update(dict(
        T=30, dt=0.01, theta=0.5, rho_f=1.0E3, mu_f=1.0, Um=2.0,
        folder="TF_cfd_results", solid="no_solid", extrapolation="no_extrapolation",
        H=0.41, L=2.5))
    return default_variables


def get_mesh_domain_and_boundaries(L, H, **namespace):               # TF_cfd.py:46-69
    mesh = refine(read_dolfin_xml(_mesh_file("TF_cfd.xml.gz")))
    boundaries = Facets(mesh)
    boundaries.mark(lambda x, on: on, 4)                                          # circle and flag
    boundaries.mark(lambda x, on: near(x[:, 0], 0), 1)                            # inlet
    boundaries.mark(lambda x, on: near(x[:, 1], 0) | near(x[:, 1], H), 2)         # walls
    boundaries.mark(lambda x, on: near(x[:, 0], L), 3)                            # outlet
    mesh.markers[:] = 1
    return mesh, mesh.markers, boundaries


def initiate(**namespace):                                           # TF_cfd.py:72-78
    return dict(drag_list=[], lift_list=[], time_list=[])


class Inlet(_FsiInlet):                                              # TF_cfd.py:81-99 (Um * 1.5, same profile)
    pass


def create_bcs(DVP, mesh, Um, H, v_deg, boundaries, **namespace):    # TF_cfd.py:102-114
    inlet = Inlet(Um, H, degree=v_deg)
    u_inlet = DirichletBC(DVP, mesh, boundaries, 1, inlet, 1)
    u_wall = DirichletBC(DVP, mesh, boundaries, 1, (0.0, 0.0), 2)
    u_flag = DirichletBC(DVP, mesh, boundaries, 1, (0.0, 0.0), 4)
    p_out = DirichletBC(DVP, mesh, boundaries, 2, 0, 3)
    return dict(bcs=[u_wall, u_flag, u_inlet, p_out], inlet=inlet)


def pre_solve(t, inlet, **namespace):                                # TF_cfd.py:117-119
    inlet.update(t)


def post_solve(t, dvp_, drag_list, lift_list, time_list, mu_f, verbose, mesh, DVP, boundaries, assembler,
               **namespace):                                         # TF_cfd.py:122-137
    x = assembler.to_host(dvp_["n"])
    f4 = np.nonzero((boundaries.markers == 4) & boundaries.on_boundary)[0]
    force = surface_traction(mesh, DVP, boundaries, x, mu_f, f4)
    drag_list.append(-force[0])
    lift_list.append(-force[1])
    time_list.append(t)
    if verbose:
        print("Drag: {:e}".format(drag_list[-1]))
        print("Lift: {:e}".format(lift_list[-1]))


def finished(drag_list, lift_list, time_list, results_folder, **namespace):
    if results_folder is None:
        return
    os.makedirs(results_folder, exist_ok=True)
    np.savetxt(os.path.join(results_folder, 'Lift.txt'), lift_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'Drag.txt'), drag_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'Time.txt'), time_list, delimiter=',')
