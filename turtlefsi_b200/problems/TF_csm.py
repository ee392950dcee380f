"""Turek CSM benchmark — turtleFSI/problems/TF_csm.py re-hosted without DOLFIN (solid only, gravity)."""
import os

import numpy as np

from turtlefsi_b200.boundary import DirichletBC, Facets, point_eval
from turtlefsi_b200.mesh import read_dolfin_xml, refine
from turtlefsi_b200.problems import *  # noqa: F401,F403
from turtlefsi_b200.problems.TF_fsi import DOLFIN_EPS, _mesh_file


def set_problem_parameters(default_variables, **namespace):          # TF_csm.py:18-43
    default_variables.update(dict(
        T=10, dt=0.01, theta=0.51,
        rho_s=1.0e3, mu_s=0.5e6, nu_s=0.4, gravity=2.0, lambda_s=2e6,
        dx_f_id=0, dx_s_id=1, folder="TF_csm_results", fluid="no_fluid", extrapolation="no_extrapolation",
        R=0.05, c_x=0.2, c_y=0.2, f_L=0.35))
    return default_variables


def get_mesh_domain_and_boundaries(c_x, c_y, R, **namespace):        # TF_csm.py:46-70
    mesh = refine(read_dolfin_xml(_mesh_file("TF_csm.xml.gz")))
    boundaries = Facets(mesh)
    boundaries.mark(lambda x, on: (x[:, 0] - c_x) ** 2 + (x[:, 1] - c_y) ** 2 < R ** 2 + DOLFIN_EPS * 1e5, 1)
    mesh.markers[:] = 1
    return mesh, mesh.markers, boundaries


def initiate(f_L, R, c_x, c_y, **namespace):                         # TF_csm.py:73-82
    return dict(displacement_x_list=[], displacement_y_list=[], time_list=[], coord=[c_x + R + f_L, c_y])


def create_bcs(DVP, mesh, boundaries, **namespace):                  # TF_csm.py:85-91
    u_barwall = DirichletBC(DVP, mesh, boundaries, 0, (0.0, 0.0), 1)
    v_barwall = DirichletBC(DVP, mesh, boundaries, 1, (0.0, 0.0), 1)
    return dict(bcs=[u_barwall, v_barwall])


def post_solve(t, dvp_, coord, displacement_x_list, displacement_y_list, time_list, verbose, mesh, DVP,
               assembler, **namespace):                              # TF_csm.py:120-133
    time_list.append(t)
    d_eval = point_eval(mesh, DVP, assembler.to_host(dvp_["n"]), coord, field=0)
    displacement_x_list.append(d_eval[0])
    displacement_y_list.append(d_eval[1])
    if verbose:
        print("Distance x: {:e}".format(d_eval[0]))
        print("Distance y: {:e}".format(d_eval[1]))


def finished(results_folder, displacement_x_list, displacement_y_list, time_list, **namespace):
    if results_folder is None:
        return
    os.makedirs(results_folder, exist_ok=True)
    np.savetxt(os.path.join(results_folder, 'Time.txt'), time_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'dis_x.txt'), displacement_x_list, delimiter=',')
    np.savetxt(os.path.join(results_folder, 'dis_y.txt'), displacement_y_list, delimiter=',')
