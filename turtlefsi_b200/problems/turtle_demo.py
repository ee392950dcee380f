"""The default problem of turtleFSI — turtleFSI/problems/turtle_demo.py re-hosted without DOLFIN
(elastic turtle in a pulsatile flow, biharmonic lifting "constrained_disp", recompute = 15)."""
import os

import numpy as np

from turtlefsi_b200.boundary import DirichletBC, Facets
from turtlefsi_b200.mesh import read_turtle_demo
from turtlefsi_b200.problems import *  # noqa: F401,F403


def set_problem_parameters(default_variables, **namespace):          # turtle_demo.py:27-55
    default_variables.update(dict(
        T=0.2, dt=0.005, theta=0.505, Um=1.0, rho_f=1.0E3, mu_f=1.0, rho_s=1.0E3, mu_s=5.0E4, lambda_s=4.5E5,
        nu_s=0.45, dx_f_id=1, dx_s_id=2, extrapolation="biharmonic", extrapolation_sub_type="constrained_disp",
        recompute=15, checkpoint_step=10, folder="turtle_demo_results", save_step=1))
    return default_variables


def _mesh_folder():
    here = os.path.dirname(os.path.abspath(__file__))
    for base in (os.environ.get("TURTLEFSI_MESH_DIR", ""), os.path.join(here, "..", "..", "tests", "golden", "meshes"),
                 "/root/reference/turtleFSI/mesh"):
        p = os.path.join(base, "turtle_demo")
        if base and os.path.exists(os.path.join(p, "mc.h5")):
            return p
    raise FileNotFoundError("turtle_demo mesh")


def get_mesh_domain_and_boundaries(**namespace):                     # turtle_demo.py:58-87
    mesh, fverts, fmark = read_turtle_demo(_mesh_folder())
    boundaries = Facets(mesh)
    key = {(int(min(a, b)), int(max(a, b))): i for i, (a, b) in enumerate(boundaries.vertices)}
    for (a, b), m in zip(fverts, fmark):                             # MeshValueCollection -> MeshFunction
        boundaries.markers[key[(int(min(a, b)), int(max(a, b)))]] = m
    return mesh, mesh.markers, boundaries


class Inlet:                                                          # turtle_demo.py:90-110
    def __init__(self, Um, **kwargs):
        self.t, self.t_ramp, self.Um, self.value = 0.0, 0.5, Um, 0.0

    def update(self, t):
        self.t = t
        if self.t < self.t_ramp:
            self.value = self.Um * np.abs(np.cos(self.t / self.t_ramp * np.pi) - 1) / 2
        else:
            Um_min = self.Um / 6
            self.value = (self.Um - Um_min) * np.abs(np.cos(self.t / self.t_ramp * np.pi) - 1) / 2 + Um_min

    def __call__(self, x):
        return np.stack([np.full(len(x), self.value), np.zeros(len(x))], axis=1)


def create_bcs(DVP, mesh, boundaries, Um, v_deg, extrapolation_sub_type, verbose, **namespace):   # :113-148
    inlet = Inlet(Um, degree=v_deg)
    noslip = (0.0, 0.0)
    bottom_id, outlet_id, top_id, inlet_id, turtle_head_tail_id = 11, 12, 13, 14, 15

    def bc(field, value, marker, comp=None):
        return DirichletBC(DVP, mesh, boundaries, field, value, marker, comp=comp)
    u_inlet = bc(1, inlet, inlet_id)
    u_bot = bc(1, 0.0, bottom_id, comp=1)                             # DVP.sub(1).sub(1)
    u_top = bc(1, 0.0, top_id, comp=1)
    u_head_tail = bc(1, noslip, turtle_head_tail_id)
    p_outlet = bc(2, 0.0, outlet_id)
    bcs = [u_bot, u_top, u_inlet, p_outlet, u_head_tail]
    for i in [bc(0, noslip, m) for m in (bottom_id, top_id, outlet_id, inlet_id, turtle_head_tail_id)]:
        bcs.append(i)
    return dict(bcs=bcs, inlet=inlet)


def pre_solve(t, inlet, **namespace):                                # turtle_demo.py:151-153
    inlet.update(t)
