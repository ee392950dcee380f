"""A user's problem file written against DOLFIN's API — the Turek channel with the elastic flag (same physics and
geometry as the reference's FSI benchmark, so its results can be checked against the committed oracle history), but
written independently: SubDomain subclasses instead of AutoSubDomain lambdas, a Constant-driven inflow ramp, drag and
lift through `dot(sigma, n)`. It is run UNMODIFIED by turtlefsi_b200 through the `dolfin` stand-in
(turtlefsi_b200/compat) in tests/test_dropin.py (CPU, oracle backend) and tests/test_gpu_dropin.py (CUDA path)."""
import os

import numpy as np
from dolfin import *                       # noqa: F401,F403
from mpi4py import MPI as pyMPI           # noqa: F401

from turtleFSI.modules import *            # noqa: F401,F403  (sigma)
from turtleFSI.problems import *           # noqa: F401,F403  (default hooks)

HERE = os.path.dirname(os.path.abspath(__file__))


def set_problem_parameters(default_variables, **namespace):
    default_variables.update(dict(
        T=30, dt=0.01, theta=0.51, Um=2.0, rho_f=1.0e3, mu_f=1.0, rho_s=1.0e3, nu_s=0.4, mu_s=2.0e6, lambda_s=4e6,
        folder="turek_flag_results", extrapolation="laplace", extrapolation_sub_type="constant", recompute=1,
        geo=dict(R=0.05, H=0.41, L=2.5, flag_len=0.35, flag_h=0.02, cx=0.2, cy=0.2)))
    default_variables["compiler_parameters"].update({"quadrature_degree": 5})
    return default_variables


class Walls(SubDomain):
    def __init__(self, H):
        super().__init__()
        self.H = H

    def inside(self, x, on_boundary):
        return near(x[1], self.H) | near(x[1], 0.0)                 # array-friendly


class Cylinder(SubDomain):
    def __init__(self, geo):
        super().__init__()
        self.g = geo

    def inside(self, x, on_boundary):
        g = self.g
        return (x[0] - g["cx"]) ** 2 + (x[1] - g["cy"]) ** 2 < g["R"] ** 2 + DOLFIN_EPS * 1e5


class FlagRoot(Cylinder):
    def inside(self, x, on_boundary):
        g = self.g
        if not Cylinder.inside(self, x, on_boundary):                # python control flow: evaluated point by point
            return False
        return g["cy"] - g["flag_h"] / 2 <= x[1] <= g["cy"] + g["flag_h"] / 2 and x[0] > g["cx"]


def get_mesh_domain_and_boundaries(geo, **namespace):
    mesh = Mesh(os.path.join(HERE, "..", "golden", "meshes", "TF_fsi.xml.gz"))
    g = geo
    tip = g["cx"] + g["R"] + g["flag_len"]
    lo, hi = g["cy"] - g["flag_h"] / 2, g["cy"] + g["flag_h"] / 2
    facets = MeshFunction("size_t", mesh, mesh.geometry().dim() - 1, 0)
    DomainBoundary().mark(facets, 1)
    Walls(g["H"]).mark(facets, 2)
    AutoSubDomain(lambda x: near(x[0], 0.0)).mark(facets, 3)                                  # inflow
    AutoSubDomain(lambda x, on_boundary: near(x[0], g["L"])).mark(facets, 4)                  # outflow
    AutoSubDomain(lambda x: near(x[1], hi) or near(x[1], lo) or near(x[0], tip)).mark(facets, 5)   # flag surface
    Cylinder(g).mark(facets, 6)
    FlagRoot(g).mark(facets, 7)
    cells = MeshFunction("size_t", mesh, mesh.geometry().dim())
    cells.set_all(1)
    AutoSubDomain(lambda x: (lo <= x[1]) & (x[1] <= hi) & (g["cx"] <= x[0]) & (x[0] <= tip)).mark(cells, 2)
    return mesh, cells, facets


class Parabola(UserExpression):
    def __init__(self, peak, H, amplitude, **kwargs):
        self.peak, self.H, self.amplitude = peak, H, amplitude
        super().__init__(**kwargs)

    def eval(self, value, x):
        value[0] = float(self.amplitude) * self.peak * x[1] * (self.H - x[1]) / (self.H / 2.0) ** 2
        value[1] = 0.0

    def value_shape(self):
        return (2,)


def initiate(geo, **namespace):
    return dict(tip=[geo["cx"] + geo["R"] + geo["flag_len"], geo["cy"]], history=dict(t=[], drag=[], lift=[], dx=[], dy=[]),
                ramp=Constant(0.0))


def create_bcs(DVP, boundaries, Um, geo, v_deg, ramp, **namespace):
    inflow = Parabola(1.5 * Um, geo["H"], ramp, degree=v_deg)
    zero = Constant((0.0, 0.0))
    bcs = [DirichletBC(DVP.sub(1), zero, boundaries, 2), DirichletBC(DVP.sub(1), inflow, boundaries, 3),
           DirichletBC(DVP.sub(1), zero, boundaries, 6), DirichletBC(DVP.sub(1), zero, boundaries, 7),
           DirichletBC(DVP.sub(2), Constant(0.0), boundaries, 4)]
    bcs += [DirichletBC(DVP.sub(0), zero, boundaries, i) for i in (2, 3, 4, 6, 7)]
    return dict(bcs=bcs)


def pre_solve(t, ramp, **namespace):
    ramp.assign(0.5 * (1 - np.cos(t * np.pi / 2)) if t < 2 else 1.0)


def post_solve(t, dvp_, mu_f, n, ds, dS, tip, history, **namespace):
    d, v, p = dvp_["n"].split(deepcopy=True)[:3]
    traction = dot(sigma(v, p, d, mu_f), n)
    traction_flag = dot(sigma(v("+"), p("+"), d("+"), mu_f), n("+"))
    history["t"].append(t)
    history["drag"].append(-assemble(traction[0] * ds(6)) - assemble(traction_flag[0] * dS(5)))
    history["lift"].append(-assemble(traction[1] * ds(6)) - assemble(traction_flag[1] * dS(5)))
    at_tip = d(tip)
    history["dx"].append(at_tip[0])
    history["dy"].append(at_tip[1])


def finished(results_folder, history, **namespace):
    if MPI.rank(MPI.comm_world) == 0:
        np.savetxt(os.path.join(str(results_folder), "history.txt"),
                   np.column_stack([history[k] for k in ("t", "drag", "lift", "dx", "dy")]))
