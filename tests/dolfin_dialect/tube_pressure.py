"""A user's 3-D problem file in DOLFIN dialect that EDITS A FORM, as turtleFSI's aneurysm problems do
(Test_Cylinder/problem_aneu.py:210-214): the pressure of the surrounding tissue is added to F_solid_linear as an
interior-facet load on the fluid-solid interface, with a time-dependent coefficient updated in pre_solve. Geometry,
Dirichlet conditions and parameters are those of turtlefsi_b200/problems/tube3d.py (whose `inner_pressure=True`
variant registers the same load by hand), so the two must produce the same Newton history."""
import numpy as np
from dolfin import *                       # noqa: F401,F403

from turtleFSI.problems import *           # noqa: F401,F403
from turtlefsi_b200.mesh import tube_mesh  # the synthetic tube of BASELINE config 4 (DOLFIN has no generator for it)
from turtlefsi_b200.problems import tube3d as _native


def set_problem_parameters(default_variables, **namespace):
    default_variables = _native.set_problem_parameters(default_variables)
    default_variables["folder"] = "tube_pressure_results"
    return default_variables


def get_mesh_domain_and_boundaries(ncube, r_fluid, length, fsi_id, **namespace):
    mesh = Mesh(tube_mesh(ncube, r_fluid=r_fluid, length=length))
    domains = MeshFunction("size_t", mesh, 3)
    domains.array()[:] = mesh.native.markers
    boundaries = MeshFunction("size_t", mesh, 2, 0)
    DomainBoundary().mark(boundaries, 33)
    AutoSubDomain(lambda x, on: on & (abs(x[2]) < 1e-12)).mark(boundaries, 11)
    AutoSubDomain(lambda x, on: on & (abs(x[2] - length) < 1e-12)).mark(boundaries, 12)
    fac, marks = boundaries.facets, boundaries.array()
    fluid_side = domains.array()[fac.cell0] == 1
    marks[(marks == 11) & fluid_side] = 2
    marks[(marks == 12) & fluid_side] = 3
    other = domains.array()[np.maximum(fac.cell1, 0)]
    marks[~fac.on_boundary & (domains.array()[fac.cell0] != other)] = fsi_id
    return mesh, domains, boundaries


initiate = _native.initiate


class Inflow(UserExpression):
    def __init__(self, v_max, r_fluid, t_ramp, **kwargs):
        self.v_max, self.r, self.t_ramp, self.factor = v_max, r_fluid, t_ramp, 0.0
        super().__init__(**kwargs)

    def update(self, t):
        self.factor = self.v_max * (0.5 * (1 - np.cos(np.pi * min(t, self.t_ramp) / self.t_ramp)))

    def eval(self, value, x):
        value[0] = 0.0
        value[1] = 0.0
        value[2] = self.factor * np.maximum(0.0, 1.0 - (x[0] ** 2 + x[1] ** 2) / self.r ** 2)

    def value_shape(self):
        return (3,)


class TissuePressure(UserExpression):
    def __init__(self, t, **kwargs):
        self.t = t
        super().__init__(**kwargs)

    def eval(self, value, x):
        value[0] = self.t * (10932.4 / 0.1) if self.t < 0.1 else self.t * 10023.84667 + 9930.01533

    def value_shape(self):
        return ()


def create_bcs(DVP, mesh, boundaries, psi, F_solid_linear, v_max, r_fluid, t_ramp, fsi_id, **namespace):
    inflow = Inflow(v_max, r_fluid, t_ramp, degree=2)
    p_tissue = TissuePressure(t=0.0, degree=2)
    dSS = Measure("dS", domain=mesh, subdomain_data=boundaries)
    n = FacetNormal(mesh)
    F_solid_linear += p_tissue * inner(n('+'), psi('+')) * dSS(fsi_id)
    zero = (0.0, 0.0, 0.0)
    bcs = [DirichletBC(DVP.sub(1), inflow, boundaries, 2), DirichletBC(DVP.sub(0), zero, boundaries, 2),
           DirichletBC(DVP.sub(1), zero, boundaries, 11), DirichletBC(DVP.sub(0), zero, boundaries, 11),
           DirichletBC(DVP.sub(1), zero, boundaries, 12), DirichletBC(DVP.sub(0), zero, boundaries, 12),
           DirichletBC(DVP.sub(0), zero, boundaries, 3), DirichletBC(DVP.sub(2), 0.0, boundaries, 3)]
    return dict(bcs=bcs, inflow=inflow, p_tissue=p_tissue, F_solid_linear=F_solid_linear)


def pre_solve(t, inflow, p_tissue, **namespace):
    inflow.update(t)
    p_tissue.t = t


def post_solve(t, dvp_, probe, probe_v, probe_d, time_list, **namespace):
    d, v, p = dvp_["n"].split(deepcopy=True)
    probe_v.append(v(probe))
    probe_d.append(d([0.85, 0.0, probe[2]]))
    time_list.append(t)
