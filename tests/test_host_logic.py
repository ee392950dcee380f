"""CPU tests of the host logic: mesh generators, dof map, CSR pattern, the C-ABI library."""
import ctypes as C
import os
import re

import numpy as np

from tests.helpers import square_mesh
from turtlefsi_b200.mesh import box_mesh, tube_mesh
from turtlefsi_b200.space import MixedSpace, pattern_from_cell_dofs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_box_mesh_counts_and_orientation():
    m = box_mesh(5)
    assert m.num_cells == 6 * 125 and m.num_vertices == 216
    assert (np.diff(m.cells, axis=1) > 0).all()
    X = m.coords[m.cells]
    vol = np.abs(np.linalg.det(np.swapaxes(X[:, 1:] - X[:, :1], 1, 2))) / 6
    assert abs(vol.sum() - 1.0) < 1e-12 and vol.min() > 0
    # sizes named in BASELINE.md: 6 n^3
    assert 6 * 55 ** 3 == 998250 and 6 * 69 ** 3 == 1971054


def test_tube_markers():
    m = tube_mesh(6)
    assert set(np.unique(m.markers)) == {1, 2}
    cen = m.coords[m.cells].mean(axis=1)
    assert ((np.hypot(cen[:, 0], cen[:, 1]) < 0.8) == (m.markers == 1)).all()


def test_dofmap_and_pattern_all_numberings():
    for mesh in (tube_mesh(3), square_mesh(4)):
        for bih in (False, True):
            ref = None
            for numbering in ("interleaved", "blocked"):
                sp = MixedSpace(mesh, biharmonic=bih, numbering=numbering)
                g = mesh.gdim
                assert sp.ndof_cell == {(2, False): 27, (2, True): 39, (3, False): 64, (3, True): 94}[(g, bih)]
                # every dof is used, numbering is a bijection
                assert np.array_equal(np.unique(sp.cell_dofs), np.arange(sp.num_dofs))
                rp, ci = sp.pattern()
                rp2, ci2 = pattern_from_cell_dofs(sp.cell_dofs, sp.num_dofs)
                assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
                # sorted columns, diagonal present
                for r in (0, sp.num_dofs // 2, sp.num_dofs - 1):
                    seg = ci[rp[r]:rp[r + 1]]
                    assert (np.diff(seg) > 0).all() and r in seg
                ref = ref or len(ci)
                assert len(ci) == ref       # nnz is numbering independent
    sp = MixedSpace(box_mesh(6))
    assert 1500 < len(sp.pattern()[1]) / (6 * 216) < 1800      # SURVEY: ~1550 nnz per tet on large meshes


def test_shared_edge_dofs_consistent():
    mesh = tube_mesh(3)
    sp = MixedSpace(mesh)
    X = sp.node_coords(mesh)
    # the node of local P2 function i of a cell sits at the vertex / edge midpoint (UFC order)
    from turtlefsi_b200.space import EDGES
    c = 17
    vx = mesh.coords[mesh.cells[c]]
    loc = [vx[i] for i in range(4)] + [0.5 * (vx[i] + vx[j]) for i, j in EDGES[3]]
    assert np.allclose(X[sp.cell_nodes[c]], np.array(loc))


def test_abi_library_exports_every_declared_symbol():
    from turtlefsi_b200 import lib
    lib.build()
    so = C.CDLL(lib.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "turtlefsi_b200.h")).read()
    declared = set(re.findall(r"\b(tfsi_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(lib.EXPORTS)
    for name in declared:
        assert hasattr(so, name), name


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from turtlefsi_b200.assembler import FSIAssembler
    mesh = tube_mesh(2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FSIAssembler(mesh, MixedSpace(mesh), [], 1e-3, 0.5)
    # and the raw ABI refuses too
    from turtlefsi_b200 import lib
    L = lib.load()
    tf, ms = C.c_double(), C.c_double()
    assert L.tfsi_probe_fp64_peak(C.byref(tf), C.byref(ms)) != 0


def test_biharmonic_facet_operator_matches_oracle():
    """product facet operator (Gauss rules) vs the oracle's (Simpson / degree-6 rule), 2-D and 3-D"""
    import scipy.sparse as sp
    from oracle.facets import biharmonic_facet_matrix
    from turtlefsi_b200.boundary import Facets, biharmonic_facet_operator
    for mesh in (square_mesh(5), tube_mesh(3)):
        space = MixedSpace(mesh, biharmonic=True)
        facets = Facets(mesh)
        facets.mark(lambda x, on: on & (x[:, 0] < 0.3), 7)
        facets.mark(lambda x, on: on & (x[:, 0] > 0.7), 9)
        r, c, v = biharmonic_facet_operator(mesh, space, facets, [7, 9])
        sel = np.nonzero(np.isin(facets.markers, [7, 9]) & facets.on_boundary)[0]
        assert len(sel) > 0
        r0, c0, v0 = biharmonic_facet_matrix(mesh, space, facets.vertices[sel], facets.cell0[sel], facets.local0[sel])
        n = space.num_dofs
        A = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
        B = sp.coo_matrix((v0, (r0, c0)), shape=(n, n)).tocsr()
        assert abs(A - B).max() < 1e-14 * abs(B).max()
        # antisymmetric coupling: rows(w) x cols(d) = - rows(d) x cols(w) transposed positions
        assert abs(A).sum() > 0


def test_surface_traction_3d_constant_stress():
    """3-D drag/lift functional (common.py:63-81 sigma * n over marked facets): for a linear velocity
    v = A x, constant p and d = 0 the stress is constant, so the traction on a planar face is
    sigma n * area and the traction over the closed boundary vanishes."""
    from turtlefsi_b200.boundary import Facets, interpolate, surface_traction
    from turtlefsi_b200.mesh import box_mesh
    from turtlefsi_b200.space import MixedSpace
    mesh = box_mesh(3)
    space = MixedSpace(mesh)
    A = np.array([[0.3, -1.2, 0.5], [0.7, 0.1, -0.4], [-0.9, 0.6, 0.2]])
    mu, p0 = 0.37, 2.5
    x = np.zeros(space.num_dofs)
    interpolate(space, mesh, x, 1, lambda X: X @ A.T)
    interpolate(space, mesh, x, 2, lambda X: np.full(len(X), p0))
    sig = mu * (A + A.T) - p0 * np.eye(3)
    facets = Facets(mesh)
    allb = np.nonzero(facets.on_boundary)[0]
    assert np.abs(surface_traction(mesh, space, facets, x, mu, allb)).max() < 1e-13
    face = np.nonzero(facets.on_boundary & (np.abs(facets.midpoints[:, 0] - 1.0) < 1e-12))[0]
    assert np.allclose(surface_traction(mesh, space, facets, x, mu, face), sig @ np.array([1.0, 0, 0]), rtol=0, atol=1e-13)
    # with a deformation d = B x the stress stays constant: sigma = mu (A F^-1 + F^-T A^T) - p I
    B = 0.05 * np.array([[1.0, 0.2, 0], [0, -0.5, 0.3], [0.1, 0, 0.4]])
    interpolate(space, mesh, x, 0, lambda X: X @ B.T)
    H = A @ np.linalg.inv(np.eye(3) + B)
    sig = mu * (H + H.T) - p0 * np.eye(3)
    assert np.allclose(surface_traction(mesh, space, facets, x, mu, face), sig @ np.array([1.0, 0, 0]), rtol=0, atol=1e-13)


def test_hdf5_reader_on_the_shipped_mesh_files():
    """turtlefsi_b200.utils.hdf5_lite (no h5py in this image) on the reference's own HDF5 mesh files: the turtle_demo pair
    (tests/golden/meshes, offsets known from SURVEY Appendix C) and, when the reference tree is here, the 3-D artery of
    Test_Cylinder/problem_aneu.py:58-66 through the `dolfin` stand-in's HDF5File."""
    import os
    import numpy as np
    from turtlefsi_b200.mesh import read_turtle_demo
    from turtlefsi_b200.utils.hdf5_lite import read_datasets
    folder = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meshes", "turtle_demo")
    raw = open(os.path.join(folder, "mc.h5"), "rb").read()
    d = read_datasets(os.path.join(folder, "mc.h5"))
    assert d["/data0"].shape == (3350, 2) and d["/data1"].shape == (6578, 3) and d["/data2"].shape == (6578,)
    assert np.array_equal(d["/data0"], np.frombuffer(raw, "<f8", 6700, 2048).reshape(3350, 2))       # SURVEY Appendix C offsets
    assert np.array_equal(d["/data1"], np.frombuffer(raw, "<i8", 6578 * 3, 55648).reshape(6578, 3))
    assert np.array_equal(d["/data2"], np.frombuffer(raw, "<i8", 6578, 213520))
    mesh, fv, fm = read_turtle_demo(folder)
    assert mesh.num_cells == 6578 and set(np.unique(mesh.markers)) == {1, 2} and set(np.unique(fm)) <= {11, 12, 13, 14, 15}
    artery = "/root/reference/turtleFSI/problems/Test_Cylinder/mesh/artery_coarse_rescaled.h5"
    if not os.path.exists(artery):
        return
    from turtlefsi_b200 import compat
    if not compat.install():
        return                                           # a real DOLFIN is importable: nothing to stand in for
    import dolfin
    m = dolfin.Mesh()
    with dolfin.HDF5File(m.mpi_comm(), artery, "r") as hdf:
        hdf.read(m, "/mesh", False)
        boundaries = dolfin.MeshFunction("size_t", m, 2)
        hdf.read(boundaries, "/boundaries")
        domains = dolfin.MeshFunction("size_t", m, 3)
        hdf.read(domains, "/domains")
    assert (m.num_vertices(), m.num_cells()) == (624, 3057) and set(np.unique(domains.array())) == {1, 2}
    ids = set(np.unique(boundaries.array()))
    assert {0, 2, 3, 22, 33} <= ids                      # inlet / outlet / interface / outer wall of problem_aneu.py:23-30
    fac = m.facets()
    # the interface (22) separates fluid from solid; the outer wall (33) lies on the boundary of the mesh
    f22 = np.nonzero(boundaries.array() == 22)[0]
    assert (~fac.on_boundary[f22]).all()
    assert (domains.array()[fac.cell0[f22]] != domains.array()[fac.cell1[f22]]).all()
    assert fac.on_boundary[boundaries.array() == 33].all()
    X = m.coordinates()[m.cells()]
    assert (np.abs(np.linalg.det(X[:, 1:] - X[:, :1])) > 0).all()


def test_sass_of_the_hot_kernels_keeps_its_properties():
    """properties of the shipped SASS that the measured performance depends on (cuobjdump, no GPU needed):
    every scatter add is a fire-and-forget REDG.E.ADD.F64 — a fence behind the scatter makes ptxas emit ATOMG, whose warp
    then waits for an L2 round trip per add (the pipelined experiment); the Jacobian kernels do not spill;
    the scatter ranks arrive by TMA bulk copy; no tensor-core instruction (north_star design point 2)."""
    import collections
    import re
    import shutil
    import subprocess
    import pytest
    from turtlefsi_b200 import lib
    if shutil.which("cuobjdump") is None or not os.path.exists(lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not here")
    sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True).stdout
    mangled = re.findall(r"Function : (\S+)", sass)
    names = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    seen = {}
    for name, body in zip(names, blocks):
        ops = collections.Counter(m.group(1).split(".")[0] for m in
                                  re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", body))
        seen[name] = ops
        assert ops["HMMA"] + ops["UTCMMA"] + ops["DMMA"] == 0, name
    def one(sub):
        hit = [v for k, v in seen.items() if sub in k]
        assert len(hit) == 1, (sub, [k for k in seen if sub in k])
        return hit[0]
    fluid, solid, solid_pair = one("jac_fluid_kernel<3, 14>"), one("jac_solid_kernel<3, 14, 0, 0>"), one("jac_solid_kernel<3, 14, 0, 1>")
    residual = one("residual_kernel<3, false, 16, 0>")
    assert fluid["REDG"] == 180 and solid["REDG"] == 100 and solid_pair["REDG"] == 200 and residual["REDG"] >= 1
    for ops in (fluid, solid, solid_pair, residual):
        assert ops["ATOMG"] == 0 and ops["DFMA"] > 400
    for ops in (fluid, solid, solid_pair):
        assert ops["LDL"] == 0 and ops["STL"] == 0 and ops["UBLKCP"] >= 1 and ops["SYNCS"] >= 1
