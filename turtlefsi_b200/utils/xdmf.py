"""XDMF time series of nodal fields without HDF5 (h5py is not in this image): the light data is the usual XDMF XML
(one temporal collection, one uniform grid per stored step, geometry and topology written once and referenced), the heavy
data is ONE raw little-endian binary file beside it, addressed by `Seek` offsets (`Format="Binary"`, read by ParaView's
XDMF readers). Stand-in for DOLFIN's `XDMFFile.write(function, t)` as turtleFSI uses it
(turtleFSI/problems/__init__.py:232-237,311-323: `flush_output` = the XML is complete after every write,
`rewrite_function_mesh = False` = the mesh is stored once)."""
import os

import numpy as np

_TOPOLOGY = {(2, 3): "Triangle", (3, 4): "Tetrahedron"}


class XDMFSeries:
    def __init__(self, path, coords, cells):
        """coords float64 [n, g]; cells int [m, g+1] (triangles or tetrahedra of P1 nodal values)."""
        self.path = str(path)
        self.bin_name = os.path.splitext(os.path.basename(self.path))[0] + ".bin"
        self.bin_path = os.path.join(os.path.dirname(self.path), self.bin_name)
        coords = np.ascontiguousarray(coords, dtype="<f8")
        cells = np.ascontiguousarray(cells, dtype="<i4")
        self.gdim = coords.shape[1]
        self.num_points, self.num_cells = coords.shape[0], cells.shape[0]
        self.topology = _TOPOLOGY[(self.gdim, cells.shape[1])]
        self.nodes_per_cell = cells.shape[1]
        self.steps = []                       # (time, name, ncomp, byte offset)
        with open(self.bin_path, "wb") as f:
            f.write(cells.tobytes())
            self.geometry_offset = f.tell()
            f.write(coords.tobytes())
            self.size = f.tell()
        self._write_xml()

    def write(self, name, values, t):
        """append one step: values [num_points] (scalar) or [num_points, ncomp] (vector; 2-D vectors are padded to three
        components, as DOLFIN does, so that ParaView treats them as vectors)."""
        v = np.asarray(values, dtype="<f8")
        if v.ndim == 1:
            v = v[:, None]
        if v.shape[0] != self.num_points:
            raise ValueError("XDMFSeries.write: %d values for %d points" % (v.shape[0], self.num_points))
        if v.shape[1] == 2:
            v = np.concatenate([v, np.zeros((v.shape[0], 1))], axis=1)
        v = np.ascontiguousarray(v, dtype="<f8")
        with open(self.bin_path, "ab") as f:
            f.write(v.tobytes())
        self.steps.append((float(t), str(name), v.shape[1], self.size))
        self.size += v.nbytes
        self._write_xml()

    def _item(self, dims, dtype, precision, offset):
        return ('<DataItem Format="Binary" DataType="%s" Precision="%d" Endian="Little" Dimensions="%s" Seek="%d">%s</DataItem>'
                % (dtype, precision, dims, offset, self.bin_name))

    def _write_xml(self):
        geo_type = "XY" if self.gdim == 2 else "XYZ"
        out = ['<?xml version="1.0"?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>', '<Xdmf Version="3.0">', ' <Domain>',
               '  <Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">']
        for k, (t, name, ncomp, off) in enumerate(self.steps):
            out.append('   <Grid Name="step_%d" GridType="Uniform">' % k)
            out.append('    <Time Value="%.17g"/>' % t)
            out.append('    <Topology TopologyType="%s" NumberOfElements="%d">%s</Topology>'
                       % (self.topology, self.num_cells,
                          self._item("%d %d" % (self.num_cells, self.nodes_per_cell), "Int", 4, 0)))
            out.append('    <Geometry GeometryType="%s">%s</Geometry>'
                       % (geo_type, self._item("%d %d" % (self.num_points, self.gdim), "Float", 8, self.geometry_offset)))
            out.append('    <Attribute Name="%s" AttributeType="%s" Center="Node">%s</Attribute>'
                       % (name, "Scalar" if ncomp == 1 else "Vector",
                          self._item("%d %d" % (self.num_points, ncomp), "Float", 8, off)))
            out.append('   </Grid>')
        out += ['  </Grid>', ' </Domain>', '</Xdmf>']
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            f.write("\n".join(out) + "\n")
        os.replace(tmp, self.path)


def read_series(path):
    """parse a file written by XDMFSeries back: (coords, cells, [(t, name, values), ...]) — used by the tests and as
    a reader for post-processing scripts."""
    import re
    import xml.etree.ElementTree as ET
    text = open(path).read()
    text = re.sub(r"<!DOCTYPE[^>]*>", "", text)
    root = ET.fromstring(text)
    folder = os.path.dirname(str(path))

    def load(item):
        dims = [int(x) for x in item.get("Dimensions").split()]
        dt = np.dtype("<i4") if item.get("DataType") == "Int" else np.dtype("<f8")
        n = int(np.prod(dims))
        with open(os.path.join(folder, item.text.strip()), "rb") as f:
            f.seek(int(item.get("Seek")))
            return np.frombuffer(f.read(n * dt.itemsize), dtype=dt).reshape(dims)

    coords = cells = None
    steps = []
    for grid in root.iter("Grid"):
        if grid.get("GridType") != "Uniform":
            continue
        if cells is None:
            cells = load(grid.find("Topology").find("DataItem"))
            coords = load(grid.find("Geometry").find("DataItem"))
        att = grid.find("Attribute")
        steps.append((float(grid.find("Time").get("Value")), att.get("Name"), load(att.find("DataItem"))))
    return coords, cells, steps


def p2_refined_cells(cell_nodes, gdim):
    """uniform (red) refinement of P2 cells into P1 cells on the P2 nodes: cell_nodes [nc, 6 | 10] = vertices then edge
    mid-points in the local edge order of turtlefsi_b200.space.EDGES (FIAT's)."""
    from ..space import EDGES
    cn = np.asarray(cell_nodes)
    m = {e: gdim + 1 + k for k, e in enumerate(EDGES[gdim])}          # (lower, higher local vertex) -> local node
    if gdim == 2:
        sub = [(0, m[0, 1], m[0, 2]), (m[0, 1], 1, m[1, 2]), (m[0, 2], m[1, 2], 2), (m[0, 1], m[1, 2], m[0, 2])]
    else:
        e01, e02, e03, e12, e13, e23 = m[0, 1], m[0, 2], m[0, 3], m[1, 2], m[1, 3], m[2, 3]
        sub = [(0, e01, e02, e03), (e01, 1, e12, e13), (e02, e12, 2, e23), (e03, e13, e23, 3),
               # the inner octahedron, cut along the diagonal e02 - e13
               (e01, e02, e03, e13), (e01, e02, e12, e13), (e02, e03, e13, e23), (e02, e12, e13, e23)]
    return np.concatenate([cn[:, list(t)] for t in sub], axis=0)
