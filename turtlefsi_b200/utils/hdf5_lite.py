"""Minimal reader of the HDF5 files the reference ships and reads through DOLFIN (`HDF5File.read`,
turtleFSI/problems/Test_Cylinder/problem_aneu.py:58-66; `XDMFFile.read`, turtle_demo.py:58-87). h5py is not in this
image, and the files are small and plain: superblock version 0, old-style groups (symbol-table B-trees + local heaps),
version-1 object headers, fixed-point / IEEE datasets stored contiguously or in (optionally deflated) chunks. Anything
else raises. `read_datasets(path)` returns {"/group/name": numpy array}."""
import struct
import zlib

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class _File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("%s is not an HDF5 file" % path)
        if b[8] != 0:
            raise NotImplementedError("HDF5 superblock version %d (only version 0 files are read here)" % b[8])
        self.O, self.L = b[13], b[14]
        if (self.O, self.L) != (8, 8):
            raise NotImplementedError("HDF5 offsets / lengths of %d / %d bytes" % (self.O, self.L))
        self.base = self.u64(24)
        root = 24 + 4 * 8                      # root group symbol-table entry
        self.root_header = self.u64(root + 8)

    def u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def u64(self, o):
        return struct.unpack_from("<Q", self.b, o)[0]

    # ---- object headers (version 1) ----------------------------------------------------------------------
    def messages(self, addr):
        b = self.b
        addr += self.base
        if b[addr] != 1:
            raise NotImplementedError("HDF5 object header version %d" % b[addr])
        nmsg, size = self.u16(addr + 2), self.u32(addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            o, n = blocks.pop(0)
            end = o + n
            while o + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u16(o), self.u16(o + 2)
                data = o + 8
                if mtype == 0x0010:                                   # continuation block
                    blocks.append((self.base + self.u64(data), self.u64(data + 8)))
                out.append((mtype, data, msize))
                o = data + msize
        return out

    # ---- groups ------------------------------------------------------------------------------------------
    def group_entries(self, btree, heap):
        b = self.b
        heap += self.base
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self.base + self.u64(heap + 8 + 2 * self.L)
        out = []

        def name(off):
            s = heap_data + off
            return b[s:b.index(b"\0", s)].decode()

        def walk(node):
            node += self.base
            if b[node:node + 4] == b"SNOD":
                for k in range(self.u16(node + 6)):
                    e = node + 8 + k * 40
                    out.append((name(self.u64(e)), self.u64(e + 8)))
                return
            if b[node:node + 4] != b"TREE" or b[node + 4] != 0:
                raise ValueError("bad group B-tree node")
            used = self.u16(node + 6)
            o = node + 8 + 2 * self.O
            for k in range(used):
                walk(self.u64(o + self.L + k * (self.L + self.O)))
        walk(btree)
        return out

    # ---- datasets ----------------------------------------------------------------------------------------
    def dataset(self, msgs):
        b = self.b
        shape = dtype = layout = None
        deflate = False
        for mtype, o, n in msgs:
            if mtype == 0x0001:                                       # dataspace
                ver, rank = b[o], b[o + 1]
                start = o + (8 if ver == 1 else 4)
                shape = tuple(self.u64(start + 8 * k) for k in range(rank))
            elif mtype == 0x0003:                                     # datatype
                cls, bits0, size = b[o] & 15, b[o + 1], self.u32(o + 4)
                if bits0 & 1:
                    raise NotImplementedError("big-endian HDF5 data")
                if cls == 0:
                    dtype = np.dtype("<%s%d" % ("i" if bits0 & 8 else "u", size))
                elif cls == 1:
                    dtype = np.dtype("<f%d" % size)
                else:
                    dtype = None                                      # strings etc.: not needed here
            elif mtype == 0x0008:                                     # layout
                if b[o] != 3:
                    raise NotImplementedError("HDF5 data layout message version %d" % b[o])
                if b[o + 1] == 1:
                    layout = ("contiguous", self.u64(o + 2), self.u64(o + 10))
                elif b[o + 1] == 2:
                    nd = b[o + 2]
                    layout = ("chunked", self.u64(o + 3), tuple(self.u32(o + 11 + 4 * k) for k in range(nd)))
                elif b[o + 1] == 0:
                    layout = ("compact", o + 4, self.u16(o + 2))
            elif mtype == 0x000B:                                     # filter pipeline
                nf = b[o + 1]
                p = o + 8
                for _ in range(nf):
                    fid, nlen, ncv = self.u16(p), self.u16(p + 2), self.u16(p + 6)
                    if fid != 1:
                        raise NotImplementedError("HDF5 filter %d (only deflate is read here)" % fid)
                    deflate = True
                    p += 8 + (nlen + 7) // 8 * 8 + 4 * ncv + (4 if ncv % 2 else 0)
        if shape is None or dtype is None or layout is None:
            return None
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "contiguous":
            if layout[1] == _UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(b, dtype, count, self.base + layout[1]).reshape(shape).copy()
        if layout[0] == "compact":
            return np.frombuffer(b, dtype, count, layout[1]).reshape(shape).copy()
        return self.chunked(layout[1], layout[2], shape, dtype, deflate)

    def chunked(self, btree, cdims, shape, dtype, deflate):
        b = self.b
        rank = len(shape)
        chunk = cdims[:rank]
        out = np.zeros(shape, dtype)

        def walk(node):
            node += self.base
            if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
                raise ValueError("bad chunk B-tree node")
            level, used = b[node + 5], self.u16(node + 6)
            ksz = 8 + 8 * (rank + 1)
            o = node + 8 + 2 * self.O
            for k in range(used):
                key = o + k * (ksz + self.O)
                nbytes = self.u32(key)
                if self.u32(key + 4):
                    raise NotImplementedError("HDF5 chunk with a filter mask")
                offs = tuple(self.u64(key + 8 + 8 * d) for d in range(rank))
                child = self.u64(key + ksz)
                if level > 0:
                    walk(child)
                    continue
                raw = b[self.base + child:self.base + child + nbytes]
                if deflate:
                    raw = zlib.decompress(raw)
                block = np.frombuffer(raw, dtype, int(np.prod(chunk))).reshape(chunk)
                sl = tuple(slice(offs[d], min(offs[d] + chunk[d], shape[d])) for d in range(rank))
                out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        if btree != _UNDEF:
            walk(btree)
        return out


def read_datasets(path):
    """every numeric dataset of the file: {"/mesh/coordinates": array, ...}."""
    f = _File(path)
    out = {}

    def visit(header, prefix):
        msgs = f.messages(header)
        sym = [(o, n) for t, o, n in msgs if t == 0x0011]
        if sym:
            o = sym[0][0]
            for name, child in f.group_entries(f.u64(o), f.u64(o + 8)):
                visit(child, prefix + "/" + name)
            return
        a = f.dataset(msgs)
        if a is not None:
            out[prefix] = a
    visit(f.root_header, "")
    return out
