"""Writes the subset of HDF5 that libhdf5 produces for Caffe weight files (superblock 0, symbol-table groups: object header v1 +
B-tree v1 + local heap + one SNOD per group, contiguous little-endian float32 datasets), for tests of the engine's minimal reader.
The reader itself is pinned on files written by the real library (tests/golden/ref_hdf5/); this writer only adds the NESTED layout
of a .caffemodel.h5 (/data/<layer>/<blob index>), which no file of the reference's test data has.  Groups hold up to 256 links (32 symbol
nodes under one B-tree node)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Writer(object):
    def __init__(self):
        self.buf = bytearray(b"\0" * 96)            # superblock (56) + root symbol table entry (40)

    def alloc(self, data, align=8):
        while len(self.buf) % align:
            self.buf += b"\0"
        off = len(self.buf)
        self.buf += data
        return off

    def dataset(self, arr):
        arr = np.ascontiguousarray(arr, dtype="<f4")
        raw = self.alloc(arr.tobytes())
        space = struct.pack("<BBBB4x", 1, arr.ndim, 0, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        # IEEE float32 LE: class 1 version 1, bit field (byte order 0, padding, mantissa norm 2 -> 0x20, sign location 31), size 4,
        # properties: bit offset 0, precision 32, exponent location 23, size 8, mantissa location 0, size 23, bias 127
        dtype = struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        layout = struct.pack("<BBQQ", 3, 1, raw, arr.nbytes)
        msgs = b""
        for t, body in ((1, space), (3, dtype), (8, layout)):
            body += b"\0" * (-len(body) % 8)
            msgs += struct.pack("<HHB3x", t, len(body), 0) + body
        return self.alloc(struct.pack("<BBHII4x", 1, 0, 3, 1, len(msgs)) + msgs)

    def group(self, links):
        """links: {name: object header address}; -> (object header address, btree address, heap address).  Up to 32 symbol nodes of 8
        links under one level-0 B-tree node (leaf K = 4, internal K = 16)."""
        assert len(links) <= 256
        names = sorted(links)
        heap_data = bytearray(b"\0" * 8)             # offset 0: the empty name
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            b = n.encode() + b"\0"
            heap_data += b + b"\0" * (-len(b) % 8)
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, data_addr))
        chunks = [names[i:i + 8] for i in range(0, len(names), 8)] or [[]]
        entries = struct.pack("<Q", 0)               # key 0: the empty name
        for ch in chunks:
            snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for n in ch:
                snod += struct.pack("<QQII16x", offs[n], links[n], 0, 0)
            snod += b"\0" * (40 * (8 - len(ch)))
            entries += struct.pack("<QQ", self.alloc(snod), offs[ch[-1]] if ch else 0)       # child i, key i+1 = its largest name
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(chunks), UNDEF, UNDEF) + entries
        tree += b"\0" * (16 * (32 - len(chunks)))    # room for the unused children / keys
        btree = self.alloc(tree)
        body = struct.pack("<QQ", btree, heap)
        hdr = self.alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, 8 + len(body)) + struct.pack("<HHB3x", 0x11, len(body), 0) + body)
        return hdr, btree, heap

    def finish(self, root):
        hdr, btree, heap = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        self.buf[:96] = sb
        return bytes(self.buf)


def write_caffemodel_h5(layers):
    """layers: {layer name: [ndarray, ...]} -> bytes of an HDF5 file with /data/<layer>/<index> datasets (Net::ToHDF5, net.cpp:905-960)."""
    w = _Writer()
    layer_groups = {}
    for name, blobs in layers.items():
        links = {str(i): w.dataset(b) for i, b in enumerate(blobs)}
        layer_groups[name] = w.group(links)[0]
    data = w.group(layer_groups)[0]
    return w.finish(w.group({"data": data}))
