"""TEST INFRASTRUCTURE — a minimal writer of MATLAB v7.3 (HDF5) files, so that the reader `mat73.py` is exercised on files
generated in the test itself (h5py is not in this image).  It emits exactly the subset of the HDF5 file format the reference's
trajectory files use (HDF5 File Format Specification, version 0 superblock): 512-byte MATLAB user block, one root group with a
version-1 B-tree + local heap + one symbol-table node, version-1 object headers with dataspace / datatype / (filter pipeline) /
layout messages, little-endian float64 datasets stored contiguously or chunked with deflate."""
import struct
import zlib

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _dataspace(dims) -> bytes:
    return struct.pack("<BBB5x", 1, len(dims), 0) + b"".join(struct.pack("<Q", d) for d in dims)


def _datatype_f64() -> bytes:
    # class 1 (floating point), version 1; little endian, IEEE double: sign 63, exponent 52..62 (bias 1023), mantissa 0..51
    return struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)


def write_mat73(path: str, variables: dict, chunked: dict | None = None):
    """variables: name -> 2-D float array (MATLAB shape).  chunked: name -> number of rows of the HDF5 dataset per chunk (the
    dataset is stored chunked with deflate, like the reference's larger file); the others contiguously."""
    chunked = chunked or {}
    base = 512
    blob = bytearray()          # everything after the user block; addresses are relative to `base`

    def alloc(b: bytes) -> int:
        while len(blob) % 8:
            blob.append(0)
        a = len(blob)
        blob.extend(b)
        return a

    blob.extend(b"\x00" * 96)   # superblock (56 bytes) + root symbol-table entry (40 bytes), filled in at the end
    names = sorted(variables)
    # local heap data: names at 8-byte aligned offsets, offset 0 is the empty string of the root
    heap_data = bytearray(b"\x00" * 8)
    name_off = {}
    for nme in names:
        name_off[nme] = len(heap_data)
        heap_data.extend(_pad8(nme.encode() + b"\x00"))
    ohdr_addr = {}
    for nme in names:
        a = np.asarray(variables[nme], dtype="<f8")
        assert a.ndim == 2
        h5 = np.ascontiguousarray(a.T)                      # HDF5 dims = MATLAB dims reversed
        msgs = _msg(0x01, _dataspace(h5.shape)) + _msg(0x03, _datatype_f64())
        if nme in chunked:
            rows = int(chunked[nme])
            cdims = (rows, h5.shape[1])
            entries = []
            for r0 in range(0, h5.shape[0], rows):
                chunk = np.zeros(cdims, dtype="<f8")
                part = h5[r0:r0 + rows]
                chunk[:part.shape[0]] = part
                raw = zlib.compress(chunk.tobytes(), 6)
                entries.append((len(raw), (r0, 0, 0), alloc(raw)))
            # version-1 B-tree node of type 1 (raw data chunks), one leaf; the last key closes the node
            node = b"TREE" + struct.pack("<BBH", 1, 0, len(entries)) + struct.pack("<QQ", _UNDEF, _UNDEF)
            for csize, offs, addr in entries:
                node += struct.pack("<II", csize, 0) + struct.pack("<3Q", *offs) + struct.pack("<Q", addr)
            node += struct.pack("<II", 0, 0) + struct.pack("<3Q", h5.shape[0], 0, 0)
            bt = alloc(node)
            # filter pipeline (version 1): one filter, deflate (id 1), one client value (level)
            msgs += _msg(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", 6) + b"\x00" * 4)
            msgs += _msg(0x08, struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", bt) + struct.pack("<3I", cdims[0], cdims[1], 8))
        else:
            data = alloc(h5.tobytes())
            msgs += _msg(0x08, struct.pack("<BB", 3, 1) + struct.pack("<QQ", data, h5.size * 8))
        nmsg = 4 if nme in chunked else 3
        ohdr_addr[nme] = alloc(struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs)) + msgs)
    heap_data_addr = alloc(bytes(heap_data))
    heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), _UNDEF, heap_data_addr))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for nme in names:
        snod += struct.pack("<QQII16x", name_off[nme], ohdr_addr[nme], 0, 0)
    snod_addr = alloc(snod)
    tree = alloc(b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", _UNDEF, _UNDEF)
                 + struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]]))
    root_ohdr = alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + _msg(0x11, struct.pack("<QQ", tree, heap)))
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", base, _UNDEF, len(blob), _UNDEF)
    sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", tree, heap)
    blob[0:96] = sb
    header = b"MATLAB 7.3 MAT-file, written by tests/hdf5_min.py".ljust(116) + b"\x00" * 8 + struct.pack("<H", 0x0200) + b"IM"
    with open(path, "wb") as f:
        f.write(header.ljust(512, b"\x00"))
        f.write(bytes(blob))
