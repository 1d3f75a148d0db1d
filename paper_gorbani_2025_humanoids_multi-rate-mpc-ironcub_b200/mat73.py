"""Minimal pure-Python reader for MATLAB v7.3 (HDF5, superblock v0) trajectory files.

Host-side I/O of the drop-in: stands in for matio's ``Mat_VarRead`` as used by the reference's
``TrajectoryManager::loadTrajectoryFromFile`` (utils/src/TrajectoryManager.cpp:67-140), so that the
``trajectoryFile`` entries of ``vs_mcp_config.xml`` (:34-40) can be loaded without matio / h5py.
``tests/golden/make_fixtures.py`` uses it to convert the reference's ``src/trajectories/*.mat`` into
``tests/golden/trajectories.npz``.

Supports exactly what those two files need: v0 superblock, v1 object headers, v1 group B-trees +
local heaps, contiguous and chunked (deflate) layouts of little-endian float64 datasets.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class Mat73:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        sig = b"\x89HDF\r\n\x1a\n"
        self.base = self.buf.find(sig)
        if self.base < 0:
            raise ValueError("not an HDF5 / MAT v7.3 file")
        sb = self.base
        ver = self.buf[sb + 8]
        if ver != 0:
            raise ValueError(f"unsupported superblock version {ver}")
        so, sl = self.buf[sb + 13], self.buf[sb + 14]
        if (so, sl) != (8, 8):
            raise ValueError("only 8-byte offsets/lengths supported")
        # root symbol-table entry follows: 8 sig + 8 versions + 4 K's + 4 flags + 4*8 addresses
        self.base_addr = struct.unpack_from("<Q", self.buf, sb + 24)[0]
        root = sb + 24 + 32
        _, ohdr, cache, _ = struct.unpack_from("<QQII", self.buf, root)
        if cache != 1:
            raise ValueError("root group without cached symbol table")
        btree, heap = struct.unpack_from("<QQ", self.buf, root + 24)
        self.vars = {}
        for name, addr in self._group_entries(btree, heap):
            self.vars[name] = addr

    # ---- low level ---------------------------------------------------------------------------
    def _a(self, rel: int) -> int:
        return rel + self.base_addr

    def _heap_string(self, heap_addr: int, off: int) -> str:
        h = self._a(heap_addr)
        assert self.buf[h:h + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", self.buf, h + 24)[0]
        s = self._a(data_addr) + off
        e = self.buf.index(b"\x00", s)
        return self.buf[s:e].decode()

    def _group_entries(self, btree_addr: int, heap_addr: int):
        t = self._a(btree_addr)
        assert self.buf[t:t + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", self.buf, t + 4)
        assert ntype == 0
        p = t + 8 + 16
        for i in range(used):
            child = struct.unpack_from("<Q", self.buf, p + 8 + i * 16)[0]
            if level > 0:
                yield from self._group_entries(child, heap_addr)
            else:
                s = self._a(child)
                assert self.buf[s:s + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", self.buf, s + 6)[0]
                for k in range(nsym):
                    e = s + 8 + k * 40
                    name_off, ohdr = struct.unpack_from("<QQ", self.buf, e)
                    yield self._heap_string(heap_addr, name_off), ohdr

    def _messages(self, ohdr_addr: int):
        o = self._a(ohdr_addr)
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.buf, o)
        assert ver == 1
        blocks = [(o + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, p)
                body = self.buf[p + 8:p + 8 + msize]
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self._a(coff), clen))
                out.append((mtype, body))
                p += 8 + msize
        return out

    # ---- datasets ----------------------------------------------------------------------------
    def read(self, name: str) -> np.ndarray:
        """Return the variable with MATLAB's shape (column-major semantics restored)."""
        msgs = self._messages(self.vars[name])
        dims = None
        layout = None
        deflate = False
        for mtype, body in msgs:
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", body, 0)
                off = 8 if ver == 1 else 4
                dims = struct.unpack_from(f"<{rank}Q", body, off)
            elif mtype == 0x03:
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                if cls != 1 or size != 8:
                    raise ValueError(f"{name}: only float64 datasets supported")
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                deflate = True
        if dims is None or layout is None:
            raise ValueError(f"{name}: not a simple dataset")
        n = int(np.prod(dims)) if dims else 1
        ver, lclass = layout[0], layout[1]
        assert ver == 3
        if lclass == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            a = self._a(addr)
            arr = np.frombuffer(self.buf[a:a + n * 8], dtype="<f8").copy()
        elif lclass == 2:
            rank = layout[2]
            btree = struct.unpack_from("<Q", layout, 3)[0]
            cdims = struct.unpack_from(f"<{rank}I", layout, 11)[:-1]
            arr = np.zeros(dims, dtype=np.float64)
            for offs, raw in self._chunks(btree, rank):
                if deflate:
                    raw = zlib.decompress(raw)
                chunk = np.frombuffer(raw, dtype="<f8").reshape(cdims)
                sl_dst = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, cdims, dims))
                sl_src = tuple(slice(0, s.stop - s.start) for s in sl_dst)
                arr[sl_dst] = chunk[sl_src]
            arr = arr.reshape(-1)
        else:  # compact: data stored inside the layout message
            size = struct.unpack_from("<H", layout, 2)[0]
            arr = np.frombuffer(layout[4:4 + size], dtype="<f8").copy()
        # HDF5 dims are MATLAB dims reversed (C order vs Fortran order)
        return arr.reshape(dims).T.copy()

    def _chunks(self, btree_addr: int, rank: int):
        t = self._a(btree_addr)
        assert self.buf[t:t + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", self.buf, t + 4)
        assert ntype == 1
        keysz = 8 + 8 * rank
        p = t + 8 + 16
        for i in range(used):
            k = p + i * (keysz + 8)
            csize, _mask = struct.unpack_from("<II", self.buf, k)
            offs = struct.unpack_from(f"<{rank}Q", self.buf, k + 8)[:-1]
            child = struct.unpack_from("<Q", self.buf, k + keysz)[0]
            if level > 0:
                yield from self._chunks(child, rank)
            else:
                a = self._a(child)
                yield offs, self.buf[a:a + csize]


def loadmat73(path: str) -> dict:
    m = Mat73(path)
    out = {}
    for name in m.vars:
        if name.startswith("#"):
            continue
        out[name] = m.read(name)
    return out
