"""Minimal pure-Python HDF5 reader -- just enough for the files the reference path consumes.

``loaders.Channels`` in the reference reads MATLAB v7.3 ``.mat`` files through ``hdf5storage``
(reference ``src/score_based_channels/loaders.py:4,29``), which is not installed here (nor is
``h5py``).  Supported subset (everything `matlab/generate_data.m:38` and the shipped
``sample_data`` files use): superblock v0 (optionally behind a 512-byte MATLAB user block), v1 object
headers with continuation blocks, v1 group B-trees + local heaps, dataspace v1/v2, fixed-point /
floating-point / compound{real,imag} datatypes, contiguous and chunked (v1 chunk B-tree) layouts,
deflate and shuffle filters.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


class Dataset:
    def __init__(self, f: "File", name: str, shape, dtype, layout, filters):
        self.f, self.name, self.shape, self.dtype, self.layout, self.filters = f, name, tuple(shape), dtype, layout, filters

    def read(self) -> np.ndarray:
        f = self.f
        n = int(np.prod(self.shape)) if self.shape else 1
        if self.layout[0] == "contiguous":
            _, addr, size = self.layout
            if addr == UNDEF:
                return np.zeros(self.shape, self.dtype)
            raw = f.buf[f.base + addr:f.base + addr + n * self.dtype.itemsize]
            return np.frombuffer(raw, self.dtype, n).reshape(self.shape).copy()
        if self.layout[0] == "compact":
            return np.frombuffer(self.layout[1], self.dtype, n).reshape(self.shape).copy()
        _, btree, cdims = self.layout
        out = np.zeros(self.shape, self.dtype)
        rank = len(self.shape)
        for offs, addr, size, mask in f._chunks(btree, rank):
            raw = f.buf[f.base + addr:f.base + addr + size]
            for i, (fid, cd) in reversed(list(enumerate(self.filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:      # shuffle
                    es = cd[0] if cd else self.dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    raw = a.reshape(es, -1).T.tobytes()
                else:
                    raise H5Error("unsupported filter %d" % fid)
            chunk = np.frombuffer(raw, self.dtype, int(np.prod(cdims))).reshape(cdims)
            sl_out, sl_in = [], []
            for d in range(rank):
                lo = offs[d]
                hi = min(lo + cdims[d], self.shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
        return out


class File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        sig = b"\x89HDF\r\n\x1a\n"
        off = 0
        while off < len(self.buf) and self.buf[off:off + 8] != sig:
            off = 512 if off == 0 else off * 2
        if off >= len(self.buf):
            raise H5Error("not an HDF5 file: %s" % path)
        b = self.buf
        ver = b[off + 8]
        if ver not in (0, 1):
            raise H5Error("superblock version %d not supported" % ver)
        self.so, self.sl = b[off + 13], b[off + 14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets/lengths supported")
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", b, p)[0]
        root = p + 32
        # root symbol-table entry: name off, header addr, cache type, reserved, scratch (btree, heap)
        _, hdr, ctype, _, bt, heap = struct.unpack_from("<QQIIQQ", b, root)
        self.datasets: Dict[str, Dataset] = {}
        self._walk_group(hdr, "", (bt, heap) if ctype == 1 else None)

    # -- groups -------------------------------------------------------------
    def _walk_group(self, hdr: int, prefix: str, cached: Optional[Tuple[int, int]]):
        if cached is None:
            msgs = self._messages(hdr)
            st = [m for m in msgs if m[0] == 0x11]
            if not st:
                return
            bt, heap = struct.unpack_from("<QQ", st[0][1], 0)
        else:
            bt, heap = cached
        hp = self.base + heap
        if self.buf[hp:hp + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data_addr = struct.unpack_from("<Q", self.buf, hp + 24)[0]
        for name_off, ohdr, ctype, scratch in self._group_entries(bt):
            s = self.base + data_addr + name_off
            name = self.buf[s:self.buf.index(b"\x00", s)].decode()
            full = prefix + "/" + name if prefix else name
            if ctype == 1:
                self._walk_group(ohdr, full, struct.unpack_from("<QQ", scratch, 0))
            else:
                self._object(ohdr, full)

    def _group_entries(self, bt: int):
        p = self.base + bt
        b = self.buf
        if b[p:p + 4] != b"TREE":
            raise H5Error("bad group B-tree")
        ntype, level, used = b[p + 4], b[p + 5], struct.unpack_from("<H", b, p + 6)[0]
        q = p + 8 + 16
        for i in range(used):
            child = struct.unpack_from("<Q", b, q + 8 + i * 16)[0]
            if level > 0:
                yield from self._group_entries(child)
            else:
                s = self.base + child
                if b[s:s + 4] != b"SNOD":
                    raise H5Error("bad symbol node")
                n = struct.unpack_from("<H", b, s + 6)[0]
                for j in range(n):
                    e = s + 8 + j * 40
                    name_off, ohdr, ctype, _ = struct.unpack_from("<QQII", b, e)
                    yield name_off, ohdr, ctype, b[e + 24:e + 40]

    # -- object headers -------------------------------------------------------
    def _messages(self, hdr: int) -> List[Tuple[int, bytes]]:
        b = self.buf
        p = self.base + hdr
        if b[p] != 1:
            raise H5Error("only version-1 object headers supported")
        nmsg = struct.unpack_from("<H", b, p + 2)[0]
        size = struct.unpack_from("<I", b, p + 8)[0]
        blocks = [(p + 16, size)]
        out: List[Tuple[int, bytes]] = []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", b, q)
                body = b[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x10:
                    addr, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.base + addr, ln))
                out.append((mtype, body))
        return out

    def _object(self, hdr: int, name: str):
        msgs = self._messages(hdr)
        kinds = {m[0] for m in msgs}
        if 0x11 in kinds:           # a group without cached scratch
            self._walk_group(hdr, name, None)
            return
        if not {0x1, 0x3, 0x8} <= kinds:
            return
        shape = dtype = layout = None
        filters: List[Tuple[int, Tuple[int, ...]]] = []
        for mtype, body in msgs:
            if mtype == 0x1:
                ver, rank, flags = body[0], body[1], body[2]
                q = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, body, q)
            elif mtype == 0x3:
                try:
                    dtype = self._dtype(body)[0]
                except H5Error:
                    return          # references, strings, ...: not needed on this path
            elif mtype == 0x8:
                ver = body[0]
                if ver != 3:
                    raise H5Error("layout message version %d not supported" % ver)
                cls = body[1]
                if cls == 0:
                    sz = struct.unpack_from("<H", body, 2)[0]
                    layout = ("compact", bytes(body[4:4 + sz]))
                elif cls == 1:
                    addr, sz = struct.unpack_from("<QQ", body, 2)
                    layout = ("contiguous", addr, sz)
                else:
                    rank1 = body[2]
                    addr = struct.unpack_from("<Q", body, 3)[0]
                    dims = struct.unpack_from("<%dI" % rank1, body, 11)
                    layout = ("chunked", addr, tuple(dims[:-1]))
            elif mtype == 0xB:
                ver, nf = body[0], body[1]
                q = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid, nlen, _, ncd = struct.unpack_from("<HHHH", body, q) if ver == 1 or struct.unpack_from("<H", body, q)[0] >= 256 \
                        else (struct.unpack_from("<H", body, q)[0], 0, *struct.unpack_from("<HH", body, q + 2))
                    if ver == 1 or fid >= 256:
                        q += 8
                        if ver == 1:
                            nlen = (nlen + 7) // 8 * 8
                        q += nlen
                    else:
                        q += 6
                    cd = struct.unpack_from("<%dI" % ncd, body, q)
                    q += 4 * ncd
                    if ver == 1 and ncd % 2:
                        q += 4
                    filters.append((fid, cd))
        if shape is not None and dtype is not None and layout is not None:
            self.datasets[name] = Dataset(self, name, shape, dtype, layout, filters)

    def _dtype(self, body: bytes, q: int = 0):
        cls_ver = body[q]
        cls = cls_ver & 0xF
        bits0 = body[q + 1]
        size = struct.unpack_from("<I", body, q + 4)[0]
        bo = ">" if bits0 & 1 else "<"
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype("%s%s%d" % (bo, "i" if signed else "u", size)), q + 8 + 4
        if cls == 1:
            return np.dtype("%sf%d" % (bo, size)), q + 8 + 12
        if cls == 6:
            nmemb = bits0 | (body[q + 2] << 8)
            ver = cls_ver >> 4
            p = q + 8
            names, fmts, offs = [], [], []
            for _ in range(nmemb):
                e = body.index(b"\x00", p)
                nm = body[p:e].decode()
                if ver < 3:
                    p = p + ((e - p) // 8 + 1) * 8
                else:
                    p = e + 1
                if ver == 1:
                    off = struct.unpack_from("<I", body, p)[0]
                    p += 4 + 1 + 3 + 4 + 4 + 16
                elif ver == 2:
                    off = struct.unpack_from("<I", body, p)[0]
                    p += 4
                else:
                    nb = max(1, (size.bit_length() + 7) // 8)
                    off = int.from_bytes(body[p:p + nb], "little")
                    p += nb
                dt, p = self._dtype(body, p)
                names.append(nm); fmts.append(dt); offs.append(off)
            return np.dtype({"names": names, "formats": fmts, "offsets": offs, "itemsize": size}), p
        raise H5Error("datatype class %d not supported" % cls)

    # -- chunk B-tree ---------------------------------------------------------
    def _chunks(self, bt: int, rank: int):
        p = self.base + bt
        b = self.buf
        if b[p:p + 4] != b"TREE" or b[p + 4] != 1:
            raise H5Error("bad chunk B-tree")
        level, used = b[p + 5], struct.unpack_from("<H", b, p + 6)[0]
        ksz = 8 + 8 * (rank + 1)
        q = p + 8 + 16
        for i in range(used):
            k = q + i * (ksz + 8)
            size, mask = struct.unpack_from("<II", b, k)
            offs = struct.unpack_from("<%dQ" % (rank + 1), b, k + 8)
            child = struct.unpack_from("<Q", b, k + ksz)[0]
            if level > 0:
                yield from self._chunks(child, rank)
            else:
                yield offs[:rank], child, size, mask

    def __getitem__(self, name: str) -> np.ndarray:
        return self.datasets[name].read()

    def keys(self):
        return self.datasets.keys()


def loadmat_v73(path: str) -> Dict[str, np.ndarray]:
    """``hdf5storage.loadmat`` look-alike for numeric arrays: MATLAB stores arrays column-major, so the
    on-disk HDF5 dimensions are reversed (``loaders.py:29-30`` sees ``output_h`` as [N, 10, Nr, Nt]);
    compound {real, imag} becomes a complex array."""
    f = File(path)
    out = {}
    for k in f.keys():
        if k.startswith("#"):
            continue
        a = f[k]
        if a.dtype.names and set(a.dtype.names) == {"real", "imag"}:
            a = a["real"] + 1j * a["imag"]
        elif a.dtype.names and set(a.dtype.names) == {"r", "i"}:
            a = a["r"] + 1j * a["i"]
        out[k] = np.transpose(a)
    return out
