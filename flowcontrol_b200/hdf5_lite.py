"""Minimal pure-Python HDF5 dataset reader (host-side, setup only).

The reference reads its meshes through ``dolfin.XDMFFile(...).read(mesh)``
(/root/reference/src/flowcontrol/flowsolver.py:233-240), which needs libhdf5.
Neither h5py nor libhdf5 is available to this build, so this module decodes the
small subset of the HDF5 file format that the shipped mesh files use:

* superblock version 0, 8-byte offsets/lengths,
* version-1 object headers (with continuation blocks),
* old-style groups (symbol-table message -> v1 B-tree -> SNOD + local heap),
* datasets with dataspace v1, fixed-point / IEEE float datatypes,
  layout message v3 (contiguous or chunked through a v1 chunk B-tree),
  optional deflate (id 1) and shuffle (id 2) filters.

Both layouts that occur are handled: meshio files (``/data0``, ``/data1``) and
dolfin-written files (``/Mesh/mesh/geometry``, ``/Mesh/mesh/topology``).
"""

from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5LiteError(RuntimeError):
    pass


class HDF5LiteFile:
    """Read-only view of an HDF5 file restricted to the subset described above."""

    def __init__(self, path: str | Path):
        self.path = Path(path)
        self.buf = self.path.read_bytes()
        if self.buf[:8] != _SIG:
            raise HDF5LiteError(f"{path}: not an HDF5 file")
        version = self.buf[8]
        if version != 0:
            raise HDF5LiteError(f"{path}: superblock version {version} unsupported")
        if self.buf[13] != 8 or self.buf[14] != 8:
            raise HDF5LiteError("only 8-byte offsets/lengths supported")
        # root group symbol-table entry starts at byte 56
        self.root_header = struct.unpack_from("<Q", self.buf, 56 + 8)[0]

    # -- object headers -----------------------------------------------------
    def _messages(self, addr: int) -> list[tuple[int, bytes]]:
        buf = self.buf
        version, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", buf, addr)
        if version != 1:
            raise HDF5LiteError(f"object header version {version} unsupported")
        out: list[tuple[int, bytes]] = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", buf, pos)
                data = buf[pos + 8 : pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:  # continuation
                    coff, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((coff, clen))
                out.append((mtype, data))
        return out

    # -- groups -------------------------------------------------------------
    def _heap_data_addr(self, heap_addr: int) -> int:
        if self.buf[heap_addr : heap_addr + 4] != b"HEAP":
            raise HDF5LiteError("bad local heap signature")
        return struct.unpack_from("<Q", self.buf, heap_addr + 24)[0]

    def _group_entries(self, header_addr: int) -> dict[str, int]:
        btree = heap = None
        for mtype, data in self._messages(header_addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", data, 0)
        if btree is None:
            raise HDF5LiteError("object is not an old-style group")
        heap_data = self._heap_data_addr(heap)
        entries: dict[str, int] = {}
        self._walk_group_btree(btree, heap_data, entries)
        return entries

    def _walk_group_btree(self, addr: int, heap_data: int, entries: dict[str, int]) -> None:
        buf = self.buf
        sig = buf[addr : addr + 4]
        if sig == b"TREE":
            _ntype, _level, nused = struct.unpack_from("<BBH", buf, addr + 4)
            pos = addr + 24
            for _ in range(nused):
                pos += 8  # key
                child = struct.unpack_from("<Q", buf, pos)[0]
                pos += 8
                self._walk_group_btree(child, heap_data, entries)
        elif sig == b"SNOD":
            nsym = struct.unpack_from("<H", buf, addr + 6)[0]
            pos = addr + 8
            for _ in range(nsym):
                name_off, obj = struct.unpack_from("<QQ", buf, pos)
                pos += 40
                start = heap_data + name_off
                stop = buf.index(b"\x00", start)
                entries[buf[start:stop].decode()] = obj
        else:
            raise HDF5LiteError(f"unexpected node signature {sig!r}")

    def _resolve(self, name: str) -> int:
        addr = self.root_header
        for part in [p for p in name.split("/") if p]:
            entries = self._group_entries(addr)
            if part not in entries:
                raise KeyError(f"{name!r}: {part!r} not found (have {sorted(entries)})")
            addr = entries[part]
        return addr

    def keys(self, group: str = "/") -> list[str]:
        return sorted(self._group_entries(self._resolve(group)))

    # -- datasets -----------------------------------------------------------
    def read(self, name: str) -> np.ndarray:
        addr = self._resolve(name)
        shape = dtype = layout = None
        filters: list[int] = []
        for mtype, data in self._messages(addr):
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", data, 0)
                if ver != 1:
                    raise HDF5LiteError(f"dataspace version {ver} unsupported")
                shape = struct.unpack_from(f"<{rank}Q", data, 8)
            elif mtype == 0x03:
                cls = data[0] & 0x0F
                bits0 = data[1]
                size = struct.unpack_from("<I", data, 4)[0]
                endian = ">" if (bits0 & 1) else "<"
                if cls == 0:
                    kind = "i" if (bits0 & 0x08) else "u"
                elif cls == 1:
                    kind = "f"
                else:
                    raise HDF5LiteError(f"datatype class {cls} unsupported")
                dtype = np.dtype(f"{endian}{kind}{size}")
            elif mtype == 0x08:
                layout = data
            elif mtype == 0x0B:
                ver, nfilt = struct.unpack_from("<BB", data, 0)
                if ver != 1:
                    raise HDF5LiteError(f"filter pipeline version {ver} unsupported")
                pos = 8
                for _ in range(nfilt):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", data, pos)
                    pos += 8 + ((nlen + 7) // 8) * 8 + 4 * ncd + (4 if ncd % 2 else 0)
                    filters.append(fid)
        if shape is None or dtype is None or layout is None:
            raise HDF5LiteError(f"{name}: incomplete dataset header")
        if layout[0] != 3:
            raise HDF5LiteError(f"layout version {layout[0]} unsupported")
        lclass = layout[1]
        count = int(np.prod(shape)) if len(shape) else 1
        if lclass == 1:  # contiguous
            daddr, dsize = struct.unpack_from("<QQ", layout, 2)
            arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=daddr)
            return arr.reshape(shape).astype(dtype.newbyteorder("="), copy=True)
        if lclass != 2:
            raise HDF5LiteError(f"layout class {lclass} unsupported")
        ndim = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from(f"<{ndim}I", layout, 11)
        chunk_shape = cdims[:-1]
        rank = len(shape)
        out = np.zeros(shape, dtype=dtype.newbyteorder("="))
        for csize, offs, caddr in self._chunks(btree, rank):
            raw = self.buf[caddr : caddr + csize]
            for fid in reversed(filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    n = len(raw) // dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(dtype.itemsize, n).T.tobytes()
                else:
                    raise HDF5LiteError(f"filter id {fid} unsupported")
            chunk = np.frombuffer(raw, dtype=dtype).reshape(chunk_shape)
            sel_out, sel_in = [], []
            for d in range(rank):
                lo = offs[d]
                hi = min(lo + chunk_shape[d], shape[d])
                sel_out.append(slice(lo, hi))
                sel_in.append(slice(0, hi - lo))
            out[tuple(sel_out)] = chunk[tuple(sel_in)]
        return out

    def _chunks(self, addr: int, rank: int):
        buf = self.buf
        if addr == _UNDEF:
            return
        if buf[addr : addr + 4] != b"TREE":
            raise HDF5LiteError("bad chunk B-tree signature")
        _ntype, level, nused = struct.unpack_from("<BBH", buf, addr + 4)
        keysize = 8 + 8 * (rank + 1)
        pos = addr + 24
        for _ in range(nused):
            csize, _mask = struct.unpack_from("<II", buf, pos)
            offs = struct.unpack_from(f"<{rank + 1}Q", buf, pos + 8)
            child = struct.unpack_from("<Q", buf, pos + keysize)[0]
            pos += keysize + 8
            if level == 0:
                yield csize, offs[:rank], child
            else:
                yield from self._chunks(child, rank)


def read_xdmf_mesh(xdmf_path: str | Path) -> tuple[np.ndarray, np.ndarray]:
    """Return ``(vertices[nV,2] float64, triangles[nT,3] int32)`` named by an XDMF file.

    Restates what ``dolfin.XDMFFile.read(mesh)`` yields for the reference's 2-D
    triangle meshes (flowsolver.py:233-240): the XDMF names one ``Geometry`` and
    one ``Topology`` HDF5 dataset.
    """
    import re

    xdmf_path = Path(xdmf_path)
    text = xdmf_path.read_text()
    items = re.findall(r"<DataItem[^>]*>\s*([^<\s]+)\s*</DataItem>", text)
    geo = topo = None
    # decide by the enclosing tag
    for m in re.finditer(r"<(Geometry|Topology)[^>]*>.*?<DataItem[^>]*>\s*([^<\s]+)\s*</DataItem>", text, re.S):
        if m.group(1) == "Geometry":
            geo = m.group(2)
        else:
            topo = m.group(2)
    if geo is None or topo is None:
        raise HDF5LiteError(f"{xdmf_path}: could not locate Geometry/Topology items in {items}")
    gfile, gpath = geo.split(":")
    tfile, tpath = topo.split(":")
    h5 = HDF5LiteFile(xdmf_path.parent / gfile)
    xy = np.asarray(h5.read(gpath), dtype=np.float64)[:, :2]
    if tfile != gfile:
        h5 = HDF5LiteFile(xdmf_path.parent / tfile)
    tri = np.asarray(h5.read(tpath)).astype(np.int32)
    return np.ascontiguousarray(xy), np.ascontiguousarray(tri)
