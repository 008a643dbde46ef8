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


# ---------------------------------------------------------------------------------------------------------------------
# Writer (same subset as the reader: superblock 0, version-1 object headers, old-style groups, contiguous datasets)
# ---------------------------------------------------------------------------------------------------------------------
_LEAF_K, _INTERNAL_K = 4, 16  # libhdf5 defaults: symbol-table nodes hold up to 2*4 entries, B-tree nodes up to 2*16 children


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    raise HDF5LiteError(f"datatype {dt} unsupported by the writer")


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # superblock, filled in at the end

    def alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        daddr = self.alloc(arr.tobytes()) if arr.size else _UNDEF
        space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
        msgs = [_message(0x01, space), _message(0x03, _datatype_message(arr.dtype)),
                _message(0x05, struct.pack("<BBBB", 2, 2, 2, 0)),  # fill value v2: late allocation, never written, undefined
                _message(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes))]
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def group(self, entries: dict) -> tuple[int, int, int]:
        """entries: name -> (object header address, btree, heap) with btree = heap = None for datasets.
        Returns (object header, btree, heap) addresses of the new group."""
        names = sorted(entries)  # libhdf5 keeps symbol tables sorted by name
        heap_data = bytearray(8)  # offset 0: the empty string
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap_data)
            heap_data += _pad8(nm.encode() + b"\x00")
        data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))
        chunks = [names[i : i + 2 * _LEAF_K] for i in range(0, len(names), 2 * _LEAF_K)] or [[]]
        nodes = []  # (address, heap offset of the largest name below)
        for chunk in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for nm in chunk:
                obj, bt, hp = entries[nm]
                if bt is None:
                    body += struct.pack("<QQII16x", name_off[nm], obj, 0, 0)
                else:
                    body += struct.pack("<QQIIQQ", name_off[nm], obj, 1, 0, bt, hp)
            body += b"\x00" * (8 + 2 * _LEAF_K * 40 - len(body))
            nodes.append((self.alloc(body), name_off[chunk[-1]] if chunk else 0))
        level = 0
        while True:  # B-tree levels above the symbol-table nodes: up to 2 * _INTERNAL_K children per node
            groups = [nodes[i : i + 2 * _INTERNAL_K] for i in range(0, len(nodes), 2 * _INTERNAL_K)]
            up = []
            lower = 0  # key 0 of a node: the largest name to its left (the empty string for the leftmost node)
            for grp in groups:
                node = b"TREE" + struct.pack("<BBHQQ", 0, level, len(grp), _UNDEF, _UNDEF) + struct.pack("<Q", lower)
                lower = grp[-1][1]
                for child, key in grp:
                    node += struct.pack("<QQ", child, key)
                node += b"\x00" * (24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - len(node))
                up.append((self.alloc(node), grp[-1][1]))
            nodes = up
            level += 1
            if len(nodes) == 1:
                break
        btree = nodes[0][0]
        msg = _message(0x11, struct.pack("<QQ", btree, heap_addr))
        header = self.alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(msg)) + msg)
        return header, btree, heap_addr

    def finish(self, root: tuple[int, int, int]) -> bytes:
        header, btree, heap = root
        eof = len(self.buf)
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF)
        sb += struct.pack("<QQIIQQ", 0, header, 1, 0, btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_hdf5(path: str | Path, datasets: dict[str, np.ndarray]) -> None:
    """Write ``{"/group/sub/name": array}`` as an HDF5 file (contiguous datasets in old-style groups: the layout libhdf5
    1.8/1.10 reads and the layout dolfin's own files use)."""
    tree: dict = {}
    for full, arr in datasets.items():
        parts = [p for p in full.split("/") if p]
        if not parts:
            raise HDF5LiteError("dataset needs a name")
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise HDF5LiteError(f"{full}: {p} is a dataset")
        node[parts[-1]] = np.asarray(arr)
    w = _Writer()

    def emit(node: dict) -> tuple[int, int, int]:
        entries = {}
        for name, child in node.items():
            entries[name] = emit(child) if isinstance(child, dict) else (w.dataset(child), None, None)
        return w.group(entries)

    Path(path).write_bytes(w.finish(emit(tree)))


def read_all(path: str | Path) -> dict[str, np.ndarray]:
    """Every dataset of a file as ``{"/group/name": array}`` (used to append to checkpoint files)."""
    f = HDF5LiteFile(path)
    out: dict[str, np.ndarray] = {}

    def walk(prefix: str, addr: int) -> None:
        for name, child in f._group_entries(addr).items():
            is_group = any(mtype == 0x11 for mtype, _ in f._messages(child))
            if is_group:
                walk(f"{prefix}/{name}", child)
            else:
                out[f"{prefix}/{name}"] = f.read(f"{prefix}/{name}")

    walk("", f.root_header)
    return out
