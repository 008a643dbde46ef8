"""dolfin-style XDMF/HDF5 function checkpoints (host side).

The reference saves and restores fields with ``dolfin.XDMFFile.write_checkpoint`` / ``read_checkpoint``
(/root/reference/src/utils/io.py:21-50, called from /root/reference/src/flowcontrol/exporter.py:85-165 and
flowsolver.py:599-663).  This module writes and reads that layout without dolfin or libhdf5 (hdf5_lite.py holds the
byte-level HDF5 reader/writer):

``<file>.xdmf``   one temporal ``Grid`` collection per function name; every time step is a uniform grid with the mesh
                  (``Topology`` / ``Geometry``; only the first step of a file carries it, later ones ``xi:include`` it, which
                  is what ``rewrite_function_mesh=False`` produces), a ``Time`` and a ``FiniteElementFunction`` attribute
                  with four data items
``<file>.h5``     ``/<name>/<name>_<k>/{vector, cell_dofs, x_cell_dofs, cells}`` and ``.../mesh/{topology, geometry}``

``cell_dofs`` lists, cell by cell, the global dof numbers in the element's local order -- for the vector P2 space
component-blocked ``[u_x: v0 v1 v2 e0 e1 e2 | u_y: ...]`` with the cell's vertices in ascending global order and edge
``e_i`` opposite vertex ``i`` (UFC) --, ``x_cell_dofs`` the offsets into it, ``cells`` the global cell numbers and
``vector`` the dof values.  ``read_checkpoint`` does not assume the writer's dof numbering: it matches cells by their
vertex triple and uses ``cell_dofs`` to pick every value, which is how dolfin reads files written on another partition,
so files written by the reference (any dolfin dof ordering) load as well as files written here (canonical numbering
``[u_x | u_y]`` / vertices).

An ensemble writes one function per trajectory into the same file: trajectory 0 under the reference's name (``U``,
``U_n``, ``P``) so the file is a drop-in for a single run, trajectory b under ``<name>_traj<b:04d>``.
"""

from __future__ import annotations

import re
from pathlib import Path

import numpy as np

from .hdf5_lite import HDF5LiteFile, read_all, write_hdf5
from .mesh import TaylorHoodTables

_HEADER = '<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n<Xdmf Version="3.0" xmlns:xi="http://www.w3.org/2001/XInclude">\n  <Domain>\n'
_FOOTER = "  </Domain>\n</Xdmf>\n"


def _sorted_cells(tab: TaylorHoodTables):
    """Cell vertices in ascending order and the P2 node of the edge opposite each (UFC local numbering)."""
    tri = np.asarray(tab.tri if hasattr(tab, "tri") else tab.cell_nodes[:, :3], dtype=np.int64)
    cn = np.asarray(tab.cell_nodes, dtype=np.int64)
    order = np.argsort(tri, axis=1, kind="stable")
    verts = np.take_along_axis(tri, order, axis=1)
    edge_nodes = np.take_along_axis(cn[:, 3:6], order, axis=1)  # cell_nodes[:, 3 + i] is the edge opposite local vertex i
    return verts, edge_nodes


def cell_dofs(tab: TaylorHoodTables, kind: str) -> np.ndarray:
    """[nT, dofs per cell] canonical dof numbers in dolfin's local order; kind = "V" (vector P2) or "P" (P1)."""
    verts, edge_nodes = _sorted_cells(tab)
    if kind == "P":
        return verts
    if kind == "V":
        nodes = np.concatenate([verts, edge_nodes], axis=1)
        return np.concatenate([nodes, nodes + tab.nN], axis=1)
    raise ValueError(kind)


def _attr(kind: str) -> str:
    return ('ElementFamily="CG" ElementDegree="2" ElementCell="triangle"', "Vector") if kind == "V" else \
           ('ElementFamily="CG" ElementDegree="1" ElementCell="triangle"', "Scalar")


def write_checkpoint(xdmf_path, name: str, tab: TaylorHoodTables, values: np.ndarray, kind: str, time: float,
                     append: bool = False) -> None:
    """``XDMFFile.write_checkpoint(func, name, time_step, HDF5, append)`` for one function (utils/io.py:21-39).

    ``values`` is the canonical dof vector ([u_x | u_y] nodal values for kind "V", vertex values for "P")."""
    write_checkpoints(xdmf_path, [name], tab, np.asarray(values, dtype=np.float64).reshape(-1, 1), kind, time, append=append)


def write_checkpoints(xdmf_path, names: list[str], tab: TaylorHoodTables, values: np.ndarray, kind: str, time: float,
                      append: bool = False) -> None:
    """One time step of several functions (``values[:, k]`` belongs to ``names[k]``: the trajectories of an ensemble) with
    a single rewrite of the file pair."""
    xdmf_path = Path(xdmf_path)
    h5_path = xdmf_path.with_suffix(".h5")
    xdmf_path.parent.mkdir(parents=True, exist_ok=True)
    values = np.asarray(values, dtype=np.float64)
    cd = cell_dofs(tab, kind)
    ndof = tab.Nv if kind == "V" else tab.nV
    if values.ndim != 2 or values.shape != (ndof, len(names)):
        raise ValueError(f"expected values of shape ({ndof}, {len(names)}), got {values.shape}")
    data = read_all(h5_path) if (append and h5_path.exists()) else {}
    text = xdmf_path.read_text() if (append and xdmf_path.exists()) else ""
    verts, _ = _sorted_cells(tab)
    fam, atype = _attr(kind)
    cd_flat = cd.ravel().astype(np.int64)[:, None]
    xcd = (np.arange(tab.nT + 1, dtype=np.int64) * cd.shape[1])[:, None]
    cells = np.arange(tab.nT, dtype=np.int64)[:, None]
    for k, name in enumerate(names):
        counter = len([q for q in data if re.fullmatch(rf"/{re.escape(name)}/{re.escape(name)}_\d+/vector", q)])
        first_of_file = not data
        base = f"/{name}/{name}_{counter}"
        data[f"{base}/vector"] = np.ascontiguousarray(values[:, k])[:, None]
        data[f"{base}/cell_dofs"] = cd_flat
        data[f"{base}/x_cell_dofs"] = xcd
        data[f"{base}/cells"] = cells
        if first_of_file:
            data[f"{base}/mesh/topology"] = verts.astype(np.int64)
            data[f"{base}/mesh/geometry"] = np.asarray(tab.node_xy[: tab.nV], dtype=np.float64)
            mesh_xml = (f'        <Topology NumberOfElements="{tab.nT}" TopologyType="Triangle" NodesPerElement="3">\n'
                        f'          <DataItem Dimensions="{tab.nT} 3" NumberType="UInt" Format="HDF">{h5_path.name}:{base}/mesh/topology</DataItem>\n'
                        f"        </Topology>\n"
                        f'        <Geometry GeometryType="XY">\n'
                        f'          <DataItem Dimensions="{tab.nV} 2" Format="HDF">{h5_path.name}:{base}/mesh/geometry</DataItem>\n'
                        f"        </Geometry>\n")
        else:
            first = re.search(r'<Grid Name="([^"]+)" GridType="Collection"', text).group(1)
            mesh_xml = (f'        <xi:include xpointer="xpointer(//Grid[@Name=&quot;{first}&quot;]/Grid[1]/*[self::Topology or self::Geometry])" />\n')
        grid = (f'      <Grid Name="{name}_{counter}" GridType="Uniform">\n' + mesh_xml +
                f'        <Time Value="{time!r}" />\n'
                f'        <Attribute ItemType="FiniteElementFunction" {fam} Name="{name}" Center="Other" AttributeType="{atype}">\n'
                f'          <DataItem Dimensions="{cd.size} 1" NumberType="UInt" Format="HDF">{h5_path.name}:{base}/cell_dofs</DataItem>\n'
                f'          <DataItem Dimensions="{ndof} 1" NumberType="Float" Format="HDF">{h5_path.name}:{base}/vector</DataItem>\n'
                f'          <DataItem Dimensions="{tab.nT + 1} 1" NumberType="UInt" Format="HDF">{h5_path.name}:{base}/x_cell_dofs</DataItem>\n'
                f'          <DataItem Dimensions="{tab.nT} 1" NumberType="UInt" Format="HDF">{h5_path.name}:{base}/cells</DataItem>\n'
                f"        </Attribute>\n      </Grid>\n")
        open_tag = f'    <Grid Name="{name}" GridType="Collection" CollectionType="Temporal">\n'
        if not text:
            text = _HEADER + open_tag + grid + "    </Grid>\n" + _FOOTER
        elif open_tag in text:  # another time step of a function that is already in the file
            head, tail = text.split(open_tag, 1)
            end = tail.index("    </Grid>\n")
            text = head + open_tag + tail[:end] + grid + tail[end:]
        else:  # a new function in an existing file
            text = text.replace(_FOOTER, open_tag + grid + "    </Grid>\n" + _FOOTER)
    write_hdf5(h5_path, data)
    xdmf_path.write_text(text)


def checkpoint_times(xdmf_path, name: str) -> list[float]:
    text = Path(xdmf_path).read_text()
    m = re.search(rf'<Grid Name="{re.escape(name)}" GridType="Collection".*?\n    </Grid>\n', text, re.S)
    if not m:
        raise KeyError(f"{xdmf_path}: no function named {name!r}")
    return [float(v) for v in re.findall(r'<Time Value="([^"]+)"', m.group(0))]


def read_checkpoint(xdmf_path, name: str, tab: TaylorHoodTables, kind: str, counter: int = -1) -> np.ndarray:
    """``XDMFFile.read_checkpoint(func, name, counter)`` (utils/io.py:42-50): canonical dof vector of time step ``counter``."""
    return read_checkpoints(xdmf_path, [name], tab, kind, counter)[:, 0]


def read_checkpoints(xdmf_path, names: list[str], tab: TaylorHoodTables, kind: str, counter: int = -1) -> np.ndarray:
    """Time step ``counter`` of several functions of one file (the trajectories of an ensemble): [ndof, len(names)], with the
    file opened and the cell matching done once."""
    xdmf_path = Path(xdmf_path)
    h5 = HDF5LiteFile(xdmf_path.with_suffix(".h5"))
    roots = h5.keys("/")
    mine = cell_dofs(tab, kind)
    ndof = mine.shape[1]
    out = np.full((tab.Nv if kind == "V" else tab.nV, len(names)), np.nan)
    order_cache: dict[bytes, np.ndarray] = {}
    shared_topo = None
    for col, name in enumerate(names):
        if name not in roots:
            raise KeyError(f"{xdmf_path}: no function named {name!r}")
        steps = sorted(int(k.rsplit("_", 1)[1]) for k in h5.keys(f"/{name}"))
        if not steps:
            raise KeyError(f"{xdmf_path}: function {name!r} has no time steps")
        base = f"/{name}/{name}_{steps[counter]}"
        vec = np.asarray(h5.read(f"{base}/vector"), dtype=np.float64).ravel()
        fcd = np.asarray(h5.read(f"{base}/cell_dofs")).ravel().astype(np.int64)
        xcd = np.asarray(h5.read(f"{base}/x_cell_dofs")).ravel().astype(np.int64)
        # the mesh of the file: the step's own, or the first one of the file (shared mesh)
        topo = None
        try:
            topo = np.asarray(h5.read(f"{base}/mesh/topology")).astype(np.int64)
        except KeyError:
            if shared_topo is None:
                for n in roots:
                    try:
                        shared_topo = np.asarray(h5.read(f"/{n}/{n}_0/mesh/topology")).astype(np.int64)
                        break
                    except KeyError:
                        continue
            topo = shared_topo
        key_bytes = b"" if topo is None else topo.tobytes()
        if key_bytes not in order_cache:
            if topo is None or (topo.shape == (tab.nT, 3) and np.array_equal(np.sort(topo, axis=1), _sorted_cells(tab)[0])):
                order_cache[key_bytes] = np.arange(tab.nT)  # same cells in the same order (dolfin keeps the order of the mesh file in serial)
            else:  # match cells by their vertex triple
                key = lambda t: (t[:, 0] * (tab.nV + 1) + t[:, 1]) * (tab.nV + 1) + t[:, 2]  # noqa: E731
                fk, mk = key(np.sort(topo, axis=1)), key(_sorted_cells(tab)[0])
                pos = np.argsort(fk)
                loc = np.searchsorted(fk[pos], mk)
                if np.any(loc >= len(fk)) or not np.array_equal(fk[pos][np.minimum(loc, len(fk) - 1)], mk):
                    raise ValueError(f"{xdmf_path}: the checkpoint was written on a different mesh")
                order_cache[key_bytes] = pos[loc]  # my cell c = file cell order[c]
        order = order_cache[key_bytes]
        if np.any(np.diff(xcd) != ndof):
            raise ValueError(f"{xdmf_path}: {name} is not a {kind} function (cell dof counts differ)")
        file_dofs = fcd.reshape(-1, ndof)[order]
        out[mine.ravel(), col] = vec[file_dofs.ravel()]
    if np.isnan(out).any():
        raise ValueError(f"{xdmf_path}: checkpoint does not cover every dof")
    return out


def trajectory_name(name: str, b: int) -> str:
    """Function name of trajectory ``b`` of an ensemble checkpoint (trajectory 0 keeps the reference's name)."""
    return name if b == 0 else f"{name}_traj{b:04d}"
