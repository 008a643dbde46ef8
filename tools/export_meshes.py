"""Convert the reference's shipped XDMF/HDF5 meshes to small .npz fixtures.

Run once in the build container (needs /root/reference); the outputs under
data/meshes/ are committed so that tests, smoke() and bench.py never read
/root/reference at run time (it does not exist on the GPU box).

    python tools/export_meshes.py
"""
from pathlib import Path
import sys

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from flowcontrol_b200.hdf5_lite import read_xdmf_mesh  # noqa: E402

REF = Path("/root/reference/src/examples")
MESHES = {
    "cylinder_O1": REF / "cylinder/data_input/O1.xdmf",
    "lidcavity_mesh64": REF / "lidcavity/data_input/mesh64.xdmf",
    "cavity_coarse": REF / "cavity/data_input/cavity_coarse.xdmf",
    "pinball_middle": REF / "pinball/data_input/mesh_middle_gmsh.xdmf",
}

if __name__ == "__main__":
    out = ROOT / "data" / "meshes"
    out.mkdir(parents=True, exist_ok=True)
    for name, path in MESHES.items():
        xy, tri = read_xdmf_mesh(path)
        np.savez_compressed(out / f"{name}.npz", vertices=xy, triangles=tri.astype(np.int32))
        print(name, xy.shape, tri.shape, (out / f"{name}.npz").stat().st_size)
