"""The C-ABI library: builds for sm_100a, loads, and exports every symbol include/fcb200.h declares."""
import ctypes
import re

import numpy as np
import pytest


def declared_functions(root):
    text = (root / "include" / "fcb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcb_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree(root):
    from flowcontrol_b200 import libfcb

    assert declared_functions(root) == sorted(libfcb.EXPORTS)


def test_library_exports_every_declared_symbol(root, built_lib):
    for name in declared_functions(root):
        assert hasattr(built_lib, name), name
    assert b"sm_100a" in built_lib.fcb_version()


def test_struct_layout_matches_header(root):
    """sizeof checks against a tiny C program compiled with gcc from the same header."""
    import subprocess
    import tempfile
    from pathlib import Path

    from flowcontrol_b200 import libfcb

    src = '#include <stdio.h>\n#include "fcb200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(fcb_plan), sizeof(fcb_problem), sizeof(fcb_controllers), sizeof(fcb_assembly), sizeof(fcb_symbolic));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "s.c"
        c.write_text(src)
        exe = Path(d) / "s"
        subprocess.run(["gcc", f"-I{root / 'include'}", str(c), "-o", str(exe)], check=True)
        sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(libfcb.fcb_plan), ctypes.sizeof(libfcb.fcb_problem), ctypes.sizeof(libfcb.fcb_controllers),
                     ctypes.sizeof(libfcb.fcb_assembly), ctypes.sizeof(libfcb.fcb_symbolic)]


def test_no_cpu_fallback(built_lib):
    """Without a GPU fcb_create must fail loudly (never silently compute on the host)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flowcontrol_b200 import libfcb

    h = ctypes.c_void_p()
    prob = libfcb.fcb_problem()
    rc = built_lib.fcb_create(ctypes.byref(prob), 4, 0, ctypes.byref(h))
    assert rc == -3 and not h.value
    assert b"no CUDA device" in built_lib.fcb_last_error(None)


def test_product_never_imports_oracle(root):
    for p in (root / "flowcontrol_b200").rglob("*.py"):
        txt = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), p
