"""ctypes binding of libfcb200.so (the C ABI declared in include/fcb200.h).

PyTorch is not involved here: numpy arrays (host) or raw device pointers cross
the boundary.  The library has no CPU fallback; ``load()`` raises if the shared
object is missing, ``fcb_create`` fails if there is no GPU.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libfcb200.so"

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)

FCB_NPHASES = 7
PHASE_NAMES = ("rhs", "forward", "backward", "post", "spmm", "element", "measure")


class fcb_plan(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("nU", C.c_int32), ("nblocks", C.c_int32),
        ("blk_K", c_i32p), ("blk_M", c_i32p), ("blk_nsrc", c_i32p), ("blk_out0", c_i32p), ("blk_ystore", c_i32p),
        ("blk_iptr", c_i64p), ("blk_vptr", c_i64p), ("blk_eptr", c_i64p),
        ("i0", c_i32p), ("i1", c_i32p), ("i2", c_i32p), ("e0", c_i32p), ("e1", c_i32p), ("vals", c_f64p),
        ("nlaunch", C.c_int32), ("launch_ptr", c_i32p), ("n_forward_launches", C.c_int32),
        ("asm_n", C.c_int32), ("asm_ptr", c_i32p), ("asm_src", c_i32p), ("asm_dst", c_i32p), ("asm_lptr", c_i32p),
        ("ntier", C.c_int32), ("tier_ptr", c_i32p),
        ("ncluster", C.c_int32), ("cl_fptr", c_i32p), ("cl_ustore", c_i32p), ("cl_iptr", c_i64p),
        ("imp_src", c_i32p), ("imp_dst", c_i32p),
        ("nfront", C.c_int32), ("fr_c0", c_i32p), ("fr_w", c_i32p), ("fr_m", c_i32p), ("fr_sptr", c_i64p), ("fr_struct", c_i32p),
        ("fr_eptr", c_i64p), ("fr_bptr", c_i64p), ("cl_vals", c_f64p),
    ]


class fcb_problem(C.Structure):
    _fields_ = [
        ("nT", C.c_int32), ("nN", C.c_int32), ("nV", C.c_int32),
        ("cell_nodes", c_i32p), ("Jinv", c_f64p), ("detJ", c_f64p), ("node_xy", c_f64p),
        ("n_free", C.c_int32), ("perm", c_i32p),
        ("n_bc", C.c_int32), ("bc_dofs", c_i32p),
        ("na", C.c_int32), ("bc_shape", c_f64p), ("ctrl_rhs", c_f64p * 2),
        ("plan", fcb_plan * 2),
        ("ns", C.c_int32), ("sensor_ptr", c_i32p), ("sensor_idx", c_i32p), ("sensor_val", c_f64p),
        ("dt", C.c_double), ("nonlinear", C.c_int32),
        ("scheme", C.c_int32), ("cn_ptr", c_i32p), ("cn_idx", c_i32p), ("cn_val", c_f64p), ("ctrl_rhs_prev", c_f64p),
    ]


class fcb_assembly(C.Structure):
    _fields_ = [
        ("nT", C.c_int32), ("nN", C.c_int32),
        ("cell_nodes", c_i32p), ("Jinv", c_f64p), ("detJ", c_f64p),
        ("ncolour", C.c_int32), ("colour_ptr", c_i32p), ("colour_cells", c_i32p),
        ("nnz", C.c_int32), ("pos", c_i32p),
    ]


class fcb_symbolic(C.Structure):
    _fields_ = [
        ("nfront", C.c_int32), ("w", c_i32p), ("m", c_i32p),
        ("nlevel", C.c_int32), ("level_ptr", c_i32p), ("level_fronts", c_i32p),
        ("a_ptr", c_i64p), ("a_src", c_i32p), ("a_dst", c_i32p),
        ("c_ptr", c_i64p), ("c_front", c_i32p), ("c_lptr", c_i64p), ("c_loc", c_i32p),
    ]


class fcb_controllers(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nu", C.c_int32),
        ("Ad", c_f64p), ("Bd", c_f64p), ("Cd", c_f64p), ("Dd", c_f64p), ("x0", c_f64p),
        ("Ky", c_f64p), ("Fu", c_f64p),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "fcb_version": (C.c_char_p, []),
    "fcb_last_error": (C.c_char_p, [C.c_void_p]),
    "fcb_create": (C.c_int, [C.POINTER(fcb_problem), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "fcb_destroy": (C.c_int, [C.c_void_p]),
    "fcb_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "fcb_set_controllers": (C.c_int, [C.c_void_p, C.POINTER(fcb_controllers)]),
    "fcb_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fcb_run_closed_loop": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "fcb_run_open_loop": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fcb_set_controller_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fcb_get_fields": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "fcb_get_measurement": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fcb_get_controller_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fcb_get_costs": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fcb_assemble_advection": (C.c_int, [C.POINTER(fcb_assembly), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fcb_factorize": (C.c_int, [C.POINTER(fcb_symbolic), C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fcb_profile_step": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "fcb_launch_count": (C.c_int64, [C.c_void_p]),
    "fcb_stream": (C.c_void_p, [C.c_void_p]),
    "fcb_synchronize": (C.c_int, [C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load libfcb200.so and declare every entry point of include/fcb200.h."""
    global _lib
    if _lib is None:
        import os

        global LIB_PATH
        if os.environ.get("FCB_LIB"):  # tuning builds of the same library (tools/spmm_variants.sh); never a different code path
            LIB_PATH = Path(os.environ["FCB_LIB"]).resolve()
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class FcbError(RuntimeError):
    pass


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def as_voidp(x) -> C.c_void_p:
    """numpy array, int device pointer, torch tensor (via data_ptr) or None -> void*."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


class ProblemPack:
    """Keeps the numpy arrays referenced by an ``fcb_problem`` alive."""

    def __init__(self, prob, batch: int | None = None):
        tab = prob.tab
        plans = prob.plans if batch is None else prob.plans_for_batch(batch)
        keep = self.keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a

        s = fcb_problem()
        s.nT, s.nN, s.nV = tab.nT, tab.nN, tab.nV
        s.cell_nodes = _ptr(arr(tab.cell_nodes, np.int32), c_i32p)
        s.Jinv = _ptr(arr(tab.Jinv.reshape(-1, 4), np.float64), c_f64p)
        s.detJ = _ptr(arr(tab.detJ, np.float64), c_f64p)
        s.node_xy = _ptr(arr(tab.node_xy, np.float64), c_f64p)
        s.n_free = prob.sym.n
        s.perm = _ptr(arr(prob.sym.perm, np.int32), c_i32p)
        s.n_bc = len(prob.dirichlet.dofs)
        s.bc_dofs = _ptr(arr(prob.dirichlet.dofs, np.int32), c_i32p)
        s.na = prob.na
        s.bc_shape = _ptr(arr(prob.dirichlet.shape, np.float64), c_f64p)
        for o in (1, 2):
            s.ctrl_rhs[o - 1] = _ptr(arr(prob.ctrl_rhs[o], np.float64), c_f64p)
            p, q = plans[o], s.plan[o - 1]
            q.n, q.nU, q.nblocks = p.n, p.nU, len(p.blk_K)
            for name in ("blk_K", "blk_M", "blk_nsrc", "blk_out0", "blk_ystore", "i0", "i1", "i2", "e0", "e1"):
                setattr(q, name, _ptr(arr(getattr(p, name), np.int32), c_i32p))
            for name in ("blk_iptr", "blk_vptr", "blk_eptr"):
                setattr(q, name, _ptr(arr(getattr(p, name), np.int64), c_i64p))
            q.vals = _ptr(arr(p.vals, np.float64), c_f64p)
            q.n_forward_launches = p.n_forward_launches
            q.nlaunch = len(p.launch_ptr) - 1
            q.launch_ptr = _ptr(arr(p.launch_ptr, np.int32), c_i32p)
            q.asm_n = len(p.asm_dst)
            q.asm_ptr = _ptr(arr(p.asm_ptr, np.int32), c_i32p)
            q.asm_src = _ptr(arr(p.asm_src if len(p.asm_src) else np.zeros(1), np.int32), c_i32p)
            q.asm_dst = _ptr(arr(p.asm_dst if len(p.asm_dst) else np.zeros(1), np.int32), c_i32p)
            q.asm_lptr = _ptr(arr(p.asm_lptr, np.int32), c_i32p)
            pad = lambda a: a if len(a) else np.zeros(1)  # noqa: E731
            q.ntier = len(p.tier_ptr) - 1
            q.ncluster = len(p.cl_fptr) - 1
            q.nfront = len(p.fr_c0)
            for name in ("tier_ptr", "cl_fptr", "cl_ustore", "imp_src", "imp_dst", "fr_c0", "fr_w", "fr_m", "fr_struct"):
                setattr(q, name, _ptr(arr(pad(getattr(p, name)), np.int32), c_i32p))
            for name in ("cl_iptr", "fr_sptr", "fr_eptr", "fr_bptr"):
                setattr(q, name, _ptr(arr(pad(getattr(p, name)), np.int64), c_i64p))
            q.cl_vals = _ptr(arr(pad(p.cl_vals), np.float64), c_f64p)
        s.ns = prob.ns
        s.sensor_ptr = _ptr(arr(prob.sensor_ptr, np.int32), c_i32p)
        s.sensor_idx = _ptr(arr(prob.sensor_idx, np.int32), c_i32p)
        s.sensor_val = _ptr(arr(prob.sensor_val, np.float64), c_f64p)
        s.dt = prob.dt
        s.nonlinear = int(prob.nonlinear)
        s.scheme = 1 if prob.time_scheme == "cn" else 0
        if prob.time_scheme == "cn":
            # explicit operator E in solver row order: row r = canonical dof perm[r] (velocity rows only)
            import scipy.sparse as sp

            n, Nv = prob.sym.n, tab.Nv
            perm = np.asarray(prob.sym.perm)
            vel = np.flatnonzero(perm < Nv)
            sel = sp.csr_matrix((np.ones(len(vel)), (vel, perm[vel])), shape=(n, Nv))
            E = (sel @ prob.E_cn).tocsr()
            E.sort_indices()
            s.cn_ptr = _ptr(arr(E.indptr, np.int32), c_i32p)
            s.cn_idx = _ptr(arr(E.indices, np.int32), c_i32p)
            s.cn_val = _ptr(arr(E.data, np.float64), c_f64p)
        s.ctrl_rhs_prev = _ptr(arr(prob.ctrl_rhs_prev, np.float64), c_f64p)
        self.struct = s
