"""slam-envire_b200 -- B200-native trimmed ICP behind envire's icp/ interface.

This module is the thin Python (ctypes) binding of the C ABI in include/icp_b200.h,
used by tests/, bench.py and __graft_entry__.py.  The product itself is
libicp_b200.so (hand-written sm_100a kernels + C++ host orchestration) and the C++
host mirror in icp/icp_gpu.hpp (envire::icp::TrimmedGPU).  There is no CPU
fallback: constructing `Icp` without the built library or without a CUDA
device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ICPB_LIB: another build of the same library (kernel tuning variants, tools/build_variant.py); default = the in-tree one
LIB_PATH = os.environ.get("ICPB_LIB") or os.path.join(_HERE, "libicp_b200.so")
NO_PAIR = 0xFFFFFFFF
DEFERRED = 0xFFFFFFFE
TRACE_COLS = 28
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64

OK, ERR_CUDA, ERR_ARG, ERR_NO_MODEL, ERR_NO_SOURCE, ERR_NCCL, ERR_CAPACITY, ERR_NOT_ENOUGH_PAIRS = range(8)

# every symbol include/icp_b200.h declares
SYMBOLS = [
    "icpb_create", "icpb_destroy", "icpb_strerror", "icpb_last_error", "icpb_version", "icpb_set_option",
    "icpb_get_option", "icpb_model_add", "icpb_model_add_ean", "icpb_source_upload_ean", "icpb_model_clear", "icpb_model_size", "icpb_model_build",
    "icpb_source_upload", "icpb_align_resident", "icpb_align", "icpb_pairs_distance", "icpb_debug_pairs",
    "icpb_trace", "icpb_last_timing", "icpb_iteration_times", "icpb_find_pairs", "icpb_trim", "icpb_pairs_transform", "icpb_nccl_unique_id", "icpb_comm_init",
    "icpb_source_upload_shard", "icpb_p2p_export", "icpb_p2p_connect",
    "icpb_create_multi", "icpb_destroy_multi", "icpb_multi_last_error", "icpb_multi_size", "icpb_multi_context",
    "icpb_multi_set_option", "icpb_multi_model_add", "icpb_multi_model_clear", "icpb_multi_source_upload",
    "icpb_multi_align_resident", "icpb_multi_align", "icpb_multi_pairs_distance",
]


class Params(C.Structure):
    _fields_ = [("max_iter", C.c_size_t), ("min_mse", C.c_double), ("min_mse_diff", C.c_double),
                ("overlap", C.c_double)]


class Result(C.Structure):
    _fields_ = [("C", C.c_double * 16), ("iter", C.c_size_t), ("pairs", C.c_size_t), ("mse", C.c_double),
                ("mse_diff", C.c_double), ("d_box", C.c_double), ("overlap", C.c_double), ("status", C.c_int)]

    def matrix(self):
        return np.array(self.C, dtype=np.float64).reshape(4, 4).T.copy()


class GridInfo(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("cell_size", C.c_double), ("dims", C.c_uint32 * 3),
                ("levels", C.c_uint32), ("n_cells", C.c_uint64), ("n_points", C.c_uint64),
                ("n_occupied", C.c_uint64), ("build_ms", C.c_double), ("merges", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("upload_ms", C.c_double), ("search_ms", C.c_double),
                ("trim_ms", C.c_double), ("solve_ms", C.c_double), ("iterations", C.c_uint64),
                ("kernels_launched", C.c_uint64)]


class IcpError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("icp_b200: %s (status %d)" % (msg, status))
        self.status = status


_lib = None


def load_library():
    """dlopen libicp_b200.so and declare the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IcpError(ERR_CUDA, "libicp_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    dp, vp, up = C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_uint32)
    sp = C.POINTER(C.c_size_t)
    L.icpb_create.argtypes = [C.POINTER(vp), C.c_int]
    L.icpb_destroy.argtypes = [vp]
    L.icpb_destroy.restype = None
    L.icpb_strerror.argtypes = [C.c_int]
    L.icpb_strerror.restype = C.c_char_p
    L.icpb_last_error.argtypes = [vp]
    L.icpb_last_error.restype = C.c_char_p
    L.icpb_version.restype = C.c_char_p
    L.icpb_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.icpb_get_option.argtypes = [vp, C.c_char_p, dp]
    L.icpb_model_add.argtypes = [vp, dp, C.c_size_t, dp, C.c_double]
    L.icpb_model_add_ean.argtypes = [vp, dp, dp, C.POINTER(C.c_int32), C.c_size_t, dp, C.c_double]
    L.icpb_source_upload_ean.argtypes = [vp, dp, dp, C.POINTER(C.c_int32), C.c_size_t, dp, C.c_double]
    L.icpb_model_clear.argtypes = [vp]
    L.icpb_model_size.argtypes = [vp, sp]
    L.icpb_model_build.argtypes = [vp, C.POINTER(GridInfo)]
    L.icpb_source_upload.argtypes = [vp, dp, C.c_size_t, dp, C.c_double]
    L.icpb_align_resident.argtypes = [vp, C.POINTER(Params), C.POINTER(Result)]
    L.icpb_align.argtypes = [vp, dp, C.c_size_t, dp, C.c_double, C.POINTER(Params), C.POINTER(Result)]
    L.icpb_pairs_distance.argtypes = [vp, dp, C.c_size_t, sp]
    L.icpb_debug_pairs.argtypes = [vp, up, dp, C.POINTER(C.c_uint8), C.c_size_t, sp]
    L.icpb_trace.argtypes = [vp, dp, C.c_size_t, sp]
    L.icpb_last_timing.argtypes = [vp, C.POINTER(Timing)]
    L.icpb_iteration_times.argtypes = [vp, dp, C.c_size_t, sp]
    L.icpb_find_pairs.argtypes = [vp, dp, C.c_size_t, C.c_double, up, dp]
    L.icpb_trim.argtypes = [vp, dp, C.c_size_t, C.c_size_t, dp, sp, sp]
    L.icpb_pairs_transform.argtypes = [vp, dp, dp, dp, C.c_size_t, C.c_size_t, dp, dp, dp, sp]
    L.icpb_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.icpb_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.icpb_p2p_export.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.icpb_p2p_connect.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.icpb_source_upload_shard.argtypes = [vp, dp, C.c_size_t, C.c_size_t, dp]
    _lib = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _cm(T):
    return _d(np.asarray(np.eye(4) if T is None else T, dtype=np.float64).T.reshape(16))


def nccl_unique_id():
    buf = (C.c_uint8 * NCCL_ID_BYTES)()
    rc = load_library().icpb_nccl_unique_id(buf)
    if rc != OK:
        raise IcpError(rc, "icpb_nccl_unique_id failed")
    return bytes(buf)


class Icp:
    """One icpb_ctx: the B200 counterpart of envire::icp::TrimmedKD's model + _align (host buffers in, results out)."""

    def __init__(self, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        rc = self._L.icpb_create(C.byref(self._h), device)
        if rc != OK:
            msg = self._L.icpb_last_error(self._h).decode() if self._h else ""
            if self._h:
                self._L.icpb_destroy(self._h)
            self._h = None
            raise IcpError(rc, "icpb_create failed: %s %s" % (self._L.icpb_strerror(rc).decode(), msg))

    def close(self):
        if getattr(self, "_h", None):
            self._L.icpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise IcpError(rc, "%s: %s" % (self._L.icpb_strerror(rc).decode(),
                                           self._L.icpb_last_error(self._h).decode()))

    # -- options / model ------------------------------------------------
    def set_option(self, key, value):
        self._ck(self._L.icpb_set_option(self._h, key.encode(), float(value)))

    def add_to_model(self, xyz, T=None, density=1.0):
        a, ap = _d(xyz)
        t, tp = _cm(T)
        self._ck(self._L.icpb_model_add(self._h, ap, a.shape[0], tp, density))

    def add_to_model_ean(self, xyz, normals, attrs, T=None, density=1.0):
        a, ap = _d(xyz)
        nn, np_ = _d(normals)
        at = np.ascontiguousarray(attrs, dtype=np.int32)
        t, tp = _cm(T)
        self._ck(self._L.icpb_model_add_ean(self._h, ap, np_, at.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[0], tp,
                                            density))

    def upload_source_ean(self, xyz, normals, attrs, T=None, density=1.0):
        a, ap = _d(xyz)
        nn, np_ = _d(normals)
        at = np.ascontiguousarray(attrs, dtype=np.int32)
        t, tp = _cm(T)
        self._ck(self._L.icpb_source_upload_ean(self._h, ap, np_, at.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[0],
                                                tp, density))

    def clear_model(self):
        self._ck(self._L.icpb_model_clear(self._h))

    def model_size(self):
        n = C.c_size_t()
        self._ck(self._L.icpb_model_size(self._h, C.byref(n)))
        return n.value

    def build(self):
        gi = GridInfo()
        self._ck(self._L.icpb_model_build(self._h, C.byref(gi)))
        return gi

    # -- align ----------------------------------------------------------
    def upload_source(self, xyz, T=None, density=1.0):
        a, ap = _d(xyz)
        t, tp = _cm(T)
        self._keep = a
        self._ck(self._L.icpb_source_upload(self._h, ap, a.shape[0], tp, density))

    def upload_source_shard(self, xyz_local, adapter_size_global, T=None):
        a, ap = _d(xyz_local)
        t, tp = _cm(T)
        self._ck(self._L.icpb_source_upload_shard(self._h, ap, a.shape[0], adapter_size_global, tp))

    def align_resident(self, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, overlap=1.0):
        prm = Params(max_iter, min_mse, min_mse_diff, overlap)
        res = Result()
        self._ck(self._L.icpb_align_resident(self._h, C.byref(prm), C.byref(res)))
        return res

    def align(self, xyz, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, overlap=1.0):
        a, ap = _d(xyz)
        t, tp = _cm(T)
        prm = Params(max_iter, min_mse, min_mse_diff, overlap)
        res = Result()
        self._ck(self._L.icpb_align(self._h, ap, a.shape[0], tp, density, C.byref(prm), C.byref(res)))
        return res

    def pairs_distance(self):
        n = C.c_size_t()
        self._ck(self._L.icpb_pairs_distance(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value)
        if n.value:
            self._ck(self._L.icpb_pairs_distance(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), n.value,
                                                 C.byref(n)))
        return out[:n.value]  # after a sharded align: this rank's kept distances (fewer than the global count)

    def debug_pairs(self):
        n = C.c_size_t()
        self._ck(self._L.icpb_debug_pairs(self._h, None, None, None, 0, C.byref(n)))
        idx = np.zeros(n.value, dtype=np.uint32)
        d = np.zeros(n.value)
        kept = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self._ck(self._L.icpb_debug_pairs(self._h, idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                              d.ctypes.data_as(C.POINTER(C.c_double)),
                                              kept.ctypes.data_as(C.POINTER(C.c_uint8)), n.value, C.byref(n)))
        return idx, d, kept.astype(bool)

    def trace(self):
        n = C.c_size_t()
        self._ck(self._L.icpb_trace(self._h, None, 0, C.byref(n)))
        rows = np.zeros((n.value, TRACE_COLS))
        if n.value:
            self._ck(self._L.icpb_trace(self._h, rows.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
        out = []
        for r in rows:
            out.append(dict(iter=int(r[0]), found=int(r[1]), kept=int(r[2]), tau=r[3], d_box=r[4], mse=r[5],
                            mse_diff=r[6], far=int(r[7]), skipped=int(r[24]), deferred=int(r[25]), redone=int(r[26]), C=r[8:24].reshape(4, 4).T.copy()))
        return out

    def timing(self):
        t = Timing()
        self._ck(self._L.icpb_last_timing(self._h, C.byref(t)))
        return dict(total_ms=t.total_ms, upload_ms=t.upload_ms, search_ms=t.search_ms, trim_ms=t.trim_ms,
                    solve_ms=t.solve_ms, iterations=t.iterations, kernels_launched=t.kernels_launched)

    def iteration_times(self):
        n = C.c_size_t()
        self._ck(self._L.icpb_iteration_times(self._h, None, 0, C.byref(n)))
        rows = np.zeros((n.value, 3))
        if n.value:
            self._ck(self._L.icpb_iteration_times(self._h, rows.ctypes.data_as(C.POINTER(C.c_double)), n.value,
                                                  C.byref(n)))
        return rows

    # -- single stages ----------------------------------------------------
    def find_pairs(self, q, d_box=np.inf):
        a, ap = _d(q)
        idx = np.zeros(a.shape[0], dtype=np.uint32)
        d = np.zeros(a.shape[0])
        self._ck(self._L.icpb_find_pairs(self._h, ap, a.shape[0], d_box, idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                         d.ctypes.data_as(C.POINTER(C.c_double))))
        return idx, d

    def trim(self, dist, n_po):
        a, ap = _d(dist)
        tau = C.c_double()
        kept, ties = C.c_size_t(), C.c_size_t()
        self._ck(self._L.icpb_trim(self._h, ap, a.shape[0], n_po, C.byref(tau), C.byref(kept), C.byref(ties)))
        return tau.value, kept.value, ties.value

    def pairs_transform(self, x, p, dist, n_po):
        """Pairs::add xN + trim(n_po) + getTransform on the device.  Returns (T 4x4, mse, tau, kept)."""
        x_, xp = _d(x)
        p_, pp = _d(p)
        d_, dp_ = _d(dist)
        T = np.zeros(16)
        mse, tau, kept = C.c_double(), C.c_double(), C.c_size_t()
        rc = self._L.icpb_pairs_transform(self._h, xp, pp, dp_, len(d_), n_po, T.ctypes.data_as(C.POINTER(C.c_double)),
                                          C.byref(mse), C.byref(tau), C.byref(kept))
        if rc == ERR_NOT_ENOUGH_PAIRS:
            raise IcpError(rc, "not enough pairs to get transform")
        self._ck(rc)
        return T.reshape(4, 4).T.copy(), mse.value, tau.value, kept.value

    # -- multi-GPU ----------------------------------------------------------
    def p2p_export(self):
        buf = (C.c_uint8 * IPC_HANDLE_BYTES)()
        self._ck(self._L.icpb_p2p_export(self._h, buf))
        return bytes(buf)

    def p2p_connect(self, n_ranks, rank, handles):
        buf = (C.c_uint8 * (IPC_HANDLE_BYTES * n_ranks)).from_buffer_copy(handles)
        self._ck(self._L.icpb_p2p_connect(self._h, n_ranks, rank, buf))

    def comm_init(self, n_ranks, rank, unique_id):
        buf = (C.c_uint8 * NCCL_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(self._L.icpb_comm_init(self._h, n_ranks, rank, buf))
