"""ctypes front-end of the CPU parity oracle (oracle/icp_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product
(slam-envire_b200/) never imports this module.

Every function mirrors one piece of the reference (icp/icp.hpp, icp/icp.cpp);
see icp_oracle.h for the file:line map.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libicp_oracle.so")
NO_PAIR = 0xFFFFFFFF


def build(force=False):
    """Compile the oracle with gcc (and oracle/_ref when the reference tree is mounted)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("icp_oracle.c", "icp_oracle.h", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libicp_oracle.so"])
    subprocess.call(["make", "-s", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


class Params(C.Structure):
    _fields_ = [("max_iter", C.c_size_t), ("min_mse", C.c_double), ("min_mse_diff", C.c_double),
                ("overlap", C.c_double)]


class Result(C.Structure):
    _fields_ = [("C", C.c_double * 16), ("iter", C.c_size_t), ("pairs", C.c_size_t), ("mse", C.c_double),
                ("mse_diff", C.c_double), ("d_box", C.c_double), ("overlap", C.c_double)]

    def matrix(self):
        return np.array(self.C, dtype=np.float64).reshape(4, 4).T.copy()


class TraceRow(C.Structure):
    _fields_ = [("iter", C.c_double), ("found", C.c_double), ("kept", C.c_double), ("tau", C.c_double),
                ("d_box", C.c_double), ("mse", C.c_double), ("mse_diff", C.c_double), ("C", C.c_double * 16)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    dp, sp, up = C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.orc_model_new.restype = vp
    L.orc_model_free.argtypes = [vp]
    L.orc_model_clear.argtypes = [vp]
    L.orc_model_add.argtypes = [vp, dp, C.c_size_t, dp, C.c_double]
    L.orc_model_add_ean.argtypes = [vp, dp, dp, C.POINTER(C.c_int32), C.c_size_t, dp, C.c_double]
    L.orc_align_ean.argtypes = [vp, dp, dp, C.POINTER(C.c_int32), C.c_size_t, dp, C.c_double, C.POINTER(Params),
                                C.POINTER(Result), vp, C.c_int]
    L.orc_model_size.argtypes = [vp]
    L.orc_model_size.restype = C.c_size_t
    L.orc_model_points.argtypes = [vp, dp]
    L.orc_model_depth.argtypes = [vp]
    L.orc_density_indices.argtypes = [C.c_size_t, C.c_double, sp, C.c_size_t]
    L.orc_density_indices.restype = C.c_size_t
    L.orc_adapter_size.argtypes = [C.c_size_t, C.c_double]
    L.orc_adapter_size.restype = C.c_size_t
    L.orc_transform_points.argtypes = [dp, dp, C.c_size_t, dp]
    L.orc_find_pairs.argtypes = [vp, dp, C.c_size_t, C.c_double, up, dp, C.c_int]
    L.orc_find_pairs_brute.argtypes = [vp, dp, C.c_size_t, C.c_double, up, dp]
    L.orc_trim.argtypes = [dp, C.c_size_t, C.c_size_t, sp, dp]
    L.orc_trim.restype = C.c_size_t
    L.orc_get_transform.argtypes = [dp, dp, dp, sp, C.c_size_t, dp, dp]
    L.orc_pose_quat_closed.argtypes = [dp, dp]
    L.orc_build_Q.argtypes = [dp, dp]
    L.orc_run_new.restype = vp
    L.orc_run_free.argtypes = [vp]
    L.orc_align.argtypes = [vp, dp, C.c_size_t, dp, C.c_double, C.POINTER(Params), C.POINTER(Result), vp, C.c_int]
    L.orc_align_golden.argtypes = [vp, dp, C.c_size_t, dp, C.c_double, C.c_size_t, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_double, C.POINTER(Result), vp, C.c_int]
    L.orc_apply_transform.argtypes = [dp, dp, dp, dp]
    L.orc_run_pairs_distance.argtypes = [vp, dp, C.c_size_t]
    L.orc_run_pairs_distance.restype = C.c_size_t
    L.orc_run_last_pairs.argtypes = [vp, up, dp, C.c_size_t]
    L.orc_run_last_pairs.restype = C.c_size_t
    L.orc_run_trace.argtypes = [vp, C.POINTER(TraceRow), C.c_size_t]
    L.orc_run_trace.restype = C.c_size_t
    L.orc_run_timing.argtypes = [vp, dp, dp, dp]
    _lib = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _cm(T):
    """4x4 numpy (row, col) -> column-major 16 doubles."""
    return _d(np.asarray(T, dtype=np.float64).T.reshape(16))


def density_indices(n, density):
    L = lib()
    cnt = L.orc_density_indices(n, density, None, 0)
    out = np.zeros(cnt, dtype=np.uintp)
    L.orc_density_indices(n, density, out.ctypes.data_as(C.POINTER(C.c_size_t)), cnt)
    return out


def adapter_size(n, density):
    return lib().orc_adapter_size(n, density)


def transform_points(M, xyz):
    a, ap = _d(xyz)
    m, mp = _cm(M)
    out = np.empty_like(a)
    lib().orc_transform_points(mp, ap, a.shape[0], out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def pose_quat_closed(sigma):
    s, sp_ = _d(np.asarray(sigma).reshape(9))
    q = np.zeros(4)
    lib().orc_pose_quat_closed(sp_, q.ctypes.data_as(C.POINTER(C.c_double)))
    return q


def build_Q(sigma):
    s, sp_ = _d(np.asarray(sigma).reshape(9))
    q = np.zeros(16)
    lib().orc_build_Q(sp_, q.ctypes.data_as(C.POINTER(C.c_double)))
    return q.reshape(4, 4)


def trim(dist, n_po):
    d, dp = _d(dist)
    order = np.zeros(max(len(d), 1), dtype=np.uintp)
    tau = C.c_double()
    kept = lib().orc_trim(dp, len(d), n_po, order.ctypes.data_as(C.POINTER(C.c_size_t)), C.byref(tau))
    return order[:kept].copy(), tau.value


def get_transform(p, x, dist, order):
    p_, pp = _d(p)
    x_, xp = _d(x)
    d_, dp = _d(dist)
    o = np.ascontiguousarray(order, dtype=np.uintp)
    T = np.zeros(16)
    mse = C.c_double()
    rc = lib().orc_get_transform(pp, xp, dp, o.ctypes.data_as(C.POINTER(C.c_size_t)), len(o),
                                 T.ctypes.data_as(C.POINTER(C.c_double)), C.byref(mse))
    if rc != 0:
        raise RuntimeError("not enough pairs to get transform")
    return T.reshape(4, 4).T.copy(), mse.value


def apply_transform(fn_T, C_l2g, C_g2gnew):
    a, ap = _cm(fn_T)
    b, bp = _cm(C_l2g)
    c, cp = _cm(C_g2gnew)
    out = np.zeros(16)
    lib().orc_apply_transform(ap, bp, cp, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(4, 4).T.copy()


class Run:
    """Outputs of the last _align: pairs_distance, last correspondences, per-iteration trace."""

    def __init__(self):
        self._h = lib().orc_run_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_run_free(self._h)
            self._h = None

    def pairs_distance(self):
        n = lib().orc_run_pairs_distance(self._h, None, 0)
        out = np.zeros(n)
        lib().orc_run_pairs_distance(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), n)
        return out

    def last_pairs(self):
        n = lib().orc_run_last_pairs(self._h, None, None, 0)
        idx = np.zeros(n, dtype=np.uint32)
        d = np.zeros(n)
        lib().orc_run_last_pairs(self._h, idx.ctypes.data_as(C.POINTER(C.c_uint32)),
                                 d.ctypes.data_as(C.POINTER(C.c_double)), n)
        return idx, d

    def trace(self):
        n = lib().orc_run_trace(self._h, None, 0)
        rows = (TraceRow * max(n, 1))()
        lib().orc_run_trace(self._h, rows, n)
        out = []
        for r in rows[:n]:
            out.append(dict(iter=int(r.iter), found=int(r.found), kept=int(r.kept), tau=r.tau, d_box=r.d_box,
                            mse=r.mse, mse_diff=r.mse_diff, C=np.array(r.C).reshape(4, 4).T.copy()))
        return out

    def timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        lib().orc_run_timing(self._h, C.byref(a), C.byref(b), C.byref(c))
        return dict(pairs=a.value, trim=b.value, solve=c.value)


class Model:
    """FindPairsKDTree model side: accumulating, un-rebalanced kd-tree."""

    def __init__(self):
        self._h = lib().orc_model_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_model_free(self._h)
            self._h = None

    def add(self, xyz, T=None, density=1.0):
        a, ap = _d(xyz)
        t, tp = _cm(np.eye(4) if T is None else T)
        lib().orc_model_add(self._h, ap, a.shape[0], tp, density)

    def add_ean(self, xyz, normals, attrs, T=None, density=1.0):
        a, ap = _d(xyz)
        nn, np_ = _d(normals)
        at = np.ascontiguousarray(attrs, dtype=np.int32)
        t, tp = _cm(np.eye(4) if T is None else T)
        lib().orc_model_add_ean(self._h, ap, np_, at.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[0], tp, density)

    def align_ean(self, xyz, normals, attrs, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5,
                  overlap=1.0, run=None, threads=0):
        a, ap = _d(xyz)
        nn, np_ = _d(normals)
        at = np.ascontiguousarray(attrs, dtype=np.int32)
        t, tp = _cm(np.eye(4) if T is None else T)
        prm = Params(max_iter, min_mse, min_mse_diff, overlap)
        res = Result()
        rc = lib().orc_align_ean(self._h, ap, np_, at.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[0], tp, density,
                                 C.byref(prm), C.byref(res), run._h if run is not None else None, threads)
        if rc != 0:
            raise RuntimeError("model has no normals/attributes")
        return res

    def clear(self):
        lib().orc_model_clear(self._h)

    def size(self):
        return lib().orc_model_size(self._h)

    def depth(self):
        return lib().orc_model_depth(self._h)

    def points(self):
        out = np.zeros((self.size(), 3))
        lib().orc_model_points(self._h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def find_pairs(self, q, d_box=np.inf, threads=0, brute=False):
        a, ap = _d(q)
        n = a.shape[0]
        idx = np.zeros(n, dtype=np.uint32)
        d = np.zeros(n)
        ip, dp = idx.ctypes.data_as(C.POINTER(C.c_uint32)), d.ctypes.data_as(C.POINTER(C.c_double))
        if brute:
            lib().orc_find_pairs_brute(self._h, ap, n, d_box, ip, dp)
        else:
            lib().orc_find_pairs(self._h, ap, n, d_box, ip, dp, threads)
        return idx, d

    def align(self, xyz, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, overlap=1.0,
              run=None, threads=0):
        a, ap = _d(xyz)
        t, tp = _cm(np.eye(4) if T is None else T)
        prm = Params(max_iter, min_mse, min_mse_diff, overlap)
        res = Result()
        lib().orc_align(self._h, ap, a.shape[0], tp, density, C.byref(prm), C.byref(res),
                        run._h if run is not None else None, threads)
        return res

    def align_golden(self, xyz, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, alpha=0.4,
                     beta=1.0, eps=0.05, run=None, threads=0):
        a, ap = _d(xyz)
        t, tp = _cm(np.eye(4) if T is None else T)
        res = Result()
        evals = lib().orc_align_golden(self._h, ap, a.shape[0], tp, density, max_iter, min_mse, min_mse_diff, alpha,
                                       beta, eps, C.byref(res), run._h if run is not None else None, threads)
        return res, evals
