"""ctypes front-end of oracle/_ref/libicp_ref.so: the REFERENCE's own icp/icp.cpp + icp/icp.hpp compiled against
restated third-party shims (see oracle/ref_driver.cpp).  TEST INFRASTRUCTURE; exists only where /root/reference was
mounted at build time (the dev container) or where the prebuilt .so travelled to."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libicp_ref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp, sp, vp = C.POINTER(C.c_double), C.POINTER(C.c_size_t), C.c_void_p
        L.ref_new.restype = vp
        L.ref_free.argtypes = [vp]
        L.ref_clear_model.argtypes = [vp]
        L.ref_add_model.argtypes = [vp, dp, C.c_size_t, dp, C.c_double]
        L.ref_align.argtypes = [vp, dp, C.c_size_t, dp, C.c_double, C.c_size_t, C.c_double, C.c_double, C.c_double, dp,
                                dp, C.c_size_t, sp]
        L.ref_align_golden.argtypes = [vp, dp, C.c_size_t, dp, C.c_double, C.c_size_t, C.c_double, C.c_double,
                                       C.c_double, C.c_double, C.c_double, dp, dp, C.c_size_t, sp]
        ip = C.POINTER(C.c_int32)
        L.ref_add_model_ean.argtypes = [vp, dp, dp, ip, C.c_size_t, dp, C.c_double]
        L.ref_align_ean.argtypes = [vp, dp, dp, ip, C.c_size_t, dp, C.c_double, C.c_size_t, C.c_double, C.c_double,
                                    C.c_double, dp, dp, C.c_size_t, sp]
        L.ref_pairs.argtypes = [dp, dp, dp, C.c_size_t, C.c_size_t, dp, dp, dp]
        L.ref_pairs.restype = C.c_long
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _cm(T):
    return _d(np.asarray(np.eye(4) if T is None else T, dtype=np.float64).T.reshape(16))


def _unpack(out, pd, n_pd):
    return dict(fn_T=out[:16].reshape(4, 4).T.copy(), iter=int(out[16]), pairs=int(out[17]), mse=out[18],
                mse_diff=out[19], overlap=out[20], pairs_distance=pd[:n_pd].copy())


class Ref:
    """envire::icp::TrimmedKD inside a private envire::Environment."""

    def __init__(self):
        self._h = lib().ref_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_free(self._h)
            self._h = None

    def add_to_model(self, xyz, T=None, density=1.0):
        a, ap = _d(xyz)
        t, tp = _cm(T)
        lib().ref_add_model(self._h, ap, a.shape[0], tp, density)

    def clear_model(self):
        lib().ref_clear_model(self._h)

    def add_to_model_ean(self, xyz, normals, attrs, T=None, density=1.0):
        a, ap = _d(xyz)
        n_, np_ = _d(normals)
        at = np.ascontiguousarray(attrs, dtype=np.int32)
        t, tp = _cm(T)
        lib().ref_add_model_ean(self._h, ap, np_, at.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[0], tp, density)

    def align_ean(self, xyz, normals, attrs, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, overlap=1.0):
        a, ap = _d(xyz)
        n_, np_ = _d(normals)
        at = np.ascontiguousarray(attrs, dtype=np.int32)
        t, tp = _cm(T)
        out, pd, n = np.zeros(32), np.zeros(max(len(a), 1)), C.c_size_t()
        lib().ref_align_ean(self._h, ap, np_, at.ctypes.data_as(C.POINTER(C.c_int32)), a.shape[0], tp, density, max_iter,
                            min_mse, min_mse_diff, overlap, out.ctypes.data_as(C.POINTER(C.c_double)),
                            pd.ctypes.data_as(C.POINTER(C.c_double)), len(pd), C.byref(n))
        return _unpack(out, pd, n.value)

    def align(self, xyz, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, overlap=1.0):
        a, ap = _d(xyz)
        t, tp = _cm(T)
        out, pd, n = np.zeros(32), np.zeros(max(len(a), 1)), C.c_size_t()
        lib().ref_align(self._h, ap, a.shape[0], tp, density, max_iter, min_mse, min_mse_diff, overlap,
                        out.ctypes.data_as(C.POINTER(C.c_double)), pd.ctypes.data_as(C.POINTER(C.c_double)), len(pd),
                        C.byref(n))
        return _unpack(out, pd, n.value)

    def align_golden(self, xyz, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, alpha=0.4, beta=1.0,
                     eps=0.05):
        a, ap = _d(xyz)
        t, tp = _cm(T)
        out, pd, n = np.zeros(32), np.zeros(max(len(a), 1)), C.c_size_t()
        lib().ref_align_golden(self._h, ap, a.shape[0], tp, density, max_iter, min_mse, min_mse_diff, alpha, beta, eps,
                               out.ctypes.data_as(C.POINTER(C.c_double)), pd.ctypes.data_as(C.POINTER(C.c_double)),
                               len(pd), C.byref(n))
        return _unpack(out, pd, n.value)


def pairs_transform(x, p, d, n_po):
    """Pairs::add xN, trim(n_po), getTransform()."""
    x_, xp = _d(x)
    p_, pp = _d(p)
    d_, dp = _d(d)
    T, mse, tau = np.zeros(16), C.c_double(), C.c_double()
    kept = lib().ref_pairs(xp, pp, dp, len(d_), n_po, T.ctypes.data_as(C.POINTER(C.c_double)), C.byref(mse),
                           C.byref(tau))
    return kept, T.reshape(4, 4).T.copy(), mse.value, tau.value
