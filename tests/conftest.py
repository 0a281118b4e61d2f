import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has a hyphen, hence importlib)."""
    return importlib.import_module("slam-envire_b200")


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module("slam-envire_b200.synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def icp(pkg):
    """A live context on cuda:0.  GPU tests only; fails loudly when the extension or the device is missing."""
    ctx = pkg.Icp(0)
    yield ctx
    ctx.close()


class Harness:
    """tests/harness/cpu_harness.cpp: the product's __host__ __device__ search/pose code run on the CPU."""

    def __init__(self):
        src = os.path.join(ROOT, "tests", "harness", "cpu_harness.cpp")
        lib = os.path.join(ROOT, "tests", "harness", "libcpu_harness.so")
        deps = [src] + [os.path.join(ROOT, "slam-envire_b200", "csrc", f) for f in ("grid_view.h", "pose.h", "density.h")]
        if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", lib, src])
        self.L = C.CDLL(lib)

    @staticmethod
    def _p(a, t=C.c_double):
        return a.ctypes.data_as(C.POINTER(t))

    def search(self, model, h, q, d_box=np.inf, warm=None):
        model = np.ascontiguousarray(model, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        idx = np.zeros(len(q), np.uint32)
        d = np.zeros(len(q))
        lv = np.zeros(len(q), np.int32)
        w = None if warm is None else np.ascontiguousarray(warm, dtype=np.uint32)
        self.L.hs_search(self._p(model), C.c_size_t(len(model)), C.c_double(h), self._p(q), C.c_size_t(len(q)),
                         C.c_double(d_box), self._p(w, C.c_uint32) if w is not None else None,
                         self._p(idx, C.c_uint32), self._p(d), self._p(lv, C.c_int))
        return idx, d, lv

    def density_indices(self, n, density):
        """(visited vertex indices, number of segments) from the closed form of the density stepping (csrc/density.h)."""
        out = np.zeros(int(n * max(density, 1.0)) + 16, np.uint64)
        ns = C.c_int()
        self.L.hs_density_indices.restype = C.c_longlong
        m = self.L.hs_density_indices(C.c_size_t(n), C.c_double(density), self._p(out, C.c_uint64), C.c_size_t(len(out)), C.byref(ns))
        return (out[:m].astype(np.int64) if m >= 0 else None), ns.value

    def transform_from_sums(self, sums):
        s = np.ascontiguousarray(sums, dtype=np.float64)
        T = np.zeros(12)
        mse = C.c_double()
        self.L.hs_transform_from_sums(self._p(s), self._p(T), C.byref(mse))
        M = np.eye(4)
        M[:3, :] = T.reshape(3, 4)
        return M, mse.value

    def skip_sequence(self, model, h, q_steps, d_box, reach_cells=0.25, floor_cells=0.05):
        """q_steps: (steps, nq, 3) positions of the same queries over time; d_box: one per step."""
        model = np.ascontiguousarray(model, dtype=np.float64)
        q = np.ascontiguousarray(q_steps, dtype=np.float64)
        steps, nq = q.shape[0], q.shape[1]
        db = np.ascontiguousarray(np.broadcast_to(np.asarray(d_box, float), (steps,)))
        idx = np.zeros((steps, nq), np.uint32)
        d = np.zeros((steps, nq))
        sk = np.zeros((steps, nq), np.uint8)
        self.L.hs_skip_sequence(self._p(model), C.c_size_t(len(model)), C.c_double(h), self._p(q), C.c_size_t(nq),
                                C.c_int(steps), self._p(db), C.c_double(reach_cells), C.c_double(floor_cells),
                                self._p(idx, C.c_uint32), self._p(d), self._p(sk, C.c_uint8))
        return idx, d, sk.astype(bool)

    def align(self, model, h, src, T=None, density=1.0, max_iter=10, min_mse=1e-4, min_mse_diff=1e-5, overlap=1.0,
              warm=1, skip=1, reach_cells=0.25, floor_cells=0.05, lazy_margin=0.05):
        model = np.ascontiguousarray(model, dtype=np.float64)
        src = np.ascontiguousarray(src, dtype=np.float64)
        res = np.zeros(32)
        trace = np.zeros((max(max_iter, 1), 28))
        Tcm = np.ascontiguousarray(np.asarray(np.eye(4) if T is None else T, float).T.reshape(16))
        ex = self.L.hs_align(self._p(model), C.c_size_t(len(model)), C.c_double(h), self._p(src),
                             C.c_size_t(len(src)), C.c_size_t(int(len(src) * density)), self._p(Tcm),
                             C.c_size_t(max_iter), C.c_double(min_mse), C.c_double(min_mse_diff), C.c_double(overlap),
                             C.c_int(warm), self._p(res), self._p(trace), C.c_size_t(len(trace)), C.c_int(skip), C.c_double(reach_cells),
                             C.c_double(floor_cells), C.c_double(lazy_margin))
        return dict(C=res[:16].reshape(4, 4).T.copy(), iter=int(res[16]), pairs=int(res[17]), mse=res[18],
                    mse_diff=res[19], d_box=res[20], early=bool(res[21]), trace=trace[:ex])


@pytest.fixture(scope="session")
def harness():
    return Harness()


def rot_angle(R):
    return float(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)))


def assert_transform_close(C_got, C_ref, extent, rot_tol=1e-5, trans_tol_rel=1e-6):
    """north_star tolerance: rotation within 1e-5 rad, translation within 1e-6 of the scene extent."""
    E = C_got[:3, :3] @ C_ref[:3, :3].T
    assert rot_angle(E) <= rot_tol, "rotation differs by %g rad" % rot_angle(E)
    dt = np.linalg.norm(C_got[:3, 3] - C_ref[:3, 3])
    assert dt <= trans_tol_rel * extent, "translation differs by %g (extent %g)" % (dt, extent)
