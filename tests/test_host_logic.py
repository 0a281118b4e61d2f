"""CPU tier: the product's own host/device-shared code (no GPU needed).

  * tests/harness runs grid_view.h (icpb_nn_search) and pose.h (state machine, closed-form pose) on the CPU
  * the C-ABI library loads and exports every symbol include/icp_b200.h declares
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_transform_close


@pytest.mark.parametrize("nm,nq,h,d_box,seed", [
    (5000, 2000, 0.7, np.inf, 1), (5000, 2000, 0.25, np.inf, 2), (5000, 2000, 4.0, np.inf, 3),
    (5000, 2000, 0.5, 0.4, 4), (5000, 2000, 0.5, 1.7, 5), (1, 50, 1.0, np.inf, 6), (9, 50, 1.0, 0.5, 7),
    # array tails: the scanners read four candidates at a time and rely on the sentinels behind the last point
    (2, 64, 1.0, np.inf, 8), (3, 64, 0.3, 0.9, 9), (4999, 1500, 0.6, np.inf, 10), (5001, 1500, 0.6, 0.8, 11),
    (4998, 1500, 40.0, np.inf, 12),  # one cell holds everything
])
def test_grid_search_exact(harness, oracle, nm, nq, h, d_box, seed):
    for bl in (1, 0, -1):  # flat block scan for start levels <= 1 / 0 only / never (fast path + pyramid walk)
        harness.L.hs_set_block_levels(bl)
        _grid_search_exact(harness, oracle, nm, nq, h, d_box, seed)
    harness.L.hs_set_block_levels(1)


def _grid_search_exact(harness, oracle, nm, nq, h, d_box, seed):
    rng = np.random.default_rng(seed)
    m = rng.uniform(-10, 10, (nm, 3))
    q = rng.uniform(-14, 14, (nq, 3))
    M = oracle.Model()
    M.add(m)
    i0, d0 = M.find_pairs(q, d_box=d_box, brute=True)
    d0 = np.where(i0 == 0xFFFFFFFF, d_box, d0)
    i1, d1, lv = harness.search(m, h, q, d_box)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    # a warm-start candidate (previous iteration's pair) never changes the answer
    w = rng.integers(0, nm, nq).astype(np.uint32)
    i2, d2, _ = harness.search(m, h, q, d_box, warm=w)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_grid_search_ties_lowest_index(harness, oracle):
    g = np.arange(0, 8, dtype=np.float64)
    m = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    m = np.concatenate([m, m[:64]])  # exact duplicates
    q = np.concatenate([m[:300] + 0.5, m[:50]])
    M = oracle.Model()
    M.add(m)
    i0, d0 = M.find_pairs(q, brute=True)
    for h in (0.5, 1.0, 3.0):
        i1, d1, _ = harness.search(m, h, q)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_grid_search_deep_pyramid(harness, oracle):
    """An elongated cloud in small cells: 2000 cells along x, 11 pyramid levels -- the walk's trail uses its second
    register (levels 9 and up), cold searches start at the top level."""
    rng = np.random.default_rng(21)
    m = np.stack([rng.uniform(0, 20, 3000), rng.uniform(0, 0.05, 3000), rng.uniform(0, 0.05, 3000)], 1)
    q = np.concatenate([m[:400] + rng.normal(0, 0.3, (400, 3)), rng.uniform(-3, 23, (300, 3)) * [1, 0.1, 0.1]])
    M = oracle.Model()
    M.add(m)
    for d_box in (np.inf, 2.5, 0.2):
        i0, d0 = M.find_pairs(q, d_box=d_box, brute=True)
        d0 = np.where(i0 == 0xFFFFFFFF, d_box, d0)
        i1, d1, lv = harness.search(m, 0.01, q, d_box)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
        if d_box == np.inf:
            assert lv.max() == 11
        # 8000 cells along x: 13 levels; walks start at level 12 at most (two root nodes along x)
        i2, d2, lv2 = harness.search(m, 0.0025, q, d_box)
        assert np.array_equal(i0, i2) and np.array_equal(d0, d2)
        if d_box == np.inf:
            assert lv2.max() == 12


def test_grid_search_surface_scene(harness, oracle, synth):
    m, s, T, prm = synth.config("cfg1", 0.2)
    M = oracle.Model()
    M.add(m)
    for d_box in (np.inf, 0.8):
        i0, d0 = M.find_pairs(s, d_box=d_box)
        d0 = np.where(i0 == 0xFFFFFFFF, d_box, d0)
        i1, d1, lv = harness.search(m, 0.45, s, d_box)
        assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_pose_step_matches_oracle(harness, oracle):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(400, 3)) * [5, 3, 1] + [10, -4, 2]
    p = x @ np.array([[0.98, -0.17, 0.05], [0.17, 0.98, 0.02], [-0.05, -0.01, 1.0]]) + [0.3, -0.2, 0.1]
    d = np.linalg.norm(p - x, axis=1)
    T_ref, mse_ref = oracle.get_transform(p, x, d, np.arange(400))
    sums = np.concatenate([p.sum(0), x.sum(0), (p[:, :, None] * x[:, None, :]).sum(0).reshape(9), [(d * d).sum(), 400]])
    T, mse = harness.transform_from_sums(sums)
    assert np.allclose(T, T_ref, rtol=0, atol=1e-12) and abs(mse - mse_ref) <= 1e-14 * mse_ref


@pytest.mark.parametrize("h,d_box,step,reach,floor", [
    (0.5, np.inf, 0.01, 0.25, 0.05), (0.5, 0.6, 0.02, 1.0, 0.2), (0.25, 0.3, 0.004, 0.0, 0.0), (1.5, 0.2, 0.03, 8.0, 0.4),
    (0.5, 0.05, 0.001, 0.25, 0.05),
])
def test_search_skipping_is_exact(harness, oracle, synth, h, d_box, step, reach, floor):
    """Search skipping (icpb_nn_skip, grid_view.h): queries drift a little from step to step, as the source of an align
    does; every step's answers -- searched or skipped -- must be the exact nearest neighbours the oracle finds, with
    bit-identical distances.  The model is a surface scene plus exact duplicates and a lattice patch (ties)."""
    rng = np.random.default_rng(17)
    m = synth.scene(6000, 12.0, 5)
    g = np.arange(0, 6, dtype=np.float64) * 0.25
    lat = np.stack(np.meshgrid(g, g, [0.0], indexing="ij"), -1).reshape(-1, 3) + [1.0, 1.0, 2.0]
    m = np.concatenate([m, m[:200], lat])
    q0 = np.concatenate([m[rng.choice(len(m), 1500, replace=False)] + rng.normal(0, 0.02, (1500, 3)),
                         lat[:20] + [0.125, 0.125, 0.0],                       # equidistant from four lattice points
                         rng.uniform(-8, 8, (200, 3))])
    steps = 12
    drift = rng.normal(0, 1, 3)
    drift *= step / np.linalg.norm(drift)
    q = np.stack([q0 + t * drift + (rng.normal(0, step * 0.1, q0.shape) if t else 0) for t in range(steps)])
    dbs = np.full(steps, d_box) * np.linspace(1.0, 0.7, steps)                 # d_box shrinks like 2*tau does
    idx, d, sk = harness.skip_sequence(m, h, q, dbs, reach, floor)
    M = oracle.Model()
    M.add(m)
    for t in range(steps):
        i0, d0 = M.find_pairs(q[t], d_box=dbs[t], brute=True)
        d0 = np.where(i0 == 0xFFFFFFFF, dbs[t], d0)
        assert np.array_equal(i0, idx[t]) and np.array_equal(d0, d[t]), t
    assert not sk[0].any()
    if floor > 0 and 1.5 * step <= 0.25 * floor * h:  # slow enough for the searches to widen their reach
        assert sk[1:].mean() > 0.1  # the rule does fire


@pytest.mark.parametrize("warm,skip,lazy", [(0, 0, -1), (1, 0, -1), (1, 1, -1), (1, 1, 0.05), (1, 1, 0.0), (1, 1, -0.3)])
def test_emulated_device_loop_matches_oracle(harness, oracle, synth, warm, skip, lazy):
    """lazy > -1: lazy search with that margin.  -0.3 puts the guard BELOW the previous tau: most trims fail to verify and
    their loop bodies run twice -- the results must not change."""
    m, s, T, prm = synth.config("cfg1", 0.1)
    M = oracle.Model()
    M.add(m)
    run = oracle.Run()
    r = M.align(s, run=run, **prm)
    e = harness.align(m, 0.5, s, warm=warm, skip=skip, lazy_margin=lazy, **prm)
    if skip:
        assert e["trace"][:, 24].sum() > 0 and e["trace"][0, 24] == 0
    if lazy >= 0:
        assert e["trace"][:, 25].sum() > 0  # queries were deferred
    if lazy == -0.3:
        assert e["trace"][-1, 26] > 0  # loop bodies were run again
    assert e["iter"] == r.iter and e["pairs"] == r.pairs and not e["early"]
    assert_transform_close(e["C"], r.matrix(), 20.0)
    assert abs(e["mse"] - r.mse) <= 1e-5 * r.mse
    tr = run.trace()
    assert len(tr) == len(e["trace"])
    for a, b in zip(tr, e["trace"]):
        assert a["found"] == b[1] and a["kept"] == b[2] and abs(a["tau"] - b[3]) <= 1e-9


def test_emulated_device_loop_edge_cases(harness, oracle, synth):
    m = synth.scene(500, 10.0, 1)
    M = oracle.Model()
    M.add(m)
    for src, prm in [(m[:2], dict(max_iter=5)), (m[:50], dict(max_iter=5, overlap=0.001)), (m[:50], dict(max_iter=0)),
                     (m, dict(max_iter=5))]:
        r = M.align(src, **prm)
        e = harness.align(m, 0.6, src, **prm)
        assert e["iter"] == r.iter and e["pairs"] == r.pairs
        assert (np.isnan(e["d_box"]) and np.isnan(r.d_box)) or e["d_box"] == r.d_box
        assert e["early"] == (r.pairs < 3 and prm["max_iter"] > 0)


def test_density_stepping_closed_form(harness):
    """K0: PointcloudAdapter visits idx = size_t(index), index += 1/density with the rounding of every fp64 addition
    (icp/icp.hpp:126-138).  csrc/density.h restates the accumulation as a few linear segments in exact integer
    arithmetic so that the device can gather all visits in parallel: every visited index must equal the loop's."""
    def loop(n, density):
        out, index, inc = [], 0.0, 1.0 / density
        while int(index) < n:
            out.append(int(index))
            index += inc
        return np.asarray(out, dtype=np.int64)
    rng = np.random.default_rng(3)
    cases = [(1000, 0.37), (17, 0.5), (5, 1.0), (100000, 0.3), (100000, 1.0 / 3.0), (300000, 0.999999), (250000, 0.7),
             (1 << 18, 1.0 - 2.0 ** -20), (200000, 0.4), (200000, 2.0 / 3.0), (50000, 0.05), (3, 0.01), (40000, 1.7), (1, 0.5),
             (0, 0.5), (2000000, 0.61803398875)]
    cases += [(int(rng.integers(1, 400000)), float(rng.uniform(0.02, 1.0))) for _ in range(24)]
    # step exactly half-way between grid points of a binade: 1/density = 1 + 2^-40 ... -> ties in the binade with ulp 2^-39
    cases += [(1 << 16, 1.0 / (1.0 + 2.0 ** -40)), (1 << 16, 1.0 / (1.5 + 2.0 ** -45)), (100000, 1.0 / (2.0 + 2.0 ** -49))]
    for n, density in cases:
        got, nseg = harness.density_indices(n, density)
        want = loop(n, density)
        assert got is not None and nseg <= 200, (n, density, nseg)
        assert len(got) == len(want) and np.array_equal(got, want), (n, density)
    got, nseg = harness.density_indices(1000, 900.0)  # absurd densities are left to the host loop
    assert got is None and nseg == -1


def test_c_abi_exports_every_declared_symbol(pkg):
    """include/icp_b200.h is the boundary: every prototype in it must be exported by libicp_b200.so."""
    hdr = open(os.path.join(ROOT, "include", "icp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(icpb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pkg.SYMBOLS)
    L = pkg.load_library()
    for name in declared:
        assert hasattr(L, name), name
    assert L.icpb_version().startswith(b"icp_b200")


def test_create_multi_rejects_bad_device_lists(pkg):
    """icpb_create_multi: argument errors are reported before any device is touched (runs without a GPU)."""
    import ctypes as C
    L = pkg.load_library()
    L.icpb_create_multi.restype = C.c_int
    h = C.c_void_p()
    for ids in ([0, 0], [1, 2, 1], []):
        arr = (C.c_int * max(len(ids), 1))(*ids)
        assert L.icpb_create_multi(C.byref(h), arr, len(ids)) == pkg.ERR_ARG
        assert not h.value
    arr = (C.c_int * 9)(*range(9))
    assert L.icpb_create_multi(C.byref(h), arr, 9) == pkg.ERR_ARG  # more than 8 ranks: the exchange rows are sized for 8
    assert L.icpb_create_multi(None, arr, 2) == pkg.ERR_ARG


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product must fail loudly, not fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.IcpError) as ei:
        pkg.Icp(0)
    assert ei.value.status == pkg.ERR_CUDA


def test_product_does_not_touch_oracle():
    """oracle/ is test infrastructure: nothing under slam-envire_b200/ or include/ may reference it."""
    for base in ("slam-envire_b200", "include"):
        for dp, dn, fn in os.walk(os.path.join(ROOT, base)):
            for f in fn:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "icp_oracle" not in txt and "oracle." not in txt.replace("the oracle.", ""), f
