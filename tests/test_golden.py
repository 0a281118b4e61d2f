"""CPU tier: the C oracle against golden vectors produced by the REFERENCE's own icp.cpp/icp.hpp (oracle/_ref, see
tests/golden/make_golden.py), and -- where _ref is present (dev container) -- directly against it."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_transform_close

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases as golden_cases  # noqa: E402

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["cases"]
CASES = golden_cases.cases()


def check_against_golden(name, C, it, pairs, mse, overlap, pd, oracle, rot_tol=1e-5):
    g, c = GOLDEN[name], CASES[name]
    T_src = np.eye(4) if c["T_src"] is None else c["T_src"]
    fn_T = oracle.apply_transform(T_src, T_src, C)  # PointcloudAdapter::applyTransform
    extent = float(np.ptp(c["model"][0][0], axis=0).max())
    assert it == g["iter"] and pairs == g["pairs"]
    assert abs(overlap - g["overlap"]) < 1e-12
    assert_transform_close(fn_T, np.array(g["fn_T"]), extent, rot_tol=rot_tol)
    assert abs(mse - g["mse"]) <= 1e-5 * abs(g["mse"]) + 1e-300
    assert len(pd) == g["n_pairs_distance"]
    if len(pd):
        assert abs(np.sum(pd) - g["pd_sum"]) <= 1e-7 * abs(g["pd_sum"]) + 1e-12
        assert np.allclose(pd[-4:], g["pd_tail"], rtol=1e-6, atol=1e-10)
        assert np.all(np.diff(pd) >= 0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(oracle, name):
    c = CASES[name]
    M = oracle.Model()
    run = oracle.Run()
    if c["kind"] == "ean":
        for xyz, T, dens in c["model"]:
            M.add_ean(xyz, c["normals"], c["attrs"], T=T, density=dens)
        r = M.align_ean(c["src"], c.get("src_normals", c["normals"]), c["attrs"], T=c["T_src"], density=c["density"],
                        run=run, **c["params"])
        check_against_golden(name, r.matrix(), r.iter, r.pairs, r.mse, r.overlap, run.pairs_distance(), oracle)
        return
    for xyz, T, dens in c["model"]:
        M.add(xyz, T=T, density=dens)
    if c["kind"] == "golden":
        r, _ = M.align_golden(c["src"], T=c["T_src"], density=c["density"], run=run, **c["params"])
    else:
        r = M.align(c["src"], T=c["T_src"], density=c["density"], run=run, **c["params"])
    check_against_golden(name, r.matrix(), r.iter, r.pairs, r.mse, r.overlap, run.pairs_distance(), oracle)


def test_golden_file_is_current():
    assert set(GOLDEN) == set(CASES)


# ---- direct comparisons with the compiled reference (dev container only) ----
def _ref():
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    return R


def test_ref_pairs_trim_and_transform(oracle):
    """Pairs::add/trim/getTransform of the reference vs orc_trim/orc_get_transform, incl. the non-Horn pose solve."""
    R = _ref()
    rng = np.random.default_rng(0)
    for trial in range(20):
        n = int(rng.integers(3, 400))
        x = rng.normal(size=(n, 3)) * rng.uniform(0.1, 30) + rng.normal(size=3) * 10
        ang = rng.uniform(-1.5, 1.5)
        A = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        p = x @ A + rng.normal(size=3) + 0.01 * rng.normal(size=(n, 3))
        d = np.linalg.norm(p - x, axis=1)
        n_po = int(rng.integers(3, n + 5))
        kept, T_ref, mse_ref, tau_ref = R.pairs_transform(x, p, d, n_po)
        order, tau = oracle.trim(d, n_po)
        T, mse = oracle.get_transform(p, x, d, order)
        assert kept == len(order) and tau == tau_ref
        assert np.allclose(T, T_ref, rtol=0, atol=1e-11 * (1 + np.abs(x).max()))
        assert abs(mse - mse_ref) <= 1e-12 * mse_ref
    kept, _, _, tau = R.pairs_transform(x[:2], p[:2], d[:2], 5)
    assert kept == -1  # std::runtime_error("not enough pairs to get transform")


def test_ref_align_random_clouds(oracle, synth):
    R = _ref()
    rng = np.random.default_rng(5)
    for trial in range(4):
        m = synth.scene(3000, 30.0, 100 + trial)
        s, _ = synth.perturbed_copy(m, float(rng.uniform(1, 8)), float(rng.uniform(0, 0.5)), 0.005, seed=trial)
        prm = dict(max_iter=int(rng.integers(3, 30)), min_mse=1e-9, min_mse_diff=1e-10, overlap=float(rng.uniform(0.4, 1.0)))
        dens = float(rng.choice([1.0, 0.5, 0.37]))
        r = R.Ref()
        r.add_to_model(m)
        g = r.align(s, density=dens, **prm)
        M = oracle.Model()
        M.add(m)
        run = oracle.Run()
        o = M.align(s, density=dens, run=run, **prm)
        assert o.iter == g["iter"] and o.pairs == g["pairs"]
        assert abs(o.mse - g["mse"]) <= 1e-9 * g["mse"]
        fn = oracle.apply_transform(np.eye(4), np.eye(4), o.matrix())
        assert np.allclose(fn, g["fn_T"], rtol=0, atol=1e-10)
        assert np.allclose(run.pairs_distance(), g["pairs_distance"], rtol=1e-9, atol=1e-13)


def test_oracle_reproduces_device_trace_at_full_size(oracle):
    """BASELINE configs[1] at FULL size (1M vs 1M): tests/golden/cfg2_poses.npz holds, for every loop body of the align
    as the B200 path ran it (tools/make_cfg2_poses.py), the pose it started from and its found / kept / tau / mse.  The
    oracle evaluates two of those loop bodies on the same input and must reproduce the recorded numbers: found and kept
    exactly, tau bit for bit (the same distance), mse to 1e-9 (the 17 sums are reduced in a different order)."""
    import os
    import bench
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cfg2_poses.npz"))
    model, src, _ = bench.workload(0)
    ref = bench.CpuReference(model)
    n_po = int(len(src) * bench.ALIGN["overlap"])
    for k in (len(z["tau"]) // 2, len(z["tau"]) - 1):
        r = ref.iteration(src, z["C_in"][k], float(z["d_box_in"][k]), n_po, 0)
        assert r["found"] == z["found"][k] and r["kept"] == z["kept"][k]
        assert r["tau"] == z["tau"][k]
        assert abs(r["mse"] - z["mse"][k]) <= 1e-9 * z["mse"][k]
