"""Inputs of the golden cases (shared by make_golden.py and the tests): all seeded / closed-form, so only the
OUTPUTS of the reference need to be stored in golden.json."""
import importlib

import numpy as np

synth = importlib.import_module("slam-envire_b200.synth")


def sine_wave():
    """test/unit/icp.cpp:52-91 (setSineWave): 21x21 lattice, z = 0.2 cos(10 r)."""
    pts = []
    for i in range(-10, 11):
        for j in range(-10, 11):
            x, y = i * .1, j * .1
            pts.append((x, y, .2 * np.cos(10.0 * np.sqrt(x * x + y * y))))
    return np.array(pts)


def sine_wave_ean():
    """setSineWave incl. its normals and scan-edge attributes (test/unit/icp.cpp:52-91)."""
    pts, nrm, att = [], [], []
    for i in range(-10, 11):
        for j in range(-10, 11):
            edge = i in (-10, 10) or j in (-10, 10)
            x, y = i * .1, j * .1
            d = np.sqrt(x * x + y * y)
            z = .2 * np.cos(10.0 * d)
            zd = -.2 * 10.0 * np.sin(10.0 * d)
            nx = np.sqrt(1.0 / (1.0 + zd * zd))
            ny = zd * nx
            a = np.arctan2(y, x)
            Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
            pts.append((x, y, z))
            nrm.append(Rz @ np.array([-ny, 0.0, nx]))
            att.append(int(edge) << 1)  # edge << TriMesh::SCAN_EDGE
    return np.array(pts), np.array(nrm), np.array(att, dtype=np.int32)


def half_cube():
    """test/unit/icp.cpp:23-50 (setHalfCube): 3 adjacent walls of a unit cube, 363 points."""
    pts = []
    for i in range(11):
        for j in range(11):
            x, y = i * .1, j * .1
            pts += [(x, y, 0.0), (0.0, x, y), (x, 0.0, y)]
    return np.array(pts)


def test_transform():
    """Translation3d(0,0,5) * AngleAxisd(0.1, UnitX)  (test/unit/icp.cpp:150-152)."""
    c, s = np.cos(0.1), np.sin(0.1)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 5.0], [0, 0, 0, 1.0]])


def small_transform():
    return synth.rigid((0.2, 1.0, 0.1), 0.05, (0.03, -0.02, 0.04))


def cases():
    """name -> dict(model=[(xyz, T, density)...], src, T_src, density, kind, params)"""
    out = {}
    sw, hc = sine_wave(), half_cube()
    gui = dict(max_iter=10, min_mse=1e-4, min_mse_diff=1e-5)
    out["icp_test2_sine_golden_section"] = dict(model=[(sw, None, 1.0)], src=sw, T_src=test_transform(), density=1.0,
                                                kind="golden", params=dict(gui, alpha=0.4, beta=1.0, eps=0.05))
    for ov in (1.0, 0.7, 0.4):
        out["sine_overlap_%.1f" % ov] = dict(model=[(sw, None, 1.0)], src=sw, T_src=test_transform(), density=1.0,
                                             kind="fixed", params=dict(gui, overlap=ov))
    out["halfcube_small_motion"] = dict(model=[(hc, None, 1.0)], src=hc, T_src=small_transform(), density=1.0,
                                        kind="fixed", params=dict(max_iter=40, min_mse=1e-12, min_mse_diff=1e-14, overlap=0.9))
    out["halfcube_golden_section"] = dict(model=[(hc, None, 1.0)], src=hc, T_src=small_transform(), density=1.0,
                                          kind="golden", params=dict(gui, alpha=0.4, beta=1.0, eps=0.05))
    # TrimmedKDEAN: edge/normal pair filter (icp/icp.hpp:206-221, 513-521)
    p_e, n_e, a_e = sine_wave_ean()
    for ov in (0.8, 0.5):
        out["ean_sine_overlap_%.1f" % ov] = dict(model=[(p_e, None, 1.0)], src=p_e, T_src=small_transform(), density=1.0,
                                                 kind="ean", normals=n_e, attrs=a_e,
                                                 params=dict(max_iter=10, min_mse=1e-9, min_mse_diff=1e-12, overlap=ov))
    m, s, T, prm = synth.config("cfg1", 0.1)
    # a scene with analytic normals: ground normals from the height field, wall normals axis-aligned; 3 % edge flags
    rng = np.random.default_rng(9)
    nrm = np.zeros_like(m)
    g = np.abs(m[:, 2] - 0.3 * np.sin(0.5 * m[:, 0]) * np.cos(0.4 * m[:, 1])) < 1e-12
    gx = 0.15 * np.cos(0.5 * m[:, 0]) * np.cos(0.4 * m[:, 1])
    gy = -0.12 * np.sin(0.5 * m[:, 0]) * np.sin(0.4 * m[:, 1])
    nrm[g] = np.stack([-gx[g], -gy[g], np.ones(g.sum())], 1)
    nrm[~g & (m[:, 1] == 10.0)] = (0, -1, 0)
    nrm[~g & (m[:, 1] != 10.0)] = (1, 0, 0)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    att = (rng.random(len(m)) < 0.03).astype(np.int32) << 1
    Rt = T[:3, :3]
    out["ean_cfg1_scale0.1"] = dict(model=[(m, None, 1.0)], src=s, T_src=None, density=1.0, kind="ean", normals=nrm,
                                    attrs=att, src_normals=nrm @ Rt, params=dict(prm, max_iter=25))
    out["cfg1_scale0.1"] = dict(model=[(m, None, 1.0)], src=s, T_src=None, density=1.0, kind="fixed", params=prm)
    out["cfg1_gui_defaults"] = dict(model=[(m, None, 1.0)], src=s, T_src=None, density=1.0, kind="fixed",
                                    params=dict(max_iter=5, min_mse=1e-5, min_mse_diff=1e-8, overlap=0.8))
    T_src = synth.rigid((1, 0, 0), 0.02, (0.1, -0.2, 0.05))
    out["cfg1_density_and_frames"] = dict(model=[(m[:3000], None, 1.0), (m[3000:], synth.rigid((0, 0, 1), 0.0, (0, 0, 0)), 0.5)],
                                          src=(s - T_src[:3, 3]) @ T_src[:3, :3], T_src=T_src, density=0.5, kind="fixed",
                                          params=dict(prm, max_iter=20))
    m2, s2, T2, prm2 = synth.config("cfg2", 0.01)
    out["cfg2_scale0.01"] = dict(model=[(m2, None, 1.0)], src=s2, T_src=None, density=1.0, kind="fixed",
                                 params=dict(prm2, max_iter=40))
    m3 = synth.scene(6000, 60.0, 4)
    s3, _ = synth.perturbed_copy(m3, 1.0, 0.3, 0.01, seed=5, replace_frac=0.5, extent=60.0)
    out["cfg3_style_outliers"] = dict(model=[(m3, None, 1.0)], src=s3, T_src=None, density=1.0, kind="fixed",
                                      params=dict(max_iter=15, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.5))
    return out
