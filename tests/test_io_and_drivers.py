"""The rows SURVEY.md section 8f marks "next": PLY I/O (f3), the ICPLocalization caller (f1), and the env_icp /
env_slam6d-style drivers written against the kept API (cfg1 / cfg4 of BASELINE.json), each checked against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_transform_close

PK = os.path.join(ROOT, "slam-envire_b200")


def _build(src, exe, link=True):
    deps = [src, os.path.join(PK, "icp", "icp_gpu.hpp"), os.path.join(PK, "io", "scene.hpp"),
            os.path.join(PK, "io", "ply.hpp"), os.path.join(PK, "compat", "Eigen", "mini_eigen.hpp")]
    if link:
        deps.append(os.path.join(PK, "libicp_b200.so"))
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + PK, "-I" + os.path.join(PK, "compat"),
               "-I" + os.path.join(ROOT, "include"), "-o", exe, src]
        if link:
            cmd += ["-L" + PK, "-licp_b200", "-Wl,-rpath," + PK]
        subprocess.check_call(cmd)
    return exe


def _read_ply_numpy(path):
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    dt = "<f8" if "property double x" in head.decode() else "<f4"
    xyz = np.frombuffer(body, dtype=dt, count=3 * n).reshape(n, 3)
    rest = body[xyz.nbytes:]
    normals = np.frombuffer(rest, dtype=dt, count=3 * n).reshape(n, 3) if any(l.startswith("element normal") for l in lines) else None
    return xyz, normals


def test_ply_roundtrip(tmp_path, synth):
    exe = _build(os.path.join(ROOT, "tests", "cpp", "ply_roundtrip.cpp"), os.path.join(ROOT, "tests", "cpp", "ply_roundtrip"),
                 link=False)
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(1234, 3)) * 50
    nrm = rng.normal(size=(1234, 3))
    a, b = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    synth.write_ply(a, xyz, nrm)  # the reference writer's layout: vertex element, then a separate normal element
    out = subprocess.check_output([exe, a, b]).decode().split()
    assert out == ["1234", "1"]
    x2, n2 = _read_ply_numpy(b)
    assert np.array_equal(x2, xyz) and np.array_equal(n2, nrm)
    # float precision and no normals
    synth.write_ply(a, xyz, None, double=False)
    assert subprocess.check_output([exe, a, b, "0"]).decode().split() == ["1234", "0"]
    x3, n3 = _read_ply_numpy(b)
    assert n3 is None and np.array_equal(x3, xyz.astype(np.float32))
    # ascii with faces and colours to skip
    with open(a, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nelement face 1\nproperty list uchar int vertex_index\nend_header\n"
                "0 0 0 255\n1 0.5 0 0\n0 1 2.25 7\n3 0 1 2\n")
    assert subprocess.check_output([exe, a, b]).decode().split() == ["3", "0"]
    x4, _ = _read_ply_numpy(b)
    assert np.array_equal(x4, [[0, 0, 0], [1, 0.5, 0], [0, 1, 2.25]])


@pytest.mark.gpu
def test_env_icp_driver(tmp_path, oracle, synth):
    exe = _build(os.path.join(PK, "tools", "env_icp.cpp"), os.path.join(PK, "tools", "env_icp"))
    m, s, T, prm = synth.config("cfg1", 0.2)
    mp, sp, op = str(tmp_path / "m.ply"), str(tmp_path / "s.ply"), str(tmp_path / "o.ply")
    synth.write_ply(mp, m)
    synth.write_ply(sp, s)
    init = synth.rigid((0, 0, 1), 0.01, (0.02, 0.0, -0.01))
    args = [exe, mp, sp, "--overlap", str(prm["overlap"]), "--max-iter", "60", "--min-mse", "1e-10", "--min-mse-diff", "1e-12",
            "--out", op, "--init"] + ["%.17g" % v for v in init.reshape(16)]
    out = json.loads(subprocess.check_output(args).decode())
    M = oracle.Model()
    M.add(m)
    r = M.align(s, T=init, max_iter=60, min_mse=1e-10, min_mse_diff=1e-12, overlap=prm["overlap"])
    fn = oracle.apply_transform(init, init, r.matrix())
    assert out["iter"] == r.iter and out["pairs"] == r.pairs
    assert_transform_close(np.array(out["scan_frame"]), fn, 20.0)
    assert abs(out["mse"] - r.mse) <= 1e-5 * r.mse
    aligned, _ = _read_ply_numpy(op)
    assert np.allclose(aligned, s @ fn[:3, :3].T + fn[:3, 3], atol=1e-9)
    # golden-section mode runs and reports an overlap inside the bracket
    out = json.loads(subprocess.check_output([exe, mp, sp, "--golden", "0.4", "1.0", "0.05", "--max-iter", "10"]).decode())
    g, _ = M.align_golden(s, max_iter=10, min_mse=1e-5, min_mse_diff=1e-8, alpha=0.4, beta=1.0, eps=0.05)
    assert abs(out["overlap"] - g.overlap) < 1e-12 and out["iter"] == g.iter and out["pairs"] == g.pairs


@pytest.mark.gpu
def test_env_slam6d_driver_growing_model(tmp_path, synth):
    """cfg4 in miniature: 12 scans x 4000 points registered one after another into the growing model."""
    exe = _build(os.path.join(PK, "tools", "env_slam6d.cpp"), os.path.join(PK, "tools", "env_slam6d"))
    seq = synth.scan_sequence(12, 4000, seed=5)
    for i, sc in enumerate(seq):
        synth.write_ply(str(tmp_path / ("scan%03d.ply" % i)), sc["scan"])
        np.savetxt(str(tmp_path / ("scan%03d.pose" % i)), sc["pose_init"])
    out = json.loads(subprocess.check_output([exe, str(tmp_path), "12", "--max-iter", "40", "--overlap", "0.8"]).decode())
    assert out["scans"] == 12 and out["model_points"] == sum(len(s["scan"]) for s in seq)
    err_init, err_ref = [], []
    for i, sc in enumerate(seq):
        F = np.loadtxt(str(tmp_path / ("scan%03d.frames" % i)))
        err_ref.append(np.linalg.norm(F[:3, 3] - sc["pose_true"][:3, 3]))
        err_init.append(np.linalg.norm(sc["pose_init"][:3, 3] - sc["pose_true"][:3, 3]))
    # registration pulls the odometry-noisy poses towards the truth (drift accumulates, hence the loose bound)
    assert np.mean(err_ref[1:]) < 0.5 * np.mean(err_init[1:])
    assert max(err_ref) < 0.15


REF_LOC = os.path.join(ROOT, "oracle", "_ref", "ref_localization_cli")


def test_scene_reader_reads_the_reference_scene(tmp_path):
    """f3: the scene.yml the reference ships (test/scene.yml: two frame nodes, ids without the leading '/', transforms
    spread over four lines) read by io/scene.hpp; then a scene in the reference's emitter style with links and a PLY in
    its layout (src/tools/PlyFile.cpp:173-246); then writer -> reader."""
    exe = _build(os.path.join(ROOT, "tests", "cpp", "scene_cli.cpp"), os.path.join(ROOT, "tests", "cpp", "scene_cli"), link=False)
    ref_scene = "/root/reference/test"
    if os.path.exists(os.path.join(ref_scene, "scene.yml")):
        out = json.loads(subprocess.check_output([exe, ref_scene]).decode())
        assert "error" not in out, out
        assert [f["id"] for f in out["frames"]] == ["/1", "/2"] and out["clouds"] == []
        T1 = np.array(out["frames"][0]["transform"]).reshape(4, 4)
        assert np.array_equal(T1, np.array([[1, 0, 0, 0], [0, 1, 0, 2.5], [0, 0, 1, 0], [0, 0, 0, 1.0]]))
        assert np.array_equal(np.array(out["frames"][1]["transform"]).reshape(4, 4), np.eye(4))
    # a scene as FileSerialization writes it: objects (a root, a posed frame node below it, a cloud in that frame), links
    d = tmp_path / "scene"
    d.mkdir()
    T = np.eye(4)
    T[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]
    T[:3, 3] = [1.5, -2.0, 0.25]
    rows = ", \n               ".join(", ".join(repr(float(v)) for v in r) for r in T)
    (d / "scene.yml").write_text(
        "objects:\n - class: 'envire::FrameNode'\n   id: /0\n   label: ''\n   transform: [1, 0, 0, 0, \n               0, 1, 0, 0, \n"
        "               0, 0, 1, 0, \n               0, 0, 0, 1]\n"
        " - class: 'envire::FrameNode'\n   id: /1\n   label: ''\n   transform: [%s]\n"
        " - class: 'envire::Pointcloud'\n   id: /2\n   label: 'scan'\n   sensor_origin: [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]\n"
        "links:\n - type: frameNodeTree\n   child: /1\n   parent: /0\n - type: cartesianMapGraph\n   map: /2\n   node: /1\n" % rows)
    synth = __import__("importlib").import_module("slam-envire_b200.synth")
    pts = np.random.default_rng(5).normal(0, 3, (1234, 3))
    synth.write_ply(str(d / "envire::Pointcloud_2.ply"), pts, normals=np.tile([0.0, 0.0, 1.0], (1234, 1)))
    d2 = tmp_path / "rewritten"
    d2.mkdir()
    for args in ([str(d)], [str(d), str(d2)]):
        out = json.loads(subprocess.check_output([exe] + args).decode())
        assert "error" not in out, out
        assert out["root"] == "/0" and [f["id"] for f in out["frames"]] == ["/0", "/1"] and out["frames"][1]["parent"] == "/0"
        assert np.array_equal(np.array(out["frames"][1]["transform"]).reshape(4, 4), T)
        assert np.array_equal(np.array(out["frames"][1]["to_root"]).reshape(4, 4), T)
        (c,) = out["clouds"]
        assert c["id"] == "/2" and c["frame"] == "/1" and c["n"] == 1234 and c["normals"]
        assert np.allclose(c["sum"], pts.sum(0), rtol=0, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("cov_mode,density", [(0, 1.0), (1, 1.0), (2, 0.5)])
def test_reference_icp_localization_on_the_drop_in(tmp_path, oracle, cov_mode, density):
    """f1 / drop-in proof: the REFERENCE's own icp/icpLocalization.cpp + .hpp, compiled unmodified by oracle/Makefile with
    envire::icp::TrimmedKD = TrimmedGPU (oracle/_ref/ref_localization_cli), runs loadEnvironment (scene directory ->
    addToModel), scan-line accumulation and doScanMatch on the GPU; result and covariance heuristics
    (icp/icpLocalization.cpp:114-130,188-252) against the oracle run on the very same accumulated cloud."""
    exe = REF_LOC
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_localization_cli was not built (needs /root/reference: make -C oracle dropin)")
    (tmp_path / "scene").mkdir()
    rng = np.random.default_rng(1)
    model = np.concatenate([rng.uniform(-15, 15, (60000, 2)), np.zeros((60000, 1))], 1)  # flat ground, z = 0
    mp, cp = str(tmp_path / "m.bin"), str(tmp_path / "cloud.bin")
    model.tofile(mp)
    txt = subprocess.check_output([exe, mp, str(cov_mode), str(density), cp, str(tmp_path / "scene")]).decode()
    txt = txt[txt.index("{"):]  # the reference's loadEnvironment prints progress lines first
    out = json.loads(txt.replace(" inf", " Infinity"))  # printf prints inf for the NO_COV covariances
    assert "error" not in out, out
    cloud = np.fromfile(cp).reshape(-1, 3)
    assert out["has_new"] and out["points"] == len(cloud) and out["detached"]
    assert len(cloud) > 40 * 200  # 40 stored lines (lines_per_point_cloud), most rays hit the ground
    frm = np.array(out["from"])
    M = oracle.Model()
    M.add(model)
    run = oracle.Run()
    r = M.align(cloud, T=frm, density=density, max_iter=25, min_mse=1e-9, min_mse_diff=1e-11, overlap=0.9, run=run)
    to = oracle.apply_transform(frm, frm, r.matrix())
    assert out["pairs"] == r.pairs and out["n_pd"] == len(run.pairs_distance())
    assert_transform_close(np.array(out["to"]), to, 30.0)
    assert abs(out["mse"] - r.mse) <= 1e-5 * r.mse
    if cov_mode == 0:
        assert out["cov_pos00"] == 0.01 and out["cov_ori00"] == 0.02 and out["cov_pos01"] == 0.0
    elif cov_mode == 1:
        f = np.float32(np.float32(max(1.0, density) / 4.0) / np.sqrt(r.mse))
        assert abs(out["cov_pos00"] - 1e-3 * 2.0 / (float(f) * .5) ** 4) <= 1e-4 * out["cov_pos00"]
        assert abs(out["cov_ori00"] - (np.pi / 180) / (float(f) * .5) ** 4) <= 1e-4 * out["cov_ori00"]
    else:
        assert np.isinf(out["cov_pos00"]) and np.isinf(out["cov_pos01"])
