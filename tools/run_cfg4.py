"""cfg4 (env_slam6d-style) at reduced length: N scans x P points registered sequentially into a growing model through the
C++ driver slam-envire_b200/tools/env_slam6d (TrimmedGPU).  usage: python tools/run_cfg4.py [n_scans] [pts_per_scan]
[procedural]   (procedural: every scan samples the scene afresh, O(points) per scan -- for the full 1000 scans)"""
import importlib, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
synth = importlib.import_module("slam-envire_b200.synth")
n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
procedural = len(sys.argv) > 3 and sys.argv[3] == "procedural"
PK = os.path.join(ROOT, "slam-envire_b200")
exe = os.path.join(PK, "tools", "env_slam6d")
subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + PK, "-I" + os.path.join(PK, "compat"),
                       "-I" + os.path.join(ROOT, "include"), "-o", exe, os.path.join(PK, "tools", "env_slam6d.cpp"),
                       "-L" + PK, "-licp_b200", "-Wl,-rpath," + PK])
d = tempfile.mkdtemp()
t0 = time.time()
seq = synth.scan_sequence(n_scans, pts, seed=5, extent=max(120.0, 1.5 * n_scans + 40.0), radius=15.0, procedural=procedural)
for i, sc in enumerate(seq):
    synth.write_ply(d + "/scan%03d.ply" % i, sc["scan"])
    np.savetxt(d + "/scan%03d.pose" % i, sc["pose_init"])
gen = time.time() - t0
out = {}
for margin in ("0", "1.0"):
    t0 = time.time()
    r = json.loads(subprocess.check_output([exe, d, str(n_scans), "--max-iter", "30", "--growth-margin", margin]).decode())
    r["wall_s"] = time.time() - t0
    err = [np.linalg.norm(np.loadtxt(d + "/scan%03d.frames" % i)[:3, 3] - sc["pose_true"][:3, 3]) for i, sc in enumerate(seq)]
    r["mean_pos_err_m"], r["max_pos_err_m"] = float(np.mean(err)), float(np.max(err))
    out["growth_margin_" + margin] = r
out["init_mean_pos_err_m"] = float(np.mean([np.linalg.norm(sc["pose_init"][:3, 3] - sc["pose_true"][:3, 3]) for sc in seq]))
out["generate_s"] = gen
print(json.dumps(out))
