"""cfg4-style sequence: how many queries per loop body are skipped / deferred / far?  usage: python tools/studies/cfg4_skip.py [scans] [pts]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 12
pts = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
seq = synth.scan_sequence(n_scans, pts, seed=5, extent=120.0, radius=15.0, procedural=True)
icp = pkg.Icp(0)
icp.set_option("growth_margin", 1.0)
icp.set_option("profile", 1)
T0 = seq[0]["pose_true"]
icp.add_to_model(seq[0]["scan"], T=T0)
for i in range(1, n_scans):
    sc = seq[i]
    r = icp.align(sc["scan"], T=sc["pose_init"], max_iter=30, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.8)
    tr, tm = icp.trace(), icp.iteration_times()
    T = r.matrix() @ sc["pose_init"]
    print("scan %d: iter %d pairs %d mse %.3g | search us per body: %s" % (i, r.iter, r.pairs, r.mse, " ".join("%.0f" % (1e3 * t[0]) for t in tm[::3])))
    print("        skipped: %s" % " ".join("%d" % (t["skipped"] // 1000) for t in tr[::3]), "| deferred:", " ".join("%d" % (t["deferred"] // 1000) for t in tr[::3]),
          "| far:", " ".join("%d" % (t["far"] // 1000) for t in tr[::3]))
    icp.add_to_model(sc["scan"], T=T)
