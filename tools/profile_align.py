"""Per-iteration picture of one align on the GPU (trace + device times).  Usage: python tools/profile_align.py [cfg] [scale] [key=value ...]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
opts = dict(a.split("=") for a in sys.argv[3:])
m, s, T, prm = synth.config(cfg, scale)
for k in list(opts):
    if k in prm:
        prm[k] = type(prm[k])(float(opts.pop(k)))
icp = pkg.Icp(0)
icp.set_option("profile", 1)  # per-phase CUDA events (the default is the single-graph-launch loop)
for k, v in opts.items():
    icp.set_option(k, float(v))
icp.add_to_model(m)
gi = icp.build()
print("grid h=%.4f dims=%s levels=%d occupied=%d build=%.2f ms" % (gi.cell_size, list(gi.dims), gi.levels, gi.n_occupied, gi.build_ms))
icp.upload_source(s)
for rep in range(2):
    r = icp.align_resident(**prm)
t = icp.timing()
print("align: iter=%d pairs=%d mse=%.6g total=%.3f ms (search %.3f trim %.3f solve %.3f) launches=%d" % (
    r.iter, r.pairs, r.mse, t["total_ms"], t["search_ms"], t["trim_ms"], t["solve_ms"], t["kernels_launched"]))
E = r.matrix() @ np.linalg.inv(T)
print("vs truth: rot %.3g rad, trans %.3g m" % (np.arccos(min(1, (np.trace(E[:3, :3]) - 1) / 2)), np.linalg.norm(E[:3, 3])))
its = icp.iteration_times()
for row, tm in zip(icp.trace(), its):
    print("it %3d found %8d kept %8d tau %.5f mse %.6g far %8d skipped %8d | search %.3f trim %.3f solve %.3f ms" % (
        row["iter"], row["found"], row["kept"], row["tau"], row["mse"], row["far"], row["skipped"], tm[0], tm[1], tm[2]))
