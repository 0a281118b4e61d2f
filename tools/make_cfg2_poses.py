"""Records the pose and d_box every loop body of the cfg2 align (bench.py's workload) starts from, on the GPU, into
gpurun_out/cfg2_poses.npz -- copy it to tests/golden/cfg2_poses.npz.  bench.py --impl reference times the oracle's
loop bodies from these poses (a spread sample of the COMPLETE align instead of its first two bodies), and
tests/test_golden.py checks the oracle against the recorded found / kept / tau / mse there.
usage (GPU box): python tools/make_cfg2_poses.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
pkg = importlib.import_module("slam-envire_b200")
model, src, T = bench.workload(0)
icp = pkg.Icp(0)
icp.set_option("lazy", 0)  # every query resolved: the recorded `found` does not depend on the lazy search
icp.add_to_model(model)
r = icp.align(src, **bench.ALIGN)
tr = icp.trace()
C_in, d_in = bench.poses_from_trace(tr)
out = os.path.join(ROOT, "gpurun_out", "cfg2_poses.npz")
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez(out, C_in=np.asarray(C_in), d_box_in=np.asarray(d_in), found=np.asarray([t["found"] for t in tr]),
         kept=np.asarray([t["kept"] for t in tr]), tau=np.asarray([t["tau"] for t in tr]), mse=np.asarray([t["mse"] for t in tr]))
print("iter", r.iter, "pairs", r.pairs, "mse", r.mse, "->", out)
