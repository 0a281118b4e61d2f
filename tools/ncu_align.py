"""Short single-align program for ncu captures.  Usage: python tools/ncu_align.py [max_iter] [scale]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
max_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 30
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
m, s, T, prm = synth.config("cfg2", scale)
icp = pkg.Icp(0)
icp.set_option("profile", 1)  # stream-enqueued kernels (ncu launch lists / -s / -c counts refer to these)
icp.add_to_model(m)
r = icp.align(s, **dict(prm, max_iter=max_iter))
print("iter", r.iter, "mse", r.mse, icp.timing())
