"""Growing model: cost of addToModel on a built index.  Builds a model of N points, then appends K
batches of B points inside the grid box and prints the device time of every merge (grid_info.build_ms) and the wall
time of the call.  usage: python tools/append_profile.py [N] [B] [K] [key=value ...]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
args = [a for a in sys.argv[1:] if "=" not in a]
N = int(args[0]) if len(args) > 0 else 10_000_000
B = int(args[1]) if len(args) > 1 else 100_000
K = int(args[2]) if len(args) > 2 else 5
ext = 140.0 * (N / 1e6) ** 0.5
m = synth.scene(N, ext, 4)
icp = pkg.Icp(0)
for a in sys.argv[1:]:
    if "=" in a:
        icp.set_option(a.split("=")[0], float(a.split("=")[1]))
t0 = time.time()
icp.add_to_model(m)
gi = icp.build()
print("build: %d points, h=%.4f dims=%s levels=%d: %.3f ms on the device, %.1f ms wall (incl. H2D of %d MB)" % (
    N, gi.cell_size, list(gi.dims), gi.levels, gi.build_ms, 1e3 * (time.time() - t0), N * 24 // 1000000))
rng = np.random.default_rng(1)
for k in range(K):
    c = rng.uniform(-0.3 * ext, 0.3 * ext, 2)
    b = m[rng.integers(0, N, B)] * [0.05, 0.05, 1.0] + [c[0], c[1], 0.0] + rng.normal(0, 0.01, (B, 3))
    b[:, 2] = np.clip(b[:, 2], m[:, 2].min(), m[:, 2].max())
    t0 = time.time()
    icp.add_to_model(b)
    wall = 1e3 * (time.time() - t0)
    gi = icp.build()
    print("append %d: %d points into %d: merge %.3f ms on the device, call %.2f ms wall, merges so far %d" % (
        k, B, gi.n_points - B, gi.build_ms, wall, gi.merges))
# a second full build of the final model, warm (allocations done)
for rep in range(3):  # the first rebuild may still grow buffers; the later ones are the steady state
    icp.set_option("cell_size", 0.0)  # marks the grid dirty
    t0 = time.time()
    gi = icp.build()
    print("rebuild %d of %d points: %.3f ms on the device, %.2f ms wall" % (rep, gi.n_points, gi.build_ms, 1e3 * (time.time() - t0)))
