"""Would a cache of the K nearest model points per query let later loop bodies of the cfg2 align skip the search?
For refresh bodies t0: K nearest + a lower bound lb on everything else (the (K+1)-th nearest, capped by the covered
radius R = d1 + margin); for t > t0 the cached set still holds the exact nearest neighbour iff best_cached(t) < lb -
motion.  Prints the fraction of (kept-range) queries that stay valid.  CPU only (scipy cKDTree)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy.spatial import cKDTree
import bench
z = np.load(os.path.join(os.path.dirname(bench.__file__), "tests", "golden", "cfg2_poses.npz"))
model, src, _ = bench.workload(0)
sub = np.arange(0, len(src), 10)  # every 10th query
src = src[sub]
tree = cKDTree(model)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
margin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.10
h = 0.3464
def pos(t):
    C = z["C_in"][t]
    return src @ C[:3, :3].T + C[:3, 3]
for t0 in (5, 15, 30, 45, 60, 75):
    q0 = pos(t0)
    d, idx = tree.query(q0, k=K + 1, workers=-1)
    R = np.minimum(d[:, 0] + margin, 0.98 * h)              # radius the refresh search covers (block scan limit)
    lb = np.minimum(d[:, K], np.maximum(R, d[:, 0]))         # everything not cached is at least this far
    cached = idx[:, :K].copy()
    cached[d[:, :K] > R[:, None]] = -1                       # only points inside the covered radius are cached
    tau = z["tau"][t0]
    line = []
    for dt in (1, 2, 3, 5, 8, 12):
        t = t0 + dt
        if t >= len(z["tau"]):
            break
        q = pos(t)
        mot = np.linalg.norm(q - q0, axis=1)
        dc = np.where(cached >= 0, np.linalg.norm(model[np.maximum(cached, 0)] - q[:, None, :], axis=2), np.inf).min(axis=1)
        valid = dc < lb - mot
        rel = d[:, 0] <= 1.05 * tau                          # queries the lazy search would not defer anyway
        line.append("%d:%.2f" % (dt, valid[rel].mean()))
    print("t0 %2d tau %.3f  median d1 %.3f  cached/query %.1f  median margin %.3f  median motion/body %.4f | valid after dt bodies: %s" % (
        t0, tau, np.median(d[:, 0]), (cached >= 0).sum(1).mean(), np.median(lb - d[:, 0]), np.median(np.linalg.norm(pos(t0 + 1) - q0, axis=1)), " ".join(line)))
