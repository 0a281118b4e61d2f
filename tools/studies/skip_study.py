"""Study (CPU, SciPy): how often would the exact search-skip rule fire in each iteration of a cfg2-style align?

Per query the search keeps its winner and a lower bound `lb` on the distance to every OTHER model point.  Next
iteration the query has moved by delta: lb -= delta; if dist(winner) < lb the winner provably stays the winner.
This script replays the reference loop with a k=2 kd-tree query and reports, per iteration, the share of queries
that (a) would skip, (b) need a search whose ball stays inside 2x2x2 cells ("fast"), (c) need the pyramid walk.
Usage: python tools/studies/skip_study.py [scale] [margin_cells]
"""
import importlib
import sys
import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, ".")
synth = importlib.import_module("slam-envire_b200.synth")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
margin_cells = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
m, s, T, prm = synth.config("cfg2", scale)
h = 0.3464 / np.sqrt(scale) if scale < 1 else 0.3464
margin = margin_cells * h
tree = cKDTree(m)
n = len(s)
C = np.eye(4)
mse, mse_diff, d_box = np.inf, np.inf, np.inf
win = np.full(n, -1)
lb = np.zeros(n)
p_old = None
n_po = int(n * 0.8)
it = 0
print("it  skip   fast   walk   changed  med_d1  med_gap  max_delta  mse")
while it < 300 and mse > 1e-10 and mse_diff > 1e-9:
    p = s @ C[:3, :3].T + C[:3, 3]
    d, j = tree.query(p, k=2, workers=-1)
    d1, d2, j1 = d[:, 0], d[:, 1], j[:, 0]
    if p_old is not None:
        delta = np.linalg.norm(p - p_old, axis=1)
        lb = lb - delta
        dwin = np.linalg.norm(p - m[win], axis=1)
        skip = dwin < lb
        assert np.all(win[skip] == j1[skip])
        U = dwin
    else:
        delta = np.zeros(n)
        skip = np.zeros(n, bool)
        U = np.full(n, np.inf)
    fast = ~skip & (U <= margin)
    walk = ~skip & ~fast
    # searched queries get a fresh bound: runner-up distance, capped by the radius the search covered
    lb = np.where(fast, np.minimum(d2, margin), np.where(walk, 0.0, lb))
    changed = np.mean(win != j1)
    win = j1
    p_old = p
    ok = d1 <= d_box
    dd = np.where(ok, d1, np.inf)
    order = np.argsort(dd, kind="stable")[:min(n_po, int(ok.sum()))]
    tau = dd[order][-1]
    d_box = 2 * tau
    P, X, D = p[order], m[j1[order]], dd[order]
    mu_p, mu_x = P.mean(0), X.mean(0)
    S = P.T @ X / len(order) - np.outer(mu_p, mu_x)
    A = S - S.T
    dl = np.array([A[1, 2], A[2, 0], A[0, 1]])
    t = np.trace(S)
    lam = np.sqrt(t * t + dl @ dl)
    q = np.array([lam + t, *dl])
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    Tm = np.eye(4)
    Tm[:3, :3] = R
    Tm[:3, 3] = mu_x - R @ mu_p
    C = Tm @ C
    old = mse
    mse = float((D * D).mean())
    mse_diff = old - mse
    print("%2d  %.3f  %.3f  %.3f  %.3f   %.4f  %.4f   %.5f   %.3e" % (it, skip.mean(), fast.mean(), walk.mean(), changed,
          np.median(d1), np.median(d2 - d1), delta.max(), mse), flush=True)
    it += 1
