"""Study 2 (CPU, SciPy): cost proxy of search-skipping variants on cfg2.

For every iteration: share of queries that skip / search in the 2x2x2 block (fast) / need a wider search, and the
mean number of candidate points the fast searches scan (points in the level-0 cells that the cube of half edge R_s
around the query touches).  R_s = the radius the search is made to cover so that the bound on the runner-up is
useful: variants `fixed` (0.5 cell) and `adaptive` (U + kappa*delta + floor, capped at 0.5 cell).
Usage: python tools/studies/skip_study2.py variant kappa floor_cells [scale]
"""
import importlib, sys
import numpy as np
from scipy.spatial import cKDTree
sys.path.insert(0, ".")
synth = importlib.import_module("slam-envire_b200.synth")
variant = sys.argv[1]; kappa = float(sys.argv[2]); floor_c = float(sys.argv[3])
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
m, s, T, prm = synth.config("cfg2", scale)
h = 0.34641
o = m.min(0)
dims = (np.floor((m.max(0) - o) / h) + 1).astype(np.int64)
cid = np.floor((m - o) / h).astype(np.int64)
key = (cid[:, 2] * dims[1] + cid[:, 1]) * dims[0] + cid[:, 0]
cnt = np.bincount(key, minlength=int(dims.prod()))
def cand_count(p, R):
    lo = np.clip(np.floor((p - R[:, None] - o) / h).astype(np.int64), 0, dims - 1)
    hi = np.clip(np.floor((p + R[:, None] - o) / h).astype(np.int64), 0, dims - 1)
    tot = np.zeros(len(p), np.int64)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                c = np.stack([np.where(dx, hi[:, 0], lo[:, 0]), np.where(dy, hi[:, 1], lo[:, 1]), np.where(dz, hi[:, 2], lo[:, 2])], 1)
                use = np.ones(len(p), bool)
                if dx: use &= hi[:, 0] != lo[:, 0]
                if dy: use &= hi[:, 1] != lo[:, 1]
                if dz: use &= hi[:, 2] != lo[:, 2]
                k = (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]
                tot += np.where(use, cnt[k], 0)
    return tot
tree = cKDTree(m)
n = len(s)
C = np.eye(4)
mse, mse_diff, d_box = np.inf, np.inf, np.inf
win = np.full(n, -1); lb = np.zeros(n); p_old = None
n_po = int(n * 0.8); it = 0
tot_cost = 0.0
print("it  skip   fast   wide   cand/fast  cand_now/fast   cost/query")
while it < 300 and mse > 1e-10 and mse_diff > 1e-9:
    p = s @ C[:3, :3].T + C[:3, 3]
    d, j = tree.query(p, k=2, workers=-1)
    d1, d2, j1 = d[:, 0], d[:, 1], j[:, 0]
    if p_old is not None:
        delta = np.linalg.norm(p - p_old, axis=1)
        lb = lb - delta
        U = np.linalg.norm(p - m[win], axis=1)
        skip = U < lb
    else:
        delta = np.zeros(n); skip = np.zeros(n, bool); U = np.full(n, np.inf)
    if variant == "fixed":
        Rs = np.full(n, 0.5 * h)
    else:
        Rs = np.minimum(0.5 * h, U + kappa * delta + floor_c * h)
    fast = ~skip & (U <= 0.5 * h)
    Rs = np.maximum(Rs, np.minimum(U, 0.5 * h))
    wide = ~skip & ~fast
    lb = np.where(fast, np.minimum(d2, Rs), np.where(wide, 0.0, lb))
    cf = cand_count(p[fast], Rs[fast]).mean() if fast.any() else 0
    cn = cand_count(p[fast], U[fast]).mean() if fast.any() else 0
    cost = skip.mean() * 3 + fast.mean() * (12 + cf * 0.4) + wide.mean() * 50   # in units of ~25 instr
    tot_cost += cost
    win = j1; p_old = p
    ok = d1 <= d_box
    dd = np.where(ok, d1, np.inf)
    order = np.argsort(dd, kind="stable")[:min(n_po, int(ok.sum()))]
    tau = dd[order][-1]; d_box = 2 * tau
    P, X, D = p[order], m[j1[order]], dd[order]
    mu_p, mu_x = P.mean(0), X.mean(0)
    S = P.T @ X / len(order) - np.outer(mu_p, mu_x)
    A = S - S.T
    dl = np.array([A[1, 2], A[2, 0], A[0, 1]]); t = np.trace(S); lam = np.sqrt(t * t + dl @ dl)
    q = np.array([lam + t, *dl]); q /= np.linalg.norm(q); w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    Tm = np.eye(4); Tm[:3, :3] = R; Tm[:3, 3] = mu_x - R @ mu_p
    C = Tm @ C
    old = mse; mse = float((D * D).mean()); mse_diff = old - mse
    if it % 3 == 0 or it > 78:
        print("%2d  %.3f  %.3f  %.3f   %6.1f   %6.1f    %.1f" % (it, skip.mean(), fast.mean(), wide.mean(), cf, cn, cost), flush=True)
    it += 1
print("total cost", tot_cost)
