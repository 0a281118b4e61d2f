"""Study 3 (CPU, SciPy): distribution of the search radius U (distance to the previous winner after the motion,
in units of the level-0 cell edge) per iteration of cfg2 -- which share of queries a flat scan of the 2x2x2 / 3x3x3 /
4x4x4 cell block would cover before the pyramid walk is needed.  Usage: python tools/studies/u_hist.py [h]"""
import importlib, sys
import numpy as np
from scipy.spatial import cKDTree
sys.path.insert(0, ".")
synth = importlib.import_module("slam-envire_b200.synth")
h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.34641
m, s, T, prm = synth.config("cfg2", 1.0)
tree = cKDTree(m); n = len(s); C = np.eye(4)
mse, mse_diff, d_box = np.inf, np.inf, np.inf
win = None; n_po = int(n * 0.8); it = 0
edges = np.array([0, 0.25, 0.5, 0.75, 1.0, 1.5, 2.0, 3.0, 1e9])
print("it   U/h: " + "  ".join("<%.2g" % e for e in edges[1:]) + "   | d_box/h  mean(d1)/h")
while it < 300 and mse > 1e-10 and mse_diff > 1e-9:
    p = s @ C[:3, :3].T + C[:3, 3]
    d1, j1 = tree.query(p, workers=-1)
    if win is not None:
        U = np.minimum(np.linalg.norm(p - m[win], axis=1), d_box) / h
        hist = np.histogram(U, edges)[0] / n
        print("%2d        " % it + "  ".join("%.3f" % x for x in hist) + "   | %.2f  %.3f" % (d_box / h, d1.mean() / h), flush=True)
    win = j1
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
    it += 1
