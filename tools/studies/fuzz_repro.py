"""Re-runs one case of tests/test_gpu_fuzz.py::test_align_fuzz with option variants and prints the traces next to the oracle's.
usage: python tools/studies/fuzz_repro.py <seed>"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import test_gpu_fuzz as F
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
from oracle import oracle
seed = int(sys.argv[1])
rng = np.random.default_rng(5000 + seed)
kind = ["volume", "plane", "clusters", "lattice"][seed % 4]
n = int(rng.choice([3, 40, 900, 6000]))
m = F._cloud(rng, kind, n)
span = float(np.ptp(m, axis=0).max() + 1e-3)
ns = int(rng.choice([1, 3, 50, 2000, 7000]))
T = synth.rigid(rng.normal(size=3), float(rng.uniform(0, 0.15)), rng.normal(size=3) * span * 0.02)
s = m[rng.integers(0, len(m), ns)] + rng.normal(0, span * 1e-3, (ns, 3))
s = (s - T[:3, 3]) @ T[:3, :3]
T_src = synth.rigid(rng.normal(size=3), float(rng.uniform(0, 0.5)), rng.normal(size=3)) if seed % 3 == 0 else None
density = float(rng.choice([1.0, 1.0, 0.5, 0.31]))
prm = dict(max_iter=int(rng.integers(0, 25)), min_mse=float(rng.choice([1e-4, 1e-12])),
           min_mse_diff=float(rng.choice([1e-5, 1e-14])), overlap=float(rng.choice([1.0, 0.9, 0.6, 0.2, 0.01])))
print(kind, "n", n, "ns", ns, "density", density, prm, "T_src" if T_src is not None else "")
M = oracle.Model(); M.add(m)
run = oracle.Run()
r_ref = M.align(s, T=T_src, density=density, run=run, **prm)
tr_ref = run.trace()
print("oracle  iter", r_ref.iter, "pairs", r_ref.pairs, "mse", r_ref.mse)
print("  found", [t["found"] for t in tr_ref]); print("  kept ", [t["kept"] for t in tr_ref]); print("  tau  ", ["%.6g" % t["tau"] for t in tr_ref])
print("  mse  ", ["%.12g" % t["mse"] for t in tr_ref][:3])
print("  C after body 0:\n", np.array2string(np.asarray(tr_ref[0]["C"]), precision=12))
for name, opts in [("default", {}), ("moments 1 block/SM", dict(moments_blocks_per_sm=1)), ("moments 6 blocks/SM", dict(moments_blocks_per_sm=6)),
                   ("morton", dict(hilbert=0)), ("unsorted", dict(sort_source=0))]:
    icp = pkg.Icp(0)
    for k, v in opts.items():
        icp.set_option(k, v)
    icp.add_to_model(m)
    r = icp.align(s, T=T_src, density=density, **prm)
    tr = icp.trace()
    print("%-17s iter %d pairs %d mse %.6g" % (name, r.iter, r.pairs, r.mse))
    print("  found", [t["found"] for t in tr][:4]); print("  tau  ", ["%.6g" % t["tau"] for t in tr][:4]); print("  mse  ", ["%.12g" % t["mse"] for t in tr][:3])
    print("  C after body 0:\n", np.array2string(np.asarray(tr[0]["C"]), precision=12))
    print("  dC vs oracle %.3g" % np.abs(r.matrix() - r_ref.matrix()).max())
    icp.close()
