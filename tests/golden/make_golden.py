"""Generates tests/golden/golden.json by running the REFERENCE's own icp/icp.cpp + icp/icp.hpp (oracle/_ref, built by
oracle/Makefile from the sources under /root/reference against restated third-party shims) on the seeded inputs of
cases.py.  Run in the dev container only:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import ref as R  # noqa: E402
import cases  # noqa: E402


def main():
    assert R.available(), "oracle/_ref/libicp_ref.so missing: run `make -C oracle ref` where /root/reference is mounted"
    out = {"_generator": "tests/golden/make_golden.py", "_source": "reference icp/icp.cpp + icp/icp.hpp via oracle/_ref",
           "cases": {}}
    for name, c in cases.cases().items():
        r = R.Ref()
        if c["kind"] == "ean":
            for xyz, T, dens in c["model"]:
                r.add_to_model_ean(xyz, c["normals"], c["attrs"], T=T, density=dens)
            g = r.align_ean(c["src"], c.get("src_normals", c["normals"]), c["attrs"], T=c["T_src"], density=c["density"],
                            **c["params"])
        else:
            for xyz, T, dens in c["model"]:
                r.add_to_model(xyz, T=T, density=dens)
        if c["kind"] == "ean":
            pass
        elif c["kind"] == "golden":
            g = r.align_golden(c["src"], T=c["T_src"], density=c["density"], **c["params"])
        else:
            g = r.align(c["src"], T=c["T_src"], density=c["density"], **c["params"])
        pd = g["pairs_distance"]
        out["cases"][name] = dict(
            fn_T=g["fn_T"].tolist(), iter=g["iter"], pairs=g["pairs"], mse=float(g["mse"]),
            mse_diff=float(g["mse_diff"]), overlap=float(g["overlap"]), n_pairs_distance=len(pd),
            pd_sum=float(np.sum(pd)), pd_head=[float(v) for v in pd[:4]], pd_tail=[float(v) for v in pd[-4:]],
            pd_quartiles=[float(v) for v in (np.quantile(pd, [0.25, 0.5, 0.75]) if len(pd) else [])])
        print(name, g["iter"], g["pairs"], g["mse"], g["overlap"])
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
