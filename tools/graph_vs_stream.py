"""Device time of one align: stream-enqueued loop vs the CUDA-graph WHILE loop (profile=0).
usage: python tools/graph_vs_stream.py [key=value ...]   (options for icpb_set_option, e.g. fuse_moments=0)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
for cfg, scale, over in (("cfg1", 0.2, {}), ("cfg1", 2.0, {}), ("cfg2", 1.0, dict(max_iter=300))):
    m, s, T, prm = synth.config(cfg, scale)
    prm = dict(prm, **over)
    icp = pkg.Icp(0)
    for a in sys.argv[1:]:
        icp.set_option(a.split("=")[0], float(a.split("=")[1]))
    icp.add_to_model(m)
    icp.upload_source(s)
    out = {}
    for name, opts in (("stream+events", dict(profile=1)), ("stream", dict(profile=0, graph=0)), ("graph", dict(profile=0, graph=1))):
        for k, v in opts.items():
            icp.set_option(k, v)
        ts = []
        for rep in range(6):
            r = icp.align_resident(**prm)
            ts.append(icp.timing()["total_ms"])
        out[name] = (min(ts[1:]), r.iter)
    print(cfg, len(s), "points:", "  ".join("%s %.3f ms (%d it, %.1f us/it)" % (k, v[0], v[1], 1e3 * v[0] / max(v[1], 1)) for k, v in out.items()))
    icp.close()
