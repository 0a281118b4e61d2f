"""Warp-instructions of search_kernel by part of the search (source-line ranges of grid_view.h / kernels.cuh).
usage: python tools/ncu_source_groups.py file.ncu-rep"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stderr=subprocess.DEVNULL).decode()
GV = [(132, 182, "fp32 bounds (gap / cell / node mindist, radii)"), (183, 235, "bound refresh, exact evaluation, fp32 distance"),
      (236, 291, "marks flush + 4-run scan (fast path)"), (292, 309, "walk: cell scan"), (310, 387, "walk: traversal"),
      (388, 453, "init / start level / plan"), (454, 479, "fast path: runs of the 2x2x2 block"), (480, 499, "walk: roots"),
      (500, 515, "lower bound for the skip test"), (516, 578, "block scan: enumerate rows / cells"),
      (579, 606, "block scan: candidate loop"), (607, 651, "dispatch"), (652, 760, "skip / defer tests, motion bound")]
rows, fname, hdr = [], None, None
agg, lanes = collections.Counter(), collections.Counter()
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 12 or r[0] == "":
        continue
    ie, te = float(r[hdr.index("Instructions Executed")] or 0), float(r[hdr.index("Thread Instructions Executed")] or 0)
    ln = int(r[0])
    if fname == "grid_view.h":
        g = next((n for a, b, n in GV if a <= ln <= b), "grid_view.h other")
    elif fname == "kernels.cuh":
        g = "kernel body: transform, warm start, finish, histogram, counters"
    else:
        g = fname
    agg[g] += ie
    lanes[g] += te
tot = sum(agg.values())
print("total warp-instructions %.4g, lanes %.1f" % (tot, sum(lanes.values()) / tot))
for g, v in agg.most_common():
    print("%5.1f%%  lanes %4.1f  %s" % (100 * v / tot, lanes[g] / max(v, 1), g))
