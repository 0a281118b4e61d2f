"""Per-source-line instruction counts of one kernel: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv`
usage: python tools/studies/src_hot.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        ia = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); ism = hdr.index("# Samples")
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            out.append((int(r[ia]), int(r[it]), int(r[ism]), fname, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(o[0] for o in out); ts = sum(o[2] for o in out)
print("total warp instr", tot, "samples", ts)
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %4.1f%% lanes %4.1f  %s:%d  %s" % (100 * o[0] / tot, 100 * o[2] / max(ts, 1), o[1] / max(o[0], 1), o[3], o[4], o[5]))
