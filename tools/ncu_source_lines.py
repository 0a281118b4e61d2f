"""Per-source-line instruction counts of an .ncu-rep captured with --import-source on (compiled with -lineinfo).
usage: python tools/ncu_source_lines.py file.ncu-rep [top_n]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stderr=subprocess.DEVNULL).decode()
rows, fname, hdr = [], None, None
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
        continue  # SASS rows have an empty line number
    ie, te, smp = float(r[hdr.index("Instructions Executed")] or 0), float(r[hdr.index("Thread Instructions Executed")] or 0), float(r[hdr.index("# Samples")] or 0)
    if ie > 0 or smp > 0:
        rows.append((ie, te, smp, fname, r[0], r[1].strip()[:110]))
ti, ts = sum(r[0] for r in rows), sum(r[2] for r in rows)
print("total warp-instructions %.4g, samples %d" % (ti, ts))
for ie, te, smp, f, ln, src in sorted(rows, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  %s:%s  %s" % (100 * ie / ti, 100 * smp / max(ts, 1), te / max(ie, 1), f, ln, src))
