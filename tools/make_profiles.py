"""Turns gpurun_out ncu artefacts into the tracked summaries under profiles/.
usage: python tools/make_profiles.py <tag> <launches.csv> [<name>=<file.ncu-rep> ...] [traffic=<name>]
traffic=<name>: the capture whose DRAM bytes bench.py quotes as roofline.traffic (-> profiles/<tag>_search_traffic.json)"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, "profiles")
rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ki].split("(")[0].replace("icpb::", ""), float(r[vi].replace(",", ""))) for r in rows[1:]]
first = next(i for i, (k, _) in enumerate(seq) if "search_kernel" in k)
tot, cnt = collections.Counter(), collections.Counter()
for k, v in seq[first:]:
    tot[k] += v
    cnt[k] += 1
s = sum(tot.values())
with open(os.path.join(out, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: ncu launch list of one cfg2 align (1M vs 1M, 90 iterations)\n\n" % tag)
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv python tools/ncu_align.py 95` (tools/r2_capture.sh)\n")
    f.write("(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes).\n\n")
    f.write("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|\n")
    for k, v in tot.most_common():
        f.write("| %s | %d | %.1f | %.1f %% | %.1f |\n" % (k, cnt[k], v / 1e3, 100 * v / s, v / cnt[k] / 1e3))
    idx = [i for i, (k, _) in enumerate(seq) if "search_kernel" in k]
    for which in (0, 1, 40, len(idx) - 2):
        f.write("\nIteration %d:\n\n```\n" % which)
        for k, v in seq[idx[which]:idx[which + 1]]:
            f.write("%-28s %10.1f us\n" % (k, v / 1e3))
        f.write("```\n")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
traffic_name = None
dram = {}
for arg in sys.argv[3:]:
    name, rep = arg.split("=")
    if name == "traffic":
        traffic_name = rep
        continue
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"]).decode()
    r = list(csv.reader(txt.splitlines()))
    h, u = r[0], r[1]
    with open(os.path.join(out, "%s_%s.md" % (tag, name)), "w") as f:
        f.write("# %s %s: ncu --set full --clock-control none (source: %s)\n\n" % (tag, name, os.path.basename(rep)))
        for row in r[2:]:
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(row[h.index("dram__bytes_read.sum")]) * scale[u[h.index("dram__bytes_read.sum")]]
                wr = float(row[h.index("dram__bytes_write.sum")]) * scale[u[h.index("dram__bytes_write.sum")]]
                dram.setdefault(name, (rd + wr, float(row[h.index("gpu__time_duration.sum")])))
            except (ValueError, KeyError):
                pass
            f.write("## %s (id %s)\n\n| metric | value | unit |\n|---|---|---|\n" % (row[h.index("Kernel Name")].split("(")[0], row[0]))
            for k in KEYS:
                if k in h:
                    f.write("| %s | %s | %s |\n" % (k, row[h.index(k)], u[h.index(k)]))
            f.write("\n")
if traffic_name and traffic_name in dram:
    others = ", ".join("%s %.1f MB" % (k, v[0] / 1e6) for k, v in sorted(dram.items()) if k != traffic_name)
    with open(os.path.join(out, "%s_search_traffic.json" % tag), "w") as f:
        json.dump({"dram_bytes_per_launch": dram[traffic_name][0],
                   "source": "ncu --set full --clock-control none (cold-cache replay), search_kernel of loop body %s of the cfg2 align, "
                             "profiles/%s_%s.md; other captures: %s" % (traffic_name.replace("search_it", ""), tag, traffic_name, others)}, f)
print("written to", out)
