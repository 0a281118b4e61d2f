#!/usr/bin/env python
"""bench.py -- trimmed-ICP hot path on B200: ICP correspondences/s and ms per align.

Workload (BASELINE.json configs[1], SURVEY.md section 8d "cfg2"): 1M-point model vs 1M-point source,
perturbed copy (3 deg, 0.3 m) + Gaussian noise sigma 0.01 m, overlap 0.8.  One STEP = one complete
Trimmed::_align (all iterations until the reference's own stopping rule fires, ~90) of that workload.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (B200 kernels through the C ABI)
  python bench.py --impl reference ...                         the CPU restatement of the reference on host cores

N > 1 (torchrun, one rank per GPU): the source is sharded (each rank holds its own 1M-point shard, weak
scaling), the model grid is replicated, and the per-iteration exchanges run fused over NVLink peer memory.
Besides the headline line the run carries:
  parity           the oracle (CPU restatement of the reference) evaluated on the SAME full-size input at sampled
                   iterations of the align, from the device's own poses: found / kept / tau / mse must agree
  multi_gpu_check  (N > 1) rank 0 re-runs the align unsharded on one GPU and compares it with the sharded run
  strong           BASELINE configs[2] (10M vs 10M, overlap 0.5) and configs[4] (100M vs 20M) with the source
                   sharded over the N GPUs (strong scaling: fixed total work)
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_MODEL = 1_000_000
N_SOURCE = 1_000_000
# max_iter is a cap only: the reference's own stopping rule (mse_diff <= min_mse_diff) ends the align after ~90 iterations
ALIGN = dict(max_iter=300, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.8)
WORKLOAD = "cfg2: 1M-pt model vs 1M-pt source, 3deg/0.3m + sigma 0.01 m noise, overlap 0.8"
METRIC = "icp_correspondences_per_sec"
UNIT = "correspondences/s"
POSES = os.path.join(ROOT, "tests", "golden", "cfg2_poses.npz")  # tools/make_cfg2_poses.py (device trace of this workload)
TRAFFIC = os.path.join(ROOT, "profiles", "r2_search_traffic.json")  # tools/make_profiles.py (ncu --set full capture)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def search_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per search_kernel launch from the committed ncu capture, or None."""
    try:
        with open(TRAFFIC) as f:
            return json.load(f)
    except Exception:
        return None


def workload(rank=0):
    synth = importlib.import_module("slam-envire_b200.synth")
    model = synth.scene(N_MODEL, 100.0, 2)
    # every rank's shard is an independent noisy observation of the same scene under the same rigid motion
    # (ICPB_SAME_SHARDS=1, a study switch: every rank gets rank 0's shard -- identical work per rank, no skew)
    src, T = synth.perturbed_copy(model[:N_SOURCE], 3.0, 0.3, 0.01, seed=3 + (0 if os.environ.get("ICPB_SAME_SHARDS") else rank))
    return model, src, T


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/: test infrastructure, used here only as the timed baseline and as the checker)
# ---------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The oracle's kd-tree over the model, and single loop bodies of Trimmed::_align evaluated from a given pose."""

    def __init__(self, model):
        from oracle import oracle as O
        O.build()
        self.O = O
        t0 = time.perf_counter()
        self.M = O.Model()
        self.M.add(model)
        self.build_s = time.perf_counter() - t0
        self.model = model

    def iteration(self, src, C_in, d_box_in, n_po, threads, subsample=1):
        """One loop body (icp/icp.hpp:466-485) from pose C_in with ball d_box_in: findPairs, trim, getTransform."""
        O = self.O
        q = src[::subsample]
        t0 = time.perf_counter()
        p = O.transform_points(C_in, q)
        idx, d = self.M.find_pairs(p, d_box=d_box_in, threads=threads)
        t_pairs = time.perf_counter() - t0
        ok = idx != 0xFFFFFFFF
        out = dict(found=int(ok.sum()), secs_pairs=t_pairs)
        if subsample == 1:
            dd = d[ok]
            order, tau = O.trim(dd, n_po)
            out.update(kept=len(order), tau=tau)
            if len(order) >= 3:
                T, mse = O.get_transform(p[ok], self.model[idx[ok]], dd, order)
                out.update(mse=mse, T=T)
        out["secs"] = time.perf_counter() - t0
        return out


def poses_from_trace(trace):
    """(C_in, d_box_in) of every executed loop body, from a device trace (row k holds the state AFTER body k)."""
    C_in, d_in = [np.eye(4)], [np.inf]
    for row in trace[:-1]:
        C_in.append(np.asarray(row["C"]))
        d_in.append(float(row["d_box"]))
    return C_in, d_in


def sample_iterations(n_iter, count):
    return sorted(set(int(round(x)) for x in np.linspace(0, n_iter - 1, count)))


def cpu_leg(ref, src, C_in, d_in, iters, n_po, threads):
    rows, secs = [], 0.0
    for k in iters:
        r = ref.iteration(src, C_in[k], d_in[k], n_po, threads)
        rows.append((k, r))
        secs += r["secs"]
    return rows, len(src) * len(iters) / secs, 1e3 * secs / len(iters)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    model, src, _ = workload(0)
    ref = CpuReference(model)
    n_po = int(int(len(src) * 1.0) * ALIGN["overlap"])
    if os.path.exists(POSES):
        z = np.load(POSES)
        C_in, d_in = list(z["C_in"]), list(z["d_box_in"])
        iters = sample_iterations(len(C_in), 5)
        what = "loop bodies %s of the %d of the complete align, each from the pose the align has there (%s)" % (
            iters, len(C_in), os.path.relpath(POSES, ROOT))
    else:  # no recorded poses: the first bodies of the align as the oracle itself runs them
        C_in, d_in, iters, what = None, None, [0, 1], "first 2 loop bodies of the align"
    vals, budget, t_start = [], 150.0, time.perf_counter()
    for k in range(args.warmup + args.steps):
        if C_in is not None:
            rows, v, ms = cpu_leg(ref, src, C_in, d_in, iters, n_po, cores)
        else:
            run = ref.O.Run()
            t0 = time.perf_counter()
            ref.M.align(src, run=run, threads=cores, **dict(ALIGN, max_iter=2))
            dt = time.perf_counter() - t0
            v, ms = len(src) * 2 / dt, 1e3 * dt / 2
        if k >= args.warmup or not vals:
            vals.append((v, ms))
        if time.perf_counter() - t_start > budget:
            break
    v = float(np.mean([a for a, _ in vals]))
    ms_iter = float(np.mean([b for _, b in vals]))
    # the faithful figure: the reference is single-threaded; one mid-align loop body, 1/16 of the queries
    one = None
    if C_in is not None:
        k = iters[len(iters) // 2]
        r1 = ref.iteration(src, C_in[k], d_in[k], n_po, 1, subsample=16)
        one = {"value": (len(src) // 16) / r1["secs_pairs"], "unit": UNIT, "cores": 1,
               "sample": "findPairs of loop body %d on every 16th query, one thread" % k}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": ms_iter * 90, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "align": ALIGN, "ms_per_iteration": ms_iter,
                   "note": "CPU restatement of the reference (oracle/icp_oracle.c): the reference itself cannot be "
                           "built here (no Eigen/Boost/libkdtree++); OpenMP over queries on all host cores; "
                           "ms_per_step = sampled ms per loop body x 90 bodies; kd-tree build %.1f s not counted" % ref.build_s},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": what, "one_core": one},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def check_parity(ref, src, trace, n_po, threads, count=3):
    """The oracle on the same full-size input, at `count` loop bodies spread over the align, from the device's poses."""
    C_in, d_in = poses_from_trace(trace)
    iters = sample_iterations(len(trace), count)
    rows, v, ms = cpu_leg(ref, src, C_in, d_in, iters, n_po, threads)
    out = {"iterations_checked": iters, "found": [], "kept": [], "tau_rel": [], "mse_rel": [], "ok": True}
    for k, r in rows:
        t = trace[k]
        out["found"].append([t["found"], r["found"]])
        out["kept"].append([t["kept"], r.get("kept")])
        tau_rel = abs(t["tau"] - r["tau"]) / max(abs(r["tau"]), 1e-300)
        mse_rel = abs(t["mse"] - r.get("mse", np.nan)) / max(abs(r.get("mse", np.nan)), 1e-300)
        out["tau_rel"].append(tau_rel)
        out["mse_rel"].append(mse_rel)
        # found, kept: exact.  tau: the same distance bit for bit unless an ulp-level near-tie sits at the cut.  mse: the
        # 17 sums are reduced in a different order.
        if t["found"] != r["found"] or t["kept"] != r.get("kept") or not tau_rel <= 1e-12 or not mse_rel <= 1e-9:
            out["ok"] = False
    return out, v, ms, iters


def strong_config(name, world, rank):
    """(model, this rank's source shard, global source size, align params) of BASELINE configs[2] / configs[4]."""
    synth = importlib.import_module("slam-envire_b200.synth")
    sharding = importlib.import_module("slam-envire_b200.sharding")
    if name == "cfg3":
        m, s, T, prm = synth.config("cfg3")
        lo, hi = sharding.shard_bounds(len(s), world, rank)
        return m, np.ascontiguousarray(s[lo:hi]), len(s), prm
    # cfg5: 100M-point source vs 20M-point model, extent 500 m.  The source is generated shard by shard (8 independent
    # 12.5M-point observations of the scene under the same rigid motion) so that no host has to hold 100M points twice
    n_src, n_model, ext, parts = 100_000_000, 20_000_000, 500.0, 8
    m = synth.scene(n_model, ext, 7)
    per = n_src // parts
    mine = [p for p in range(parts) if p * world // parts == rank]
    shards = []
    for p in mine:
        base = synth.scene(per, ext, 60 + p)
        shards.append(synth.perturbed_copy(base, 1.0, 0.3, 0.01, seed=70 + p)[0])
    return m, np.ascontiguousarray(np.concatenate(shards)), n_src, dict(max_iter=10, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.8)


def run_strong(pkg, name, world, rank, local, dist, torch, comm_mode, reps=3):
    sharding = importlib.import_module("slam-envire_b200.sharding")
    m, s_local, n_global, prm = strong_config(name, world, rank)
    icp = pkg.Icp(local)
    if world > 1:
        sharding.init_comm(icp, dist, torch.device("cuda", local), mode=comm_mode)
    icp.add_to_model(m)
    gi = icp.build()
    if world > 1:
        icp.upload_source_shard(s_local, n_global)
    else:
        icp.upload_source(s_local)
    ms = []
    for k in range(1 + reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        r = icp.align_resident(**prm)
        t = torch.tensor([icp.timing()["total_ms"]], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if k:
            ms.append(float(t.item()))
    iters = len(icp.trace())
    out = {"workload": {"cfg3": "cfg3: 10M vs 10M, half of the source replaced by a disjoint region, overlap 0.5",
                        "cfg5": "cfg5: 100M-pt source vs 20M-pt replicated model grid, overlap 0.8, first 10 loop bodies"}[name],
           "n_gpus": world, "source_points_total": int(n_global), "model_points": int(len(m)), "iterations": iters,
           "ms_per_align": float(np.mean(ms)), "ms_per_iteration": float(np.mean(ms)) / max(iters, 1),
           "correspondences_per_s": n_global * iters / (np.mean(ms) * 1e-3), "grid_build_ms": gi.build_ms,
           "final": {"iter": r.iter, "pairs": r.pairs, "mse": r.mse}}
    icp.close()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("slam-envire_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    model, src, T_true = workload(rank)
    n_s = len(src)
    icp = pkg.Icp(local)
    comm_mode = os.environ.get("ICPB_COMM", "p2p")
    if world > 1:
        sharding = importlib.import_module("slam-envire_b200.sharding")
        sharding.init_comm(icp, dist, torch.device("cuda", local), mode=comm_mode)
    icp.add_to_model(model)
    gi = icp.build()

    # pinned host copy of this rank's source for the end-to-end leg
    src_pinned = torch.from_numpy(src).pin_memory()
    src_host = src_pinned.numpy()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def upload():
        if world > 1:
            icp.upload_source_shard(src_host, n_s * world)
        else:
            icp.upload_source(src_host)

    # ---- device-resident leg: `value` ----
    # profile=1: the loop is enqueued kernel by kernel with CUDA events around every phase on the library's stream, so
    # the per-kernel durations behind `roofline` come from the very region that `value` is timed on
    icp.set_option("profile", 1)
    upload()
    res = None
    for _ in range(args.warmup):
        flush.zero_()
        barrier()
        res = icp.align_resident(**ALIGN)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, search_ms, trim_ms, solve_ms, iters, launches = [], 0.0, 0.0, 0.0, 0, 0
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        res = icp.align_resident(**ALIGN)
        t = icp.timing()
        dev_ms.append(t["total_ms"])
        search_ms += t["search_ms"]
        trim_ms += t["trim_ms"]
        solve_ms += t["solve_ms"]
        iters += t["iterations"]
        launches += t["kernels_launched"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = float(np.sum(dev_ms))
    if world > 1:
        tt = torch.tensor([step_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms = float(tt.item())
    trace = icp.trace()
    it_times = icp.iteration_times()
    n_iter_per_step = len(trace)
    total_corr = float(n_s) * world * n_iter_per_step * args.steps
    value = total_corr / (step_ms * 1e-3)

    # ---- end-to-end leg: host buffers in, result out, every step; library defaults (the whole _align loop is one
    # CUDA-graph launch driven by the device-side loop condition) ----
    icp.set_option("profile", 0)
    graph_ms = []
    for _ in range(max(2, args.steps)):
        flush.zero_()
        barrier()
        icp.align_resident(**ALIGN)
        graph_ms.append(icp.timing()["total_ms"])
    e2e_s = []
    for k in range(max(1, min(args.warmup, 2)) + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            icp.upload_source_shard(src_host, n_s * world)
            r2 = icp.align_resident(**ALIGN)
        else:
            r2 = icp.align(src_host, **ALIGN)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if k >= max(1, min(args.warmup, 2)):
            e2e_s.append(dt)
    e2e_total = float(np.sum(e2e_s))
    if world > 1:
        tt = torch.tensor([e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_total = float(tt.item())
    e2e_value = float(n_s) * world * len(icp.trace()) * len(e2e_s) / e2e_total
    assert r2.iter == res.iter and r2.pairs == res.pairs, "graph loop and stream loop disagree"

    # ---- multi-GPU parity: rank 0 re-runs the align unsharded (all shards on one GPU) ----
    mg_check = None
    if world > 1 and rank == 0:
        one = pkg.Icp(local)
        one.add_to_model(model)
        full = np.concatenate([src] + [workload(r)[1] for r in range(1, world)])
        r1 = one.align(full, **ALIGN)
        t1 = one.trace()
        dC = float(np.abs(r1.matrix() - res.matrix()).max())
        same = (r1.iter == res.iter and r1.pairs == res.pairs and len(t1) == len(trace) and
                all(a["kept"] == b["kept"] and a["found"] == b["found"] for a, b in zip(t1, trace)))
        mg_check = {"result": "pass" if (same and dC < 1e-12) else "FAIL", "max_abs_dC": dC, "iter": [res.iter, r1.iter],
                    "pairs": [res.pairs, r1.pairs], "per_iteration_found_kept_equal": bool(same),
                    "what": "sharded over %d GPUs vs the same %d source points on one GPU" % (world, len(full))}
        one.close()
    icp.close()

    # ---- strong scaling on the configs BASELINE names for N > 1 ----
    strong = {}
    if not args.no_strong:
        for name in ("cfg3", "cfg5"):
            if name == "cfg5" and world not in (1, 8) and not args.cfg5:
                continue
            try:
                strong[name] = run_strong(pkg, name, world, rank, local, dist, torch, comm_mode)
            except Exception as e:  # a strong block must not take the headline down (e.g. host memory on a small box)
                strong[name] = {"error": repr(e)[:200]}

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        # K3 algorithmic bytes per launch: per source point 24 B read + 4 B idx + 8 B distance written,
        # plus the model read once at 16 B/pt (DESIGN.md "Roofline accounting")
        k3_bytes = n_s * 36.0 + len(model) * 16.0
        k3_ms = search_ms / max(iters, 1)
        achieved = k3_bytes / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else 0.0
        traffic = search_traffic()
        conv = [row[0] for row in it_times[-5:]] if len(it_times) >= 5 else []
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "align": ALIGN, "iterations_per_align": n_iter_per_step,
                       "ms_per_iteration": step_ms / args.steps / max(n_iter_per_step, 1),
                       "sharding": "source sharded x%d (1M points per GPU), model grid replicated" % world,
                       "collectives": ("none" if world == 1 else ("fused into the trim/moments kernels over NVLink peer memory"
                                                                   if comm_mode == "p2p" else "NCCL all-reduce / all-gather")),
                       "l2": "256 MB buffer written between steps (L2 flush)",
                       "grid": {"cell_size": gi.cell_size, "dims": list(gi.dims), "levels": gi.levels,
                                "occupied_cells": gi.n_occupied, "build_ms": gi.build_ms},
                       "final": {"mse": res.mse, "pairs": res.pairs, "iter": res.iter},
                       "phase_ms_per_iteration": {"search": search_ms / max(iters, 1), "trim": trim_ms / max(iters, 1),
                                                  "moments_pose": solve_ms / max(iters, 1)},
                       "search_ms_converged_iterations": float(np.mean(conv)) if conv else None,
                       "queries_per_iteration": {"skipped": float(np.mean([t["skipped"] for t in trace])),
                                                 "deferred": float(np.mean([t["deferred"] for t in trace])),
                                                 "bodies_run_again": int(trace[-1]["redone"]) if trace else 0},
                       "wall_s_timed_region": wall,
                       "graph_loop_ms_per_step": float(np.mean(graph_ms[1:])),
                       "note": "value/roofline: stream-enqueued loop with per-phase CUDA events; e2e and "
                               "graph_loop_ms_per_step: default mode, one CUDA-graph launch per align"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(src_host.nbytes),
                    "d2h_bytes_per_step": 200 + 4 * len(trace), "ms_per_step": 1e3 * e2e_total / len(e2e_s)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "search_kernel (K2+K3)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else "no ncu capture of this kernel committed",
                         "peak_source": peak_kind, "launch_ms": k3_ms, "algorithmic_bytes_per_launch": k3_bytes},
        }
        if mg_check:
            line["multi_gpu_check"] = mg_check
        if strong:
            line["strong"] = strong
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            ref = CpuReference(model)
            n_po = int(int(n_s * 1.0) * ALIGN["overlap"])
            parity, v, ms, its = check_parity(ref, src, trace, n_po, cores, count=4)
            line["parity"] = parity
            k = its[len(its) // 2]
            C_in, d_in = poses_from_trace(trace)
            r1 = ref.iteration(src, C_in[k], d_in[k], n_po, 1, subsample=16)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "loop bodies %s of the %d of the same align, each from the pose the device "
                                              "had there, OpenMP over queries (%.0f ms per body; kd-tree build %.1f s "
                                              "not counted)" % (its, len(trace), ms, ref.build_s),
                                    "one_core": {"value": (n_s // 16) / r1["secs_pairs"], "unit": UNIT, "cores": 1,
                                                 "sample": "findPairs of loop body %d on every 16th query, one thread "
                                                           "(the reference is single-threaded)" % k}}
        print(json.dumps(line))
        if line.get("parity") and not line["parity"]["ok"]:
            raise SystemExit("bench.py: PARITY FAILURE against the oracle: %s" % json.dumps(line["parity"]))
        if mg_check and mg_check["result"] != "pass":
            raise SystemExit("bench.py: multi-GPU run differs from the single-GPU run: %s" % json.dumps(mg_check))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the parity / cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling blocks (cfg3, cfg5)")
    ap.add_argument("--cfg5", action="store_true", help="run the cfg5 strong-scaling block at any N (default: N = 1 and 8)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
