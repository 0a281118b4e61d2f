"""Runs one of the BASELINE.json configurations at (or near) full size and prints one JSON line.
   python tools/run_config.py cfg3            (10M vs 10M, 50 % disjoint, overlap 0.5)
   torchrun --nproc-per-node 8 tools/run_config.py cfg5      (100M source sharded vs 20M replicated model)
   optional: scale=<f> max_iter=<n>"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
sharding = importlib.import_module("slam-envire_b200.sharding")

cfg = sys.argv[1]
opts = dict(a.split("=") for a in sys.argv[2:])
scale = float(opts.get("scale", 1.0))
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

t0 = time.time()
if cfg == "cfg5":
    n_model, n_src_total, ext = int(20_000_000 * scale), int(100_000_000 * scale), 500.0
    model = synth.scene(n_model, ext, 7)
    lo, hi = sharding.shard_bounds(n_src_total, world, rank)
    dense = synth.scene(hi - lo, ext, 600 + rank)                      # this rank's part of the 100M-point source
    src, T = synth.perturbed_copy(dense, 1.0, 0.3, 0.01, seed=60 + rank)
    prm = dict(max_iter=int(opts.get("max_iter", 30)), min_mse=1e-10, min_mse_diff=1e-9, overlap=0.8)
    size = n_src_total
else:
    model, full, T, prm = synth.config(cfg, scale)
    prm = dict(prm, max_iter=int(opts.get("max_iter", prm["max_iter"])))
    lo, hi = sharding.shard_bounds(len(full), world, rank)
    src, size = np.ascontiguousarray(full[lo:hi]), len(full)
    ext = float(np.ptp(model[:, 0]))
gen_s = time.time() - t0

icp = pkg.Icp(local)
icp.set_option("profile", 1)
if world > 1:
    sharding.init_comm(icp, dist, torch.device("cuda", local), mode=os.environ.get("ICPB_COMM", "p2p"))
t0 = time.time()
icp.add_to_model(model)
gi = icp.build()
add_s = time.time() - t0
t0 = time.time()
if world > 1:
    icp.upload_source_shard(src, size)
else:
    icp.upload_source(src)
up_s = time.time() - t0
r = icp.align_resident(**prm)      # warm-up (first launch of every kernel, peer connections)
if world > 1:
    dist.barrier()
r = icp.align_resident(**prm)
tm = icp.timing()
ms = torch.tensor([tm["total_ms"]], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
E = r.matrix() @ np.linalg.inv(T)
if rank == 0:
    its = icp.iteration_times()
    n_it = len(icp.trace())
    print(json.dumps({
        "config": cfg, "scale": scale, "n_gpus": world, "model_points": len(model), "source_points_total": size,
        "source_points_this_rank": len(src), "align": prm, "iterations": n_it, "pairs": r.pairs, "mse": r.mse,
        "ms_per_align": float(ms.item()), "ms_per_iteration": float(ms.item()) / max(n_it, 1),
        "correspondences_per_s": size * n_it / (float(ms.item()) * 1e-3),
        "search_ms_first_last": [float(its[0][0]), float(its[-1][0])] if len(its) else None,
        "phase_ms": {k: tm[k] for k in ("search_ms", "trim_ms", "solve_ms")},
        "grid": {"cell_size": gi.cell_size, "dims": list(gi.dims), "levels": gi.levels, "cells": gi.n_cells,
                 "occupied": gi.n_occupied, "build_ms": gi.build_ms},
        "host_s": {"generate": gen_s, "add_model_incl_h2d": add_s, "upload_source": up_s},
        "vs_truth": {"rot_rad": float(np.arccos(min(1.0, (np.trace(E[:3, :3]) - 1) / 2))), "trans_m": float(np.linalg.norm(E[:3, 3]))},
        "gpu_mem_gb": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}))
icp.close()
if world > 1:
    dist.destroy_process_group()
