"""torchrun --nproc-per-node N tools/multi_gpu_check.py : sharded align (NCCL) vs the same align on one GPU.
Every rank also runs the unsharded problem in a second context; results must agree (iter, pairs exactly; C, mse to
fp64 re-association)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
pkg = importlib.import_module("slam-envire_b200")
synth = importlib.import_module("slam-envire_b200.synth")
sharding = importlib.import_module("slam-envire_b200.sharding")

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
cases = [("cfg1", 0.2, 1.0, {}), ("cfg2", 0.05, 1.0, dict(max_iter=15)), ("cfg1", 0.1, 0.37, dict(max_iter=10, overlap=0.6)),
         ("ties", 24, 1.0, {}), ("ties", 72, 1.0, {}), ("ties", 150, 1.0, {}),
         ("tiny", 5, 1.0, {}), ("tiny", 1, 1.0, {})]  # fewer source points than ranks: some ranks hold no point at all
mode = os.environ.get("ICPB_COMM", "p2p")
icp = pkg.Icp(local)
sharding.init_comm(icp, dist, torch.device("cuda", local), mode=mode)
single = pkg.Icp(local)
for name, scale, density, over in cases:
    if name == "ties":  # lattice data: massive equidistant ties at the threshold, split across ranks: the trim's bit-by-bit
        # search over the union of the ranks' candidates; the largest lattice leaves more candidates per rank than one
        # exchange row carries (the distributed bit-by-bit search, one exchange per bit)
        g = np.arange(0, scale, dtype=np.float64)
        m = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
        m = np.concatenate([m, np.zeros((len(m), 1))], 1)
        s = m + [0.3, 0.2, 0.0]
        prm = dict(max_iter=6, min_mse=1e-12, min_mse_diff=1e-14, overlap=0.55)
    elif name == "tiny":
        m, s0, T, prm = synth.config("cfg1", 0.05)
        s = s0[:scale]
        prm = dict(prm, max_iter=5, overlap=1.0)
    else:
        m, s, T, prm = synth.config(name, scale)
        prm = dict(prm, **over)
    for c in (icp, single):
        c.clear_model()
        c.add_to_model(m)
    local_pts, size = sharding.shard_source(s, density, world, rank)
    icp.upload_source_shard(local_pts, size)
    r = icp.align_resident(**prm)
    r1 = single.align(s, density=density, **prm)
    same = (r.iter == r1.iter and r.pairs == r1.pairs and np.abs(r.matrix() - r1.matrix()).max() < 1e-9
            and (abs(r.mse - r1.mse) <= 1e-9 * abs(r1.mse) + 1e-20 or (np.isinf(r.mse) and np.isinf(r1.mse))))
    tr, tr1 = icp.trace(), single.trace()
    same = same and len(tr) == len(tr1) and all(a["kept"] == b["kept"] and a["found"] == b["found"] for a, b in zip(tr, tr1))
    # getPairsDistance after a sharded align: every rank returns ITS kept distances (ascending); together they are the
    # single-GPU list
    pd = torch.from_numpy(np.ascontiguousarray(icp.pairs_distance())).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(pd)], dtype=torch.int64, device="cuda"))
    mx = max(int(x.item()) for x in sizes)
    padded = torch.full((max(mx, 1),), float("inf"), dtype=torch.float64, device="cuda")
    padded[:len(pd)] = pd
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    allpd = np.sort(np.concatenate([p.cpu().numpy()[:int(n.item())] for p, n in zip(parts, sizes)]))
    pd1 = single.pairs_distance()
    same_pd = len(allpd) == len(pd1) and (len(pd1) == r1.pairs or r1.pairs < 3) and np.allclose(allpd, pd1, rtol=1e-9, atol=1e-12) and bool(np.all(np.diff(pd.cpu().numpy()) >= 0))
    if not same_pd and rank == 0:
        print("  pairs_distance differs: gathered", len(allpd), "single", len(pd1), "pairs", r1.pairs)
    same = same and same_pd
    ok = ok and same
    if rank == 0:
        print(mode, name, "world", world, "iter", r.iter, r1.iter, "pairs", r.pairs, r1.pairs, "mse", r.mse, r1.mse,
              "dC", np.abs(r.matrix() - r1.matrix()).max(), "OK" if same else "MISMATCH", icp.timing()["total_ms"], single.timing()["total_ms"])
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if flag.item() == 1 else "FAIL")
icp.close(); single.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
