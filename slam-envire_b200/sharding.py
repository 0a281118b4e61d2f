"""Source sharding for the multi-GPU path (one process per GPU, model grid replicated).

The adapter's visiting order is fixed globally first (density stepping, reference icp/icp.hpp:126-138), then cut
into `world` contiguous ranges; rank r keeps range r.  Contiguity matters: equidistant pairs at the trim threshold are
kept in global pair-index order, which the device path realises as "lower ranks first, then local position".
"""
import numpy as np


def density_indices(n, density):
    """Indices visited by PointcloudAdapter::next(): idx = size_t(index); index += 1.0/density."""
    if density == 1.0:
        return np.arange(n, dtype=np.int64)
    out, index, inc = [], 0.0, 1.0 / density
    while int(index) < n:
        out.append(int(index))
        index += inc
    return np.asarray(out, dtype=np.int64)


def adapter_size(n, density):
    """PointcloudAdapter::size() (icp/icp.hpp:143-146)."""
    return int(n * density)


def shard_bounds(n_items, world, rank):
    """Contiguous, near-equal split of n_items over world ranks."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_source(xyz, density, world, rank):
    """(local vertices already density-stepped, global adapter size) for icpb_source_upload_shard."""
    xyz = np.asarray(xyz, dtype=np.float64)
    sel = density_indices(len(xyz), density)
    lo, hi = shard_bounds(len(sel), world, rank)
    return np.ascontiguousarray(xyz[sel[lo:hi]]), adapter_size(len(xyz), density)


def init_comm(icp, dist, device, mode="p2p"):
    """Connect the ranks' contexts through an existing torch.distributed process group.

    mode "p2p": exchange CUDA-IPC handles; the per-iteration collectives then run fused inside the kernels over NVLink
    peer memory (<= 8 ranks on one box).  mode "nccl": the library opens its own NCCL communicator instead."""
    import importlib
    import torch
    pkg = importlib.import_module("slam-envire_b200")
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return
    if mode == "p2p":
        mine = torch.frombuffer(bytearray(icp.p2p_export()), dtype=torch.uint8).to(device)
        allh = [torch.zeros(pkg.IPC_HANDLE_BYTES, dtype=torch.uint8, device=device) for _ in range(world)]
        dist.all_gather(allh, mine)
        icp.p2p_connect(world, rank, b"".join(h.cpu().numpy().tobytes() for h in allh))
        dist.barrier()
        return
    uid = torch.zeros(pkg.NCCL_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    icp.comm_init(world, rank, uid.cpu().numpy().tobytes())
