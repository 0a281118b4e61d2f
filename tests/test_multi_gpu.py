"""N > 1: (a) GPU: sharded align over NCCL vs one GPU (needs >= 2 devices; gpurun --gpus 2); (b) CPU: the sharded
protocol itself -- per-shard correspondences, radix-select with all-reduced histograms, tie quota by rank order,
all-reduced moments -- restated over torch.distributed/gloo with world_size 2 and checked against the unsharded oracle."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_sharded_align_matches_single_gpu(mode):
    """p2p: collectives fused into the kernels over NVLink peer memory (CUDA IPC); nccl: the library's own communicator."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, ICPB_COMM=mode)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533" if mode == "p2p" else "29534",
                          os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600,
                         env=env)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def _cli_run(exe, tmp_path, model, src, kind, prm, devices, density=1.0):
    import json
    mp, sp = str(tmp_path / "m.bin"), str(tmp_path / "s.bin")
    np.ascontiguousarray(model, dtype=np.float64).tofile(mp)
    np.ascontiguousarray(src, dtype=np.float64).tofile(sp)
    args = [exe, mp, sp] + ["%.17g" % v for v in np.eye(4).reshape(16)] + ["%.17g" % density, kind, str(prm["max_iter"]),
                                                                           "%.17g" % prm["min_mse"], "%.17g" % prm["min_mse_diff"]]
    args += ["%.17g" % prm[k] for k in (("alpha", "beta", "eps") if kind == "golden" else ("overlap",))]
    env = dict(os.environ)
    env.pop("ICP_B200_DEVICES", None)
    if devices is not None:
        env["ICP_B200_DEVICES"] = str(devices)
    out = json.loads(subprocess.check_output(args, env=env, timeout=600).decode())
    assert "error" not in out, out
    return out


def _same_align(a, b, extent):
    from conftest import assert_transform_close
    assert a["iter"] == b["iter"] and a["pairs"] == b["pairs"] and a["n_pairs_distance"] == b["n_pairs_distance"]
    assert_transform_close(np.array(a["fn_T"]), np.array(b["fn_T"]), extent)
    assert abs(a["mse"] - b["mse"]) <= 1e-9 * abs(b["mse"]) + 1e-300
    assert a["pd_sorted"] and abs(a["pd_sum"] - b["pd_sum"]) <= 1e-9 * abs(b["pd_sum"]) + 1e-12
    assert abs(a["pd_last"] - b["pd_last"]) <= 1e-12 * abs(b["pd_last"]) + 1e-300


@pytest.fixture(scope="module")
def cli():
    src = os.path.join(ROOT, "tests", "cpp", "icp_gpu_cli.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "icp_gpu_cli")
    pk = os.path.join(ROOT, "slam-envire_b200")
    deps = [src, os.path.join(pk, "icp", "icp_gpu.hpp"), os.path.join(pk, "compat", "Eigen", "mini_eigen.hpp"),
            os.path.join(pk, "libicp_b200.so")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + pk, "-I" + os.path.join(pk, "compat"),
                               "-o", exe, src, "-L" + pk, "-licp_b200", "-Wl,-rpath," + pk])
    return exe


@pytest.mark.gpu
def test_one_process_entry_points_on_one_device(cli, tmp_path):
    """TrimmedGPU(0, 1): the several-GPUs-in-one-process entry points (icpb_create_multi ...) with a single device must
    give exactly what the plain context gives."""
    synth = importlib.import_module("slam-envire_b200.synth")
    m, s, T, prm = synth.config("cfg1", 0.2)
    prm = dict(prm, max_iter=20)
    a = _cli_run(cli, tmp_path, m, s, "fixed", prm, None, density=0.5)
    b = _cli_run(cli, tmp_path, m, s, "fixed", prm, 1, density=0.5)
    assert a == b


@pytest.mark.gpu
def test_one_process_two_gpus_matches_single_gpu(cli, tmp_path):
    """envire::icp::TrimmedGPU(0, 2) -- two GPUs, ONE process, peers mapped with cudaDeviceEnablePeerAccess, the same
    fused exchanges -- against TrimmedGPU on one GPU: a cfg3-shaped case (half of the source replaced, overlap 0.5), a
    density-stepped source, and the golden-section overlap search (several aligns on one upload)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    synth = importlib.import_module("slam-envire_b200.synth")
    m = synth.scene(200_000, 60.0, 4)
    s, T = synth.perturbed_copy(m, 1.0, 0.3, 0.01, seed=5, replace_frac=0.5, extent=60.0)
    prm = dict(max_iter=25, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.5)
    _same_align(_cli_run(cli, tmp_path, m, s, "fixed", prm, 2), _cli_run(cli, tmp_path, m, s, "fixed", prm, None), 60.0)
    _same_align(_cli_run(cli, tmp_path, m, s, "fixed", prm, 2, density=0.37), _cli_run(cli, tmp_path, m, s, "fixed", prm, None, density=0.37), 60.0)
    m1, s1, T1, p1 = synth.config("cfg1", 0.2)
    g = dict(max_iter=15, min_mse=1e-10, min_mse_diff=1e-12, alpha=0.5, beta=1.0, eps=0.05)
    a, b = _cli_run(cli, tmp_path, m1, s1, "golden", g, 2), _cli_run(cli, tmp_path, m1, s1, "golden", g, None)
    assert abs(a["overlap"] - b["overlap"]) < 1e-12
    _same_align(a, b, 20.0)


def test_shard_bounds_and_density():
    sh = importlib.import_module("slam-envire_b200.sharding")
    from oracle import oracle as O
    for n, dens in [(1000, 1.0), (1000, 0.37), (17, 0.5), (5, 1.0)]:
        sel = sh.density_indices(n, dens)
        assert np.array_equal(sel, O.density_indices(n, dens).astype(np.int64))
        assert sh.adapter_size(n, dens) == O.adapter_size(n, dens)
        for world in (1, 2, 3, 8):
            parts = [sh.shard_bounds(len(sel), world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == len(sel)
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
    xyz = np.arange(30, dtype=np.float64).reshape(10, 3)
    loc, size = sh.shard_source(xyz, 0.5, 2, 1)
    assert size == 5 and np.array_equal(loc, xyz[[6, 8]])


D_SHIFT, D_BITS = [52, 41, 30, 19, 8, 0], [11, 11, 11, 11, 11, 8]  # K4 digit schedule (kernels.cuh)


def _worker(rank, world, port, model, src, prm, density, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    from conftest import Harness
    sh = importlib.import_module("slam-envire_b200.sharding")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = Harness()
    M = O.Model()
    M.add(model)
    local, size = sh.shard_source(src, density, world, rank)
    n_po = int(size * prm["overlap"])
    C = np.eye(4)
    it, mse, mse_diff, d_box, pairs = 0, np.inf, np.inf, np.inf, 0

    def allsum(a, dtype):
        t = torch.tensor(np.asarray(a), dtype=dtype)
        dist.all_reduce(t)
        return t.numpy()

    while it < prm["max_iter"] and mse > prm["min_mse"] and mse_diff > prm["min_mse_diff"]:
        p = O.transform_points(C, local)
        idx, d = M.find_pairs(p, d_box=d_box)
        d = np.where(idx == O.NO_PAIR, np.inf, d)
        found = int(allsum([np.isfinite(d).sum()], torch.int64)[0])
        k = min(n_po, found)
        keys = d.view(np.uint64)
        # radix select over all-reduced histograms
        prefix, krem = np.uint64(0), k
        tau, kept_mask = np.nan, np.zeros(len(d), bool)
        if k > 0:
            for s_, b_ in zip(D_SHIFT, D_BITS):
                hs = np.uint64(s_ + b_)
                match = np.ones(len(d), bool) if hs >= 64 else ((keys >> hs) == (prefix >> hs))
                dig = ((keys[match] >> np.uint64(s_)) & np.uint64((1 << b_) - 1)).astype(np.int64)
                hist = allsum(np.bincount(dig, minlength=1 << b_), torch.int64)
                cum = np.cumsum(hist)
                b = int(np.searchsorted(cum, krem))
                krem -= int(cum[b] - hist[b])
                prefix |= np.uint64(b) << np.uint64(s_)
            tau = np.array([prefix], dtype=np.uint64).view(np.float64)[0]
            quota = krem  # ties to keep globally, in global pair order: lower ranks first
            ties = np.flatnonzero(d == tau)
            counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(counts, torch.tensor([len(ties)], dtype=torch.int64))
            before = int(sum(c.item() for c in counts[:rank]))
            mine = max(0, min(len(ties), quota - before))
            kept_mask = d < tau
            kept_mask[ties[:mine]] = True
        d_box = tau * 2.0
        pairs = k
        if pairs < 3:
            break
        P, X, D = p[kept_mask], M.points()[idx[kept_mask]], d[kept_mask]
        sums = np.concatenate([P.sum(0), X.sum(0), (P[:, :, None] * X[:, None, :]).sum(0).reshape(9),
                               [(D * D).sum(), float(len(D))]])
        sums = allsum(sums, torch.float64)
        T, new_mse = H.transform_from_sums(sums)  # the product's closed-form pose step (pose.h)
        C = T @ C
        mse_diff, mse = mse - new_mse, new_mse
        it += 1
    if rank == 0:
        q.put(dict(C=C, iter=it, pairs=pairs, mse=mse, d_box=d_box))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["scene", "ties"])
def test_sharded_protocol_gloo_world2(oracle, synth, case):
    import torch.multiprocessing as mp
    if case == "scene":
        m, s, T, prm = synth.config("cfg1", 0.04)
        prm, density = dict(prm, max_iter=12), 0.6
    else:
        g = np.arange(0, 16, dtype=np.float64)
        m = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
        m = np.concatenate([m, np.zeros((len(m), 1))], 1)
        s = m + [0.3, 0.2, 0.0]
        prm, density = dict(max_iter=4, min_mse=1e-12, min_mse_diff=1e-14, overlap=0.55), 1.0
    M = oracle.Model()
    M.add(m)
    ref = M.align(s, density=density, **prm)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, s, prm, density, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["iter"] == ref.iter and out["pairs"] == ref.pairs
    assert np.abs(out["C"] - ref.matrix()).max() < 1e-9
    assert abs(out["mse"] - ref.mse) <= 1e-9 * abs(ref.mse) + 1e-20
