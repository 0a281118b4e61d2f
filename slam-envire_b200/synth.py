"""Synthetic scene clouds for the BASELINE.json configs.

Pure numpy, seeded, host side only: fp64 AoS double[N][3] exactly like
envire::Pointcloud::vertices (src/maps/Pointcloud.hpp:35 of the reference).
"""
import numpy as np


def rotation(axis, angle_rad):
    """Rodrigues rotation matrix about `axis` by `angle_rad`."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle_rad) * K + (1 - np.cos(angle_rad)) * (K @ K)


def rigid(axis, angle_rad, t):
    T = np.eye(4)
    T[:3, :3] = rotation(axis, angle_rad)
    T[:3, 3] = t
    return T


def scene(n, extent, seed):
    """50 % undulating ground, 25 % wall at y=+E/2, 25 % wall at x=-E/2, uniform in-surface."""
    rng = np.random.default_rng(seed)
    half = extent / 2.0
    height = min(extent / 4.0, 10.0)
    n_g = n // 2
    n_w1 = (n - n_g) // 2
    n_w2 = n - n_g - n_w1
    pts = np.empty((n, 3), dtype=np.float64)
    x = rng.uniform(-half, half, n_g)
    y = rng.uniform(-half, half, n_g)
    pts[:n_g] = np.stack([x, y, 0.3 * np.sin(0.5 * x) * np.cos(0.4 * y)], axis=1)
    x = rng.uniform(-half, half, n_w1)
    z = rng.uniform(0.0, height, n_w1)
    pts[n_g:n_g + n_w1] = np.stack([x, np.full(n_w1, half), z], axis=1)
    y = rng.uniform(-half, half, n_w2)
    z = rng.uniform(0.0, height, n_w2)
    pts[n_g + n_w1:] = np.stack([np.full(n_w2, -half), y, z], axis=1)
    return pts[rng.permutation(n)]


def perturbed_copy(model, angle_deg, trans_norm, noise_sigma=0.0, seed=0, axis=(0.2, 0.1, 1.0),
                   replace_frac=0.0, extent=None):
    """source = R^-1 (model - t) (+ noise); optionally a fraction replaced by points of a disjoint region.

    Returns (source, T_true) with T_true the rigid transform that maps the source back onto the model.
    """
    rng = np.random.default_rng(seed)
    tdir = np.array([1.0, -0.5, 0.25])
    t = trans_norm * tdir / np.linalg.norm(tdir)
    R = rotation(axis, np.deg2rad(angle_deg))
    src = (model - t) @ R  # row-vector form of R^T (model - t)
    if noise_sigma > 0:
        src = src + rng.normal(0.0, noise_sigma, src.shape)
    if replace_frac > 0:
        n = len(src)
        k = int(n * replace_frac)
        sel = rng.choice(n, k, replace=False)
        ext = extent if extent is not None else float(np.ptp(model[:, 0]))
        other = scene(k, ext, seed + 1000)
        other[:, 0] += 2.0 * ext  # disjoint region
        src[sel] = other
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return np.ascontiguousarray(src), T


def config(name, scale=1.0):
    """(model, source, T_true, params) for a BASELINE.json configuration; `scale` shrinks the point counts."""
    if name == "cfg1":
        n, ext = int(50_000 * scale), 20.0
        m = scene(n, ext, 1)
        s, T = perturbed_copy(m, 10.0, 0.5, 0.0, seed=11)
        return m, s, T, dict(max_iter=100, min_mse=1e-10, min_mse_diff=1e-12, overlap=0.9)
    if name == "cfg2":
        n, ext = int(1_000_000 * scale), 100.0
        m = scene(n, ext, 2)
        s, T = perturbed_copy(m, 3.0, 0.3, 0.01, seed=3)
        return m, s, T, dict(max_iter=50, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.8)
    if name == "cfg3":
        n, ext = int(10_000_000 * scale), 300.0
        m = scene(n, ext, 4)
        s, T = perturbed_copy(m, 1.0, 0.3, 0.01, seed=5, replace_frac=0.5, extent=ext)
        return m, s, T, dict(max_iter=30, min_mse=1e-10, min_mse_diff=1e-9, overlap=0.5)
    raise ValueError(name)
