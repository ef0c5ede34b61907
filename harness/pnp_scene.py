"""Synthetic 3-D / 2-D correspondences for the pose solver (harness code): points in front of the camera (OpenCV
convention, z > 0 -- the reference flips its z before the call, OpenCVEPnPSolver.cpp:26), integer pixel observations
(Feature coordinates are ints) with Gaussian noise and a share of gross outliers, and the previous pose as the guess."""
import numpy as np

from .synth import KITTI_K


def _rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def scene(seed, n=None, outlier_share=1 / 15, noise=0.5, K=KITTI_K):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 500)) if n is None else n
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-5, 5, n), rng.uniform(8, 50, n)], 1).astype(np.float32)
    rv = rng.normal(0, 0.02, 3); tv = np.array([0.1, -0.05, 1.0]) + rng.normal(0, 0.1, 3)
    p = (_rodrigues(rv) @ X.T.astype(np.float64)).T + tv
    uv = (K @ p.T).T
    uv = uv[:, :2] / uv[:, 2:]
    uv = uv + rng.normal(0, noise, uv.shape)
    no = max(1, int(n * outlier_share))
    out = rng.choice(n, no, replace=False)
    uv[out] += rng.uniform(20, 60, (no, 2)) * rng.choice([-1, 1], (no, 2))
    uv = np.ascontiguousarray(np.round(uv).astype(np.float32))
    guess_r = rv + rng.normal(0, 0.01, 3); guess_t = tv + rng.normal(0, 0.2, 3)
    return {"X": np.ascontiguousarray(X), "uv": uv, "K": np.asarray(K, np.float64), "rvec": rv, "tvec": tv,
            "guess_r": guess_r, "guess_t": guess_t, "outliers": np.sort(out)}
