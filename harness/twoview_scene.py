"""Synthetic two-view correspondences for the five-point initialiser (harness code): a KITTI-like forward motion, points
in front of both cameras, integer pixel coordinates (Feature coordinates are ints, OpenCVFivePointTri.cpp:8 builds
std::vector<cv::Point>), Gaussian noise and a share of gross outliers."""
import numpy as np

from .pnp_scene import _rodrigues
from .synth import KITTI_K


def scene(seed, n=None, outlier_share=0.25, noise=0.5, integer=True, K=KITTI_K):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(60, 600)) if n is None else n
    K = np.asarray(K, np.float64)
    X = np.c_[rng.uniform(-15, 15, n), rng.uniform(-4, 4, n), rng.uniform(6, 40, n)]
    rv = rng.normal(0, 0.03, 3)
    R = _rodrigues(rv)
    t = np.array([0.1, -0.05, 1.0]) + rng.normal(0, 0.05, 3)

    def proj(P):
        x = P @ K.T
        return x[:, :2] / x[:, 2:]

    p1 = proj(X) + rng.normal(0, noise, (n, 2))
    p2 = proj(X @ R.T + t) + rng.normal(0, noise, (n, 2))
    k = int(outlier_share * n)
    out = np.sort(rng.choice(n, k, replace=False)) if k else np.zeros(0, np.int64)
    p2[out] += rng.uniform(20, 60, (k, 2)) * rng.choice([-1, 1], (k, 2))
    if integer:
        p1 = np.trunc(p1); p2 = np.trunc(p2)
    return {"p1": np.ascontiguousarray(p1, np.float64), "p2": np.ascontiguousarray(p2, np.float64), "K": K, "R": R,
            "t": t / np.linalg.norm(t), "outliers": out, "X": X}
