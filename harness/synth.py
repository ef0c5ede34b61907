"""Synthetic KITTI-shaped inputs (SURVEY.md §8d generator) -- harness data, not product code.

Value-noise texture (thousands of corners, ~50 % of pixels >= 128) and small affine
frame-to-frame motion; ``seed = 1000 * stream + frame`` as the survey fixes it.
"""
from __future__ import annotations

import numpy as np


def base_frame(seed: int, h: int = 376, w: int = 1241) -> np.ndarray:
    import cv2
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 6 + 2, w // 6 + 2), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    return cv2.GaussianBlur(img, (0, 0), 1.2)


def next_frame(img: np.ndarray, seed: int) -> np.ndarray:
    import cv2
    rng = np.random.default_rng(seed)
    tx, ty = rng.uniform(-4, 4, 2)
    a, b = rng.uniform(-0.003, 0.003, 2)
    m = np.array([[1 + a, b, tx], [-b, 1 - a, ty]], np.float64)
    h, w = img.shape
    out = cv2.warpAffine(img, m, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    noise = rng.integers(-3, 4, out.shape, dtype=np.int16)
    return np.clip(out.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def frame_pair(stream: int, frame: int = 0, h: int = 376, w: int = 1241):
    f0 = base_frame(1000 * stream + frame, h, w)
    f1 = next_frame(f0, 1000 * stream + frame + 1)
    return f0, f1


def track_points(img: np.ndarray, n: int, seed: int, jitter: float = 0.25, min_dist: float = 3.0):
    """Config-2 points: goodFeaturesToTrack corners (integer valued) + sub-pixel jitter."""
    import cv2
    pts = cv2.goodFeaturesToTrack(img, n, 0.01, min_dist)
    pts = pts.reshape(-1, 2).astype(np.float32)
    rng = np.random.default_rng(seed)
    if len(pts) < n:  # top up with uniform points so every pair carries exactly n tracks
        extra = rng.uniform([0, 0], [img.shape[1] - 1, img.shape[0] - 1], (n - len(pts), 2))
        pts = np.concatenate([pts, extra.astype(np.float32)])
    pts = pts + rng.uniform(-jitter, jitter, pts.shape).astype(np.float32)
    return np.ascontiguousarray(pts, np.float32)


# ----------------------------------------------------------------------------- bundle adjustment
KITTI_K = np.array([[718.856, 0.0, 607.1928], [0.0, 718.856, 185.2157], [0.0, 0.0, 1.0]])


def _rodrigues(a):
    th = np.linalg.norm(a)
    if th < 1e-12:
        return np.eye(3)
    w = a / th
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def project(pose, X, K=KITTI_K):
    """ProjectionResidual's camera model (ProjectionResidual.h:44-52): p = R(a)(X+c), z flipped."""
    p = (_rodrigues(pose[:3]) @ (X + pose[3:]).T).T
    z = -p[:, 2]
    return np.stack([p[:, 0] / z * K[0, 0] + K[0, 2], p[:, 1] / z * K[1, 1] + K[1, 2]], -1), p[:, 2]


def ba_window(seed: int, n_poses=20, n_points=2000, width=1241, height=376, outlier_frac=0.05, K=KITTI_K):
    """BASELINE config 4 window: forward trajectory 1 m/frame along -z with yaw noise,
    points in the frustum at depth 5-50 m, integer-rounded observations with N(0, .5 px) noise and
    5 % gross outliers, initial parameters = truth + noise.  Returns a dict of numpy arrays."""
    rng = np.random.default_rng(seed)
    poses_t = np.zeros((n_poses, 6))
    for i in range(n_poses):
        poses_t[i, :3] = [0, rng.normal(0, 0.01), 0]
        poses_t[i, 3:] = [0, 0, 1.0 * i] + rng.normal(0, 0.02, 3)          # c = -t ; camera moves along -z
    depth = rng.uniform(5, 50, n_points) + n_poses * 0.5
    u = rng.uniform(0, width, n_points); v = rng.uniform(0, height, n_points)
    # back-project through the middle camera so most points are seen by most poses
    mid = poses_t[n_poses // 2]
    pc = np.stack([(u - K[0, 2]) / K[0, 0] * depth, (v - K[1, 2]) / K[1, 1] * depth, -depth], -1)
    pts_t = (_rodrigues(mid[:3]).T @ pc.T).T - mid[3:]
    cam_idx, pt_idx, obs = [], [], []
    for i in range(n_poses):
        uv, z = project(poses_t[i], pts_t, K)
        vis = (z < -0.5) & (uv[:, 0] >= 0) & (uv[:, 0] < width) & (uv[:, 1] >= 0) & (uv[:, 1] < height)
        ids = np.nonzero(vis)[0]
        o = uv[ids] + rng.normal(0, 0.5, (len(ids), 2))
        bad = rng.random(len(ids)) < outlier_frac
        o[bad] += rng.uniform(5, 30, (bad.sum(), 2)) * rng.choice([-1, 1], (bad.sum(), 2))
        cam_idx.append(np.full(len(ids), i)); pt_idx.append(ids); obs.append(np.round(o))
    cam_idx = np.concatenate(cam_idx).astype(np.int32); pt_idx = np.concatenate(pt_idx).astype(np.int32)
    obs = np.concatenate(obs).astype(np.float64)
    poses0 = poses_t + np.concatenate([rng.normal(0, 0.005, (n_poses, 3)), rng.normal(0, 0.05, (n_poses, 3))], 1)
    pts0 = (pts_t + rng.normal(0, 0.1, pts_t.shape)).astype(np.float32).astype(np.float64)  # Feature3D stores float
    return {"poses": poses0, "points": pts0, "obs": obs, "cam_idx": cam_idx, "pt_idx": pt_idx, "K": K.copy(),
            "poses_true": poses_t, "points_true": pts_t}


def ba_large(seed: int, n_poses=1000, n_points=1_000_000, views=5, span=40, width=1241, height=376, K=KITTI_K,
             sort_by_anchor=False):
    """BASELINE config 5 (BAL scale): each point is seen by `views` poses chosen among the `span` (40)
    poses nearest to it along a straight trajectory -- the `span` poses up to and including its anchor, which all
    have the point in front of them; observations need not fall inside an image.  sort_by_anchor orders the
    points along the trajectory (as real SfM / BAL point lists are), so a contiguous point shard sees only a
    contiguous run of cameras."""
    rng = np.random.default_rng(seed)
    poses_t = np.zeros((n_poses, 6))
    poses_t[:, 1] = rng.normal(0, 0.01, n_poses)
    poses_t[:, 5] = np.arange(n_poses) * 1.0
    poses_t[:, 3:] += rng.normal(0, 0.02, (n_poses, 3))
    anchor = rng.integers(0, n_poses, n_points)
    if sort_by_anchor:
        anchor = np.sort(anchor)
    depth = rng.uniform(8, 50, n_points)
    u = rng.uniform(0.2 * width, 0.8 * width, n_points); v = rng.uniform(0.2 * height, 0.8 * height, n_points)
    pc = np.stack([(u - K[0, 2]) / K[0, 0] * depth, (v - K[1, 2]) / K[1, 1] * depth, -depth], -1)
    # anchor rotation is a tiny yaw: back-project with the exact model point by point in blocks
    pts_t = np.empty_like(pc)
    for a in np.unique(anchor):
        m = anchor == a
        pts_t[m] = (_rodrigues(poses_t[a, :3]).T @ pc[m].T).T - poses_t[a, 3:]
    # `views` distinct poses within (anchor - span, anchor]: they all stay in front of the point (z < 0)
    offs = np.stack([rng.permutation(span)[:views] for _ in range(1)], 0)  # same pattern base, shifted per point
    shift = rng.integers(0, span, n_points)
    cams = (anchor[:, None] - ((offs + shift[:, None]) % span)).clip(0, n_poses - 1)
    cams = np.sort(cams, 1)
    keep = np.ones_like(cams, bool)
    keep[:, 1:] = cams[:, 1:] != cams[:, :-1]
    pt_idx = np.repeat(np.arange(n_points), views).reshape(n_points, views)[keep].astype(np.int32)
    cam_idx = cams[keep].astype(np.int32)
    obs = np.empty((len(cam_idx), 2))
    order = np.argsort(cam_idx, kind="stable")
    bounds = np.searchsorted(cam_idx[order], np.arange(n_poses + 1))
    for c in range(n_poses):
        sel = order[bounds[c]:bounds[c + 1]]
        if len(sel):
            uv, _ = project(poses_t[c], pts_t[pt_idx[sel]], K)
            obs[sel] = np.round(uv + rng.normal(0, 0.5, uv.shape))
    poses0 = poses_t + np.concatenate([rng.normal(0, 0.002, (n_poses, 3)), rng.normal(0, 0.03, (n_poses, 3))], 1)
    pts0 = (pts_t + rng.normal(0, 0.05, pts_t.shape)).astype(np.float32).astype(np.float64)
    return {"poses": poses0, "points": pts0, "obs": obs, "cam_idx": cam_idx, "pt_idx": pt_idx, "K": K.copy(),
            "poses_true": poses_t, "points_true": pts_t}
