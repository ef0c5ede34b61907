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
