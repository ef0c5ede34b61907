"""Python re-drive of the reference's front-end control flow (OdometryPipeline::addFrame,
OdometryPipeline.cpp:329-374; initialise :428-482; getGridROI :674-692) with a pluggable backend, so the
GPU extractor + matcher can be compared with OpenCV on the *pipeline's own call pattern* (BASELINE
config 1) in a container that cannot build the C++ executable.  Harness code (tests / bench), not product;
the OpenCV-plugin backend it is compared with lives in oracle/replay_backend.py (test infrastructure).

Per frame:  LK(prev, frame, features of prev; win 32x32, maxLevel 4)  ->  keep status==1, truncate to int
            if fewer than `tol` (150) survive: split the PREVIOUS frame into 255x255 ROIs, extract
            ceil(min_tracked/n_roi) = 40 corners per ROI, drop those with a Chebyshev neighbour < 5 among the
            new frame's features (tested in ROI-local coordinates, like the reference), add ROI offset.
"""
from __future__ import annotations

import math

import numpy as np

GRID = 255
WIN = (32, 32)
MAX_LEVEL = 4


def grid_rois(rows: int, cols: int):
    """OdometryPipeline::getGridROI: (x, y, w, h) tiles in raster order."""
    out = []
    for r in range(0, rows, GRID):
        for c in range(0, cols, GRID):
            out.append((c, r, min(GRID, cols - c), min(GRID, rows - r)))
    return out


class GpuBackend:
    """GpuGoodFeatureExtractor + GpuLucasKanadeFM through the C ABI."""
    name = "pmv"

    def __init__(self, ctx):
        self.ctx = ctx

    def extract(self, img, roi, max_feats):
        xy, _ = self.ctx.gftt(img, max_feats, 0.01, 5.0, roi=roi)
        return xy.astype(np.int32)

    def track(self, prev, nxt, pts):
        if len(pts) == 0:
            return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8)
        nx, st, _ = self.ctx.lk_track(prev, nxt, pts.astype(np.float32), WIN, MAX_LEVEL)
        return nx, st


def has_neighbor(f, feats, dist=5):
    """Frame::hasNeighbor (Frame.cpp:3-12): Chebyshev distance < dist to any existing feature."""
    if len(feats) == 0:
        return False
    d = np.abs(np.asarray(feats) - np.asarray(f)).max(axis=1)
    return bool((d < dist).any())


def run_front_end(frames, backend, min_tracked=400, tol=150):
    """Returns per frame: (features (n,2) int32 [column,row], n_tracked, extracted_flag)."""
    rows, cols = frames[0].shape
    rois = grid_rois(rows, cols)
    n_grid = int(math.ceil(min_tracked / len(rois)))
    # initialise(): features of frame 0 from every ROI (OdometryPipeline.cpp:447-459)
    feats = []
    for roi in rois:
        for (x, y) in backend.extract(frames[0], roi, int(min_tracked / len(rois))):
            feats.append((roi[0] + x, roi[1] + y))
    feats = np.array(feats, np.int32).reshape(-1, 2)
    log = [(feats.copy(), len(feats), True)]
    for k in range(1, len(frames)):
        prev, cur = frames[k - 1], frames[k]
        nx, st = backend.track(prev, cur, feats)
        new = nx[st == 1].astype(np.int32)          # Feature(int, int): truncation (OpenCVLucasKanadeFM.cpp:25)
        tracked = len(new)
        extracted = False
        if tracked < tol:
            extracted = True
            cur_feats = [tuple(p) for p in new]
            for roi in rois:                        # ROIs of the PREVIOUS frame (OdometryPipeline.cpp:351)
                for (x, y) in backend.extract(prev, roi, n_grid):
                    if not has_neighbor((x, y), cur_feats):           # ROI-local test, as in the reference (:361)
                        cur_feats.append((roi[0] + x, roi[1] + y))
            new = np.array(cur_feats, np.int32).reshape(-1, 2)
        feats = new
        log.append((feats.copy(), tracked, extracted))
    return log


def synthetic_sequence(n_frames: int, stream: int = 0, h: int = 376, w: int = 1241):
    from . import synth
    frames = [synth.base_frame(1000 * stream, h, w)]
    for k in range(1, n_frames):
        frames.append(synth.next_frame(frames[-1], 1000 * stream + k))
    return frames
