"""Host-side partitioning of the hot path over the GPUs of one box.

* front end (pyramid / LK / extractors): independent streams -> one per rank, no collective;
* windowed BA (config 4): independent windows -> contiguous window ranges per rank, no collective;
* large BA (config 5): points (and their observations) sharded contiguous-by-index, poses replicated;
  the per-iteration exchange is the NCCL all-reduce inside pmv_ba_problem_solve.
Pure numpy: covered on CPU by world_size-2 gloo tests (tests/test_sharding_gloo.py)."""
from __future__ import annotations

import numpy as np


def split_range(n: int, rank: int, nranks: int):
    """Contiguous [lo, hi) of `n` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, nranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_points(points, obs, cam_idx, pt_idx, rank: int, nranks: int):
    """Local view of a bundle-adjustment problem for `rank`: its contiguous block of points, the
    observations of those points (original relative order kept) with point indices renumbered from 0.
    Returns (points_local, obs_local, cam_local, pt_local, (lo, hi), obs_selector)."""
    points = np.asarray(points)
    pt_idx = np.asarray(pt_idx)
    lo, hi = split_range(len(points), rank, nranks)
    sel = np.nonzero((pt_idx >= lo) & (pt_idx < hi))[0]
    return (np.ascontiguousarray(points[lo:hi]), np.ascontiguousarray(np.asarray(obs)[sel]),
            np.ascontiguousarray(np.asarray(cam_idx)[sel]).astype(np.int32),
            (pt_idx[sel] - lo).astype(np.int32), (lo, hi), sel)


def shard_windows(n_windows: int, rank: int, nranks: int):
    return split_range(n_windows, rank, nranks)
