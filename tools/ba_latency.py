"""Latency probe of the one-shot bundle-adjustment call (config 1 shape: 5 poses x 400 points x 5 iterations).

Usage: python tools/ba_latency.py            (needs a GPU)
Prints wall-clock ms per pmv_ba_solve call and the device time of the BA phase, for the window path and the
general path (PMV_BA_FORCE_GENERAL=1), so the host overhead (sort, upload, download) is visible.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmv_b200 as pmv  # noqa: E402
from harness import synth  # noqa: E402


def probe(label, n_poses=5, n_points=400, iters=5, reps=200):
    ctx = pmv.Context(0)
    w = synth.ba_window(3, n_poses=n_poses, n_points=n_points)
    args = (w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"])
    for _ in range(10):
        ctx.ba_solve(*args, max_iters=iters)
    ctx.profile(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.ba_solve(*args, max_iters=iters)
    wall = (time.perf_counter() - t0) / reps * 1e3
    prof = ctx.profile_collect()
    dev = prof.get("ba", (0, 0))
    print(f"{label}: wall {wall:.3f} ms/call, device BA phase {dev[0] / max(dev[1], 1):.3f} ms/call "
          f"({len(w['cam_idx'])} observations)")


if __name__ == "__main__":
    probe("default      ")
    os.environ["PMV_BA_FORCE_WINDOW"] = "1"
    probe("window path ")
    del os.environ["PMV_BA_FORCE_WINDOW"]
    os.environ["PMV_BA_FORCE_GENERAL"] = "1"
    probe("general path")
    os.environ["PMV_BA_NO_GRAPH"] = "1"
    probe("general path, no graph")
