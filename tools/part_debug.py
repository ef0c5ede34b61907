"""Partitioned / two-sided / one-sided solves of one banded BA problem side by side (debug + timing aid)."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import pmv_b200
from harness import synth

nposes, npts, span = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (400, 20000, 20)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
w = synth.ba_large(31, n_poses=nposes, n_points=npts, views=5, span=span)
args = (w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"])
ctx = pmv_b200.Context(0)
ref = None
for mode, env in (("one_sided", {"PMV_CHOL_NO_SPLIT": "1"}), ("split", {}), ("part3", {"PMV_CHOL_PARTS": "3"}),
                  ("part4", {"PMV_CHOL_PARTS": "4"}), ("part6", {"PMV_CHOL_PARTS": "6"}), ("part8", {"PMV_CHOL_PARTS": "8"})):
    for k in ("PMV_CHOL_NO_SPLIT", "PMV_CHOL_PARTS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    prob = ctx.ba_problem(*args, 1.0)
    prob.solve(iters)
    p, x, s = prob.download()
    ts = []
    for _ in range(3):
        prob.reset(); ctx.sync(); t0 = time.perf_counter(); prob.solve(iters); ctx.sync(); ts.append(time.perf_counter() - t0)
    prob.close()
    if ref is None:
        ref = (p, s[0])
    print(mode, "iters", s[0]["iterations"], "ok steps", s[0]["successful_steps"], "cost", repr(s[0]["final_cost"]), "dpose vs one-sided",
          float(np.abs(p[0] - ref[0][0]).max()), "ms/iter", min(ts) / max(s[0]["iterations"], 1) * 1e3, flush=True)
