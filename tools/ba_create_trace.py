"""Where the host time of pmv_ba_problem_create goes (PMV_BA_TRACE=1) for BASELINE configs 4 and 5."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["PMV_BA_TRACE"] = "1"
import numpy as np
import pmv_b200
from harness import synth
ctx = pmv_b200.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "5"
if which == "5":
    w = synth.ba_large(7)
    for rep in range(2):
        t0 = time.perf_counter()
        prob = ctx.ba_problem(w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"], 1.0)
        t1 = time.perf_counter(); prob.solve(5); P, X, S = prob.download(); t2 = time.perf_counter()
        prob.close(); t3 = time.perf_counter()
        print(f"config 5: create {1e3*(t1-t0):.1f} ms, solve+download {1e3*(t2-t1):.1f} ms, close {1e3*(t3-t2):.1f} ms", file=sys.stderr)
else:
    W = 4096
    ws = [synth.ba_window(i) for i in range(16)]
    sel = [ws[i % 16] for i in range(W)]
    off = np.cumsum([0] + [len(x["obs"]) for x in sel]).astype(np.int32)
    poses = np.stack([x["poses"] for x in sel]); points = np.stack([x["points"] for x in sel])
    obs = np.concatenate([x["obs"] for x in sel]); cam = np.concatenate([x["cam_idx"] for x in sel]); pt = np.concatenate([x["pt_idx"] for x in sel])
    for rep in range(2):
        t0 = time.perf_counter()
        prob = ctx.ba_problem(poses, points, obs, cam, pt, ws[0]["K"], 1.0, obs_off=off)
        t1 = time.perf_counter(); prob.solve(5); P, X, S = prob.download(); t2 = time.perf_counter()
        prob.close()
        print(f"config 4: create {1e3*(t1-t0):.1f} ms, solve+download {1e3*(t2-t1):.1f} ms", file=sys.stderr)
