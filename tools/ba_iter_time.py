"""ms per LM iteration of one BAL-scale problem (BASELINE config 5 shape by default) -- tuning aid for the run kernels."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pmv_b200
from harness import synth

nposes, npts, span = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 1000000, 40)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
w = synth.ba_large(31, n_poses=nposes, n_points=npts, views=5, span=span)
ctx = pmv_b200.Context(0)
prob = ctx.ba_problem(w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"], 1.0)
prob.solve(iters)
st = prob.download()[2][0]
ts = []
for _ in range(4):
    prob.reset(); ctx.sync(); t0 = time.perf_counter(); prob.solve(iters); ctx.sync(); ts.append(time.perf_counter() - t0)
print({"ms_per_iter": min(ts) / max(st["iterations"], 1) * 1e3, "iterations": st["iterations"], "successful_steps": st["successful_steps"],
       "final_cost": repr(st["final_cost"])}, flush=True)
