"""One banded BA problem solved with the partitioned solve (PMV_CHOL_PARTS from the environment): short, for ncu."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pmv_b200
from harness import synth
nposes, npts, span = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 200000, 40)
w = synth.ba_large(31, n_poses=nposes, n_points=npts, views=5, span=span)
ctx = pmv_b200.Context(0)
prob = ctx.ba_problem(w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"], 1.0)
prob.solve(2)
print(prob.download()[2][0])
