"""Latency of pmv_five_point_pose (one launch) beside cv2.findEssentialMat + cv2.recoverPose on the host, on synthetic
two-view scenes shaped like the pipeline's initialisation (400 integer-pixel correspondences, 25 % outliers)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cv2  # noqa: E402

import pmv_b200  # noqa: E402
from harness import twoview_scene  # noqa: E402

ctx = pmv_b200.Context(0)
scenes = [twoview_scene.scene(9000 + s, n=400) for s in range(40)]
for sc in scenes[:5]:
    ctx.five_point_pose(sc["p1"], sc["p2"], sc["K"])
t0 = time.perf_counter()
samples = 0
for sc in scenes:
    r = ctx.five_point_pose(sc["p1"], sc["p2"], sc["K"])
gpu = (time.perf_counter() - t0) / len(scenes)
t0 = time.perf_counter()
same = 0
for sc in scenes:
    E, m = cv2.findEssentialMat(sc["p1"], sc["p2"], sc["K"], cv2.RANSAC, 0.99, 1.0)
    cv2.recoverPose(E, sc["p1"], sc["p2"], sc["K"], distanceThresh=float("inf"), mask=m.copy())
cpu = (time.perf_counter() - t0) / len(scenes)
for sc in scenes:
    E, m = cv2.findEssentialMat(sc["p1"], sc["p2"], sc["K"], cv2.RANSAC, 0.99, 1.0)
    r = ctx.five_point_pose(sc["p1"], sc["p2"], sc["K"])
    same += np.array_equal(r["ransac_mask"], m.ravel())
print(json.dumps({"workload": "five-point RANSAC + recoverPose, 400 correspondences, 25% outliers", "gpu_ms_per_call": gpu * 1e3,
                  "cv2_ms_per_call": cpu * 1e3, "identical_masks": int(same), "scenes": len(scenes)}))
