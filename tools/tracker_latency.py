"""Latency probe of the resident front end (pmv_tracker_add_frame) against the call-by-call pmv_lk_track path."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import pmv_b200
from harness import replay

ctx = pmv_b200.Context(0)
frames = replay.synthetic_sequence(60, stream=1)
tr = ctx.tracker(*frames[0].shape)
for rep in range(2):
    f0 = tr.init(frames[0])
    t0 = time.perf_counter()
    for k in range(1, len(frames)):
        tr.add_frame(frames[k])
    dt = time.perf_counter() - t0
print(f"tracker add_frame: {1e3 * dt / (len(frames) - 1):.3f} ms/frame, {len(f0)} initial features")
ctx.profile(True); ctx.profile_collect()
tr.init(frames[0])
for k in range(1, len(frames)):
    tr.add_frame(frames[k])
print("phases (ms total, count):", ctx.profile_collect())
pts = f0.astype(np.float32)
for rep in range(2):
    t0 = time.perf_counter()
    for k in range(1, len(frames)):
        ctx.lk_track(frames[k - 1], frames[k], pts, (32, 32), 4)
    dt = time.perf_counter() - t0
print(f"call-by-call pmv_lk_track: {1e3 * dt / (len(frames) - 1):.3f} ms/frame")
print("phases:", ctx.profile_collect())
