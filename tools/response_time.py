"""CUDA-event timing of the batched corner-response entry points on 8 resident 4K frames (the bench's extract leg reports
the same): algorithmic bytes 1 B/px in + 4 B/px (min-eig) or 8 B/px (fp64 ShiTomasi) out, against the measured HBM peak.
Also checks the batched maps against cv2.cornerMinEigenVal / the oracle on one frame."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import pmv_b200
from harness import synth

H, W, B = 2160, 3840, 8
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
ctx = pmv_b200.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
pitch = (W + 15) // 16 * 16
d = torch.zeros(B, H, pitch, dtype=torch.uint8, device="cuda")
f = synth.frame_pair(600, h=H, w=W)[0]
frames = [np.roll(f, 37 * b, axis=1) for b in range(B)]
for b in range(B):
    d[b, :, :W] = torch.from_numpy(frames[b]).cuda()
eig = torch.empty(B, H, W, dtype=torch.float32, device="cuda"); em = torch.zeros(B, dtype=torch.float32, device="cuda")
R = torch.empty(B, H, W, dtype=torch.float64, device="cuda"); rm = torch.zeros(B, dtype=torch.float64, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(s)
    for i in range(reps):
        fn()
        ev[i + 1].record(s)
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return t[len(t) // 2] * 1e-3


t_me = timed(lambda: ctx.min_eigen_val_batched_dev(d.data_ptr(), B, H * pitch, H, W, pitch, eig.data_ptr(), em.data_ptr()))
t_st = timed(lambda: ctx.shitomasi_response_batched_dev(d.data_ptr(), B, H * pitch, H, W, pitch, R.data_ptr(), rm.data_ptr()))
out = {"frames": B, "rows": H, "cols": W,
       "min_eig": {"ms": t_me * 1e3, "gbs": B * H * W * 5 / t_me / 1e9, "frac": B * H * W * 5 / t_me / 1e9 / peak},
       "shitomasi_fp64": {"ms": t_st * 1e3, "gbs": B * H * W * 9 / t_st / 1e9, "frac": B * H * W * 9 / t_st / 1e9 / peak},
       "peak_gbs": peak}
try:
    import cv2
    b = B - 1
    ref = cv2.cornerMinEigenVal(frames[b], 3, ksize=3)
    got = eig[b].cpu().numpy()
    out["min_eig"]["max_abs_diff_vs_cv2_over_max"] = float(np.abs(got - ref).max() / ref.max())
    out["min_eig"]["max_equal"] = bool(abs(float(em[b]) - float(got.max())) == 0.0)
    one = ctx.min_eigen_val(frames[b])
    out["min_eig"]["batched_equals_single"] = bool(np.array_equal(one, got))
except Exception as e:  # noqa: BLE001
    out["min_eig"]["check_error"] = repr(e)
print(json.dumps(out))
