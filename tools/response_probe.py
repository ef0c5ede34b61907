"""Batched corner-response kernels on 8 resident 4K frames (what bench.py's extract leg times); short, for ncu."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import pmv_b200
from harness import synth
H, W, B = 2160, 3840, 8
ctx = pmv_b200.Context(0)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
pitch = (W + 15) // 16 * 16
d = torch.zeros(B, H, pitch, dtype=torch.uint8, device="cuda")
f = synth.frame_pair(600, h=H, w=W)[0]
for b in range(B):
    d[b, :, :W] = torch.from_numpy(np.roll(f, 37 * b, axis=1)).cuda()
eig = torch.empty(B, H, W, dtype=torch.float32, device="cuda"); em = torch.zeros(B, dtype=torch.float32, device="cuda")
R = torch.empty(B, H, W, dtype=torch.float64, device="cuda"); rm = torch.zeros(B, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.min_eigen_val_batched_dev(d.data_ptr(), B, H * pitch, H, W, pitch, eig.data_ptr(), em.data_ptr())
    ctx.shitomasi_response_batched_dev(d.data_ptr(), B, H * pitch, H, W, pitch, R.data_ptr(), rm.data_ptr())
torch.cuda.synchronize()
print("ok", float(em.max()), float(rm.max()))
