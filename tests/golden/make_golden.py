"""Regenerate the committed golden fixtures from cv2 4.13.0 -- the in-image build of the very
OpenCV kernels the reference calls (the reference has no tests / golden vectors of its own,
SURVEY.md §4).  Run from the repo root:  python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import pmv_b200  # noqa: E402
from pmv_b200 import synth  # noqa: E402

OUT = Path(__file__).parent


def lk_small():
    f0, f1 = synth.frame_pair(11, h=120, w=160)
    pts = synth.track_points(f0, 60, 5)
    pts = np.concatenate([pts, np.float32([[0, 0], [159, 119], [-4, 3], [170, 50], [80.5, 60.25]])])
    win, ml = (15, 15), 2
    nx, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), None, winSize=win, maxLevel=ml)
    p1 = cv2.pyrDown(f0)
    p2 = cv2.pyrDown(p1)
    np.savez_compressed(OUT / "lk_small.npz", prev=f0, next=f1, pts=pts, win=np.int32(win), max_level=np.int32(ml),
                        next_xy=nx.reshape(-1, 2), status=st.ravel(), err=err.ravel(), pyr1=p1, pyr2=p2)


if __name__ == "__main__":
    lk_small()
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
