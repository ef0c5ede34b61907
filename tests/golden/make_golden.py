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
from harness import synth  # noqa: E402

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


def corners_small():
    """cv2 outputs for GFTT / FAST / cornerMinEigenVal; the reference-flavour ShiTomasi list comes from the
    numpy + cv2.blur restatement in tests/test_oracle_corners.py (the reference source is its own spec)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from test_oracle_corners import shitomasi_numpy
    img = synth.base_frame(21, 128, 160)
    xy = cv2.goodFeaturesToTrack(img, 60, 0.01, 5).reshape(-1, 2)
    kp = cv2.FastFeatureDetector_create(10, True).detect(img)
    k = np.array([[p.pt[0], p.pt[1], p.response] for p in kp])
    R = shitomasi_numpy(img, True)
    sel = R > R.max() * 0.4
    cand = np.argwhere(sel)
    order = np.argsort(-R[sel], kind="stable")[:50]
    np.savez_compressed(OUT / "corners_small.npz", img=img, gftt_xy=xy, eig=cv2.cornerMinEigenVal(img, 3, 3),
                        fast_col=k[:, 0].astype(np.int32), fast_row=k[:, 1].astype(np.int32),
                        fast_score=k[:, 2].astype(np.float32),
                        shi_row=cand[order][:, 0].astype(np.int32), shi_col=cand[order][:, 1].astype(np.int32),
                        shi_score=R[sel][order])


if __name__ == "__main__":
    lk_small()
    corners_small()
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
