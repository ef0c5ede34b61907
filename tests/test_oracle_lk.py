"""CPU: pin the C oracle (oracle/pmv_oracle_lk.c) against the real OpenCV kernels (cv2 4.13)
that the reference calls at OpenCVLucasKanadeFM.cpp:15, and against committed fixtures."""
from pathlib import Path

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import oracle

GOLD = Path(__file__).parent / "golden"

EDGE_PTS = np.array([[0, 0], [1240, 375], [-5, 10], [1250, 100], [600, -3], [3.5, 370.2], [1239.7, 2.1],
                     [-25, -25], [1262, 380], [620.5, 188.25]], np.float32)


def test_pyr_down_bit_exact(synth):
    f0, _ = synth.frame_pair(1)
    for img in (f0, f0[:47, :156].copy(), f0[:5, :7].copy(), f0[:1, :9].copy(), f0[:33, :34].copy()):
        assert np.array_equal(oracle.pyr_down(img), cv2.pyrDown(img))


def test_pyr_levels_matches_cv2(synth):
    f0, _ = synth.frame_pair(1)
    for win, ml in [((21, 21), 3), ((32, 32), 4), ((21, 21), 4), ((21, 21), 0), ((15, 15), 6)]:
        lv, _ = cv2.buildOpticalFlowPyramid(f0, win, ml, None, False)
        assert oracle.pyr_levels(376, 1241, win[0], win[1], ml) == lv


def test_scharr_bit_exact(synth):
    f0, _ = synth.frame_pair(2)
    _, pyr = cv2.buildOpticalFlowPyramid(f0, (21, 21), 2, None, True)
    for l in range(3):
        img = np.ascontiguousarray(pyr[2 * l])
        assert np.array_equal(oracle.scharr(img), pyr[2 * l + 1])


@pytest.mark.parametrize("win,ml", [((21, 21), 3), ((32, 32), 4), ((15, 15), 2), ((9, 13), 3)])
def test_lk_matches_cv2(synth, win, ml):
    f0, f1 = synth.frame_pair(3)
    pts = np.concatenate([synth.track_points(f0, 500, 7), EDGE_PTS])
    n1, st1, e1 = cv2.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), None, winSize=win, maxLevel=ml)
    n2, st2, e2 = oracle.lk_track(f0, f1, pts, win, ml)
    n1, st1, e1 = n1.reshape(-1, 2), st1.ravel(), e1.ravel()
    assert np.array_equal(st1, st2)
    ok = st1 == 1
    assert np.abs(n1[ok] - n2[ok]).max() < 0.01  # north_star tolerance (observed ~4e-4)
    assert np.abs(e1[ok] - e2[ok]).max() < 0.01


def test_lk_flags(synth):
    f0, f1 = synth.frame_pair(4)
    pts = synth.track_points(f0, 300, 9)
    init = pts + np.float32([1.5, -1.0])
    n1, st1, e1 = cv2.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), init.reshape(-1, 1, 2).copy(),
                                           winSize=(21, 21), maxLevel=3,
                                           flags=cv2.OPTFLOW_USE_INITIAL_FLOW | cv2.OPTFLOW_LK_GET_MIN_EIGENVALS)
    n2, st2, e2 = oracle.lk_track(f0, f1, pts, (21, 21), 3, flags=4 | 8, init=init)
    assert np.array_equal(st1.ravel(), st2)
    ok = st2 == 1
    assert np.abs(n1.reshape(-1, 2)[ok] - n2[ok]).max() < 0.01
    assert np.allclose(e1.ravel()[ok], e2[ok], rtol=1e-4, atol=1e-6)


def test_lk_golden_fixture():
    """Fixture written by tests/golden/make_golden.py from cv2 (the reference's own kernel)."""
    g = np.load(GOLD / "lk_small.npz")
    nx, st, err = oracle.lk_track(g["prev"], g["next"], g["pts"], tuple(g["win"]), int(g["max_level"]))
    assert np.array_equal(st, g["status"])
    ok = st == 1
    assert np.abs(nx[ok] - g["next_xy"][ok]).max() < 0.01
    assert np.abs(err[ok] - g["err"][ok]).max() < 0.01
    for l, lvl in enumerate([g["pyr1"], g["pyr2"]]):
        img = g["prev"] if l == 0 else g["pyr1"]
        assert np.array_equal(oracle.pyr_down(np.ascontiguousarray(img)), lvl)
