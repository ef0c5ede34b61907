"""CPU: pin the corner-detector oracles (oracle/pmv_oracle_corners.c) against cv2 4.13 -- the real
OpenCV kernels behind OpenCVGoodFeatureExtractor.cpp:7 / OpenCVFASTFeatureExtractor.cpp:8 -- and the
reference's own ShiTomasiFeatureExtractor against a numpy + cv2.blur restatement of its source."""
from pathlib import Path

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import oracle

GOLD = Path(__file__).parent / "golden"


def shitomasi_numpy(img, quirk=True):
    """Line-by-line numpy form of Frame.cpp:58-86,119-138 + ShiTomasiFeatureExtractor.cpp:49-75."""
    s = img.view(np.int8).astype(np.float64) if quirk else img.astype(np.float64)
    gx, gy = np.zeros_like(s), np.zeros_like(s)
    gx[1:-1, 1:-1] = 0.5 * s[1:-1, 2:] - 0.5 * s[1:-1, :-2]
    gy[1:-1, 1:-1] = 0.5 * s[2:, 1:-1] - 0.5 * s[:-2, 1:-1]
    H3 = cv2.blur(np.stack([gx * gx, gy * gy, gx * gy], -1), (3, 3))
    Ixx, Iyy, Ixy = H3[..., 0], H3[..., 1], H3[..., 2]
    B, Cc = -Ixx - Iyy, Ixx * Iyy - Ixy ** 2
    with np.errstate(invalid="ignore"):
        l1, l2 = (-B + np.sqrt(B ** 2 - 4 * Cc)) / 2, (-B - np.sqrt(B ** 2 - 4 * Cc)) / 2
    R = np.minimum(l1, l2)
    R[:, -1] = 0
    return R


def test_min_eigen_val_vs_cv2(synth):
    img = synth.base_frame(5)
    e_cv = cv2.cornerMinEigenVal(img, 3, 3)
    e_or = oracle.min_eigen_val(img)
    assert np.abs(e_cv - e_or).max() <= 1e-5 * e_cv.max()      # north_star tie tolerance (observed 9e-7)


def test_min_eigen_val_roi_reads_parent(synth):
    """C++ sub-Mat semantics: Sobel on the full image, crop, isolated box filter on the crop."""
    img = synth.base_frame(6)
    for (x, y, w, h) in [(255, 0, 255, 255), (1020, 255, 221, 121), (0, 0, 255, 255)]:
        s = 1 / 3060.
        Dx = cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=3, scale=s)[y:y + h, x:x + w]
        Dy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=3, scale=s)[y:y + h, x:x + w]
        box = cv2.boxFilter(np.stack([Dx * Dx, Dx * Dy, Dy * Dy], -1), -1, (3, 3), normalize=False)
        a, b, c = box[..., 0] * np.float32(.5), box[..., 1], box[..., 2] * np.float32(.5)
        want = (a + c) - np.sqrt((a - c) * (a - c) + b * b)
        got = oracle.min_eigen_val(img, (x, y, w, h))
        assert np.abs(want - got).max() <= 1e-5 * want.max()


@pytest.mark.parametrize("mc,q,md", [(400, .01, 5), (40, .01, 5), (2000, .01, 3), (0, .05, 10), (100, .01, 0), (500, .01, 4.5)])
def test_gftt_ordered_list_vs_cv2(synth, mc, q, md):
    img = synth.base_frame(5)
    c1 = cv2.goodFeaturesToTrack(img, mc, q, md).reshape(-1, 2)
    c2, _ = oracle.gftt(img, mc, q, md)
    assert np.array_equal(c1, c2)


def test_gftt_select_on_cv_map_is_exact(synth):
    """Selection logic alone (same response map) must reproduce cv2's ordered list bit for bit."""
    for seed, mc, md in [(7, 400, 5), (8, 1500, 7.3), (9, 0, 12)]:
        img = synth.base_frame(seed, 300, 400)
        e = cv2.cornerMinEigenVal(img, 3, 3)
        c1 = cv2.goodFeaturesToTrack(img, mc, 0.01, md).reshape(-1, 2)
        c2, _ = oracle.gftt_select(e, mc, 0.01, md)
        assert np.array_equal(c1, c2)


def test_fast_vs_cv2(synth):
    img = synth.base_frame(5)
    for thr, nms in [(10, True), (25, False), (40, True)]:
        kp = cv2.FastFeatureDetector_create(thr, nms).detect(img)
        col, row, sc = oracle.fast(img, thr, nms)
        k = np.array([[p.pt[0], p.pt[1], p.response] for p in kp]).reshape(-1, 3)
        assert len(kp) == len(col)
        assert np.array_equal(k[:, 0], col) and np.array_equal(k[:, 1], row) and np.array_equal(k[:, 2], sc)


def test_shitomasi_reference_flavour(synth):
    img = synth.base_frame(5)
    for quirk in (True, False):
        R1, R2 = shitomasi_numpy(img, quirk), oracle.shitomasi_response(img, quirk)
        assert np.isnan(R1).sum() == np.isnan(R2).sum() == 0
        assert np.abs(R1 - R2).max() <= 1e-12 * R1.max()
    R1 = shitomasi_numpy(img, True)
    col, row, sc = oracle.shitomasi(img, 400)
    sel = R1 > R1.max() * 0.4
    cand = np.argwhere(sel)
    order = np.argsort(-R1[sel], kind="stable")[:400]
    assert np.array_equal(cand[order][:, 0], row) and np.array_equal(cand[order][:, 1], col)
    assert np.allclose(sc, R1[sel][order], rtol=1e-12)


def test_corner_golden_fixture():
    g = np.load(GOLD / "corners_small.npz")
    img = g["img"]
    c2, _ = oracle.gftt(img, 60, 0.01, 5)
    assert np.array_equal(c2, g["gftt_xy"])
    col, row, sc = oracle.fast(img, 10, True)
    assert np.array_equal(col, g["fast_col"]) and np.array_equal(row, g["fast_row"]) and np.array_equal(sc, g["fast_score"])
    assert np.abs(oracle.min_eigen_val(img) - g["eig"]).max() <= 1e-5 * g["eig"].max()
    col, row, sc = oracle.shitomasi(img, 50)
    assert np.array_equal(col, g["shi_col"]) and np.array_equal(row, g["shi_row"])
