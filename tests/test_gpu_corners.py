"""GPU parity (through the C ABI) of the extractors: GFTT (min-eig + NMS + min-distance), the
reference's own ShiTomasi extractor, FAST-9/16.  Bar (north_star): the same corner set as the
reference extractor up to response ties within 1e-5; integer detectors bit-exact."""
from pathlib import Path

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
TIE = 1e-5


def assert_same_corners_up_to_ties(got_xy, got_sc, want_xy, want_sc, vmax):
    """Identical ordered lists, or differences only where responses tie within TIE * max."""
    if got_xy.shape == want_xy.shape and np.array_equal(got_xy, want_xy):
        return
    g = {tuple(p): s for p, s in zip(got_xy.astype(int).tolist(), got_sc)}
    w = {tuple(p): s for p, s in zip(want_xy.astype(int).tolist(), want_sc)}
    assert abs(len(g) - len(w)) <= max(2, len(w) // 200)
    diff = set(g) ^ set(w)
    assert len(diff) <= max(4, len(w) // 100), f"{len(diff)} corners differ"
    # every differing corner must sit at a near-tie: its response is within TIE*max of a competitor's
    allv = np.array(sorted(list(g.values()) + list(w.values())))
    for p in diff:
        s = g.get(p, w.get(p))
        near = np.abs(allv - s) <= 4 * TIE * vmax
        assert near.sum() >= 2, f"corner {p} differs without a response tie"


@pytest.mark.parametrize("shape", [(376, 1241), (255, 255), (121, 221), (64, 70), (3, 40)])
def test_min_eigen_val_map(ctx, synth, shape):
    img = synth.base_frame(5, max(shape[0], 8), shape[1])[:shape[0]]
    img = np.ascontiguousarray(img)
    got, want = ctx.min_eigen_val(img), oracle.min_eigen_val(img)
    assert np.abs(got - want).max() <= TIE * want.max()


def test_min_eigen_val_roi_parent_borders(ctx, synth):
    img = synth.base_frame(6)
    for roi in [(255, 0, 255, 255), (1020, 255, 221, 121), (0, 0, 255, 255), (510, 255, 255, 121), (3, 5, 17, 9)]:
        got, want = ctx.min_eigen_val(img, roi), oracle.min_eigen_val(img, roi)
        assert np.abs(got - want).max() <= TIE * want.max(), roi


@pytest.mark.parametrize("mc,q,md", [(400, .01, 5), (40, .01, 5), (2000, .01, 3), (0, .05, 10), (100, .01, 0), (500, .01, 4.5)])
def test_gftt_vs_oracle_and_cv2(ctx, synth, mc, q, md):
    cv2 = pytest.importorskip("cv2")
    img = synth.base_frame(5)
    xy, sc = ctx.gftt(img, mc, q, md)
    oxy, osc = oracle.gftt(img, mc, q, md)
    vmax = float(oracle.min_eigen_val(img).max())
    assert_same_corners_up_to_ties(xy, sc, oxy, osc, vmax)
    c1 = cv2.goodFeaturesToTrack(img, mc, q, md).reshape(-1, 2)
    assert_same_corners_up_to_ties(xy, sc, c1, osc[:len(c1)] if len(osc) >= len(c1) else np.resize(osc, len(c1)), vmax)


def test_gftt_pipeline_grid_rois(ctx, synth):
    """The pipeline pattern (OdometryPipeline.cpp:351-357): 10 ROIs of <=255x255, 40 corners each."""
    img = synth.base_frame(9)
    for gy in range(0, 376, 255):
        for gx in range(0, 1241, 255):
            roi = (gx, gy, min(255, 1241 - gx), min(255, 376 - gy))
            xy, sc = ctx.gftt(img, 40, 0.01, 5, roi=roi)
            oxy, osc = oracle.gftt(img, 40, 0.01, 5, roi=roi)
            assert_same_corners_up_to_ties(xy, sc, oxy, osc, float(oracle.min_eigen_val(img, roi).max()))


def test_gftt_flat_image_and_errors(ctx, pmv):
    flat = np.full((60, 80), 128, np.uint8)
    xy, sc = ctx.gftt(flat, 50)
    assert len(xy) == 0
    with pytest.raises(pmv.PmvError):
        ctx.gftt(flat, 10, roi=(70, 0, 20, 20))          # ROI outside the image
    with pytest.raises(pmv.PmvError):
        ctx.gftt(flat, 10, block_size=5)                 # unsupported, like nothing the reference calls


def test_gftt_4k(ctx, synth):
    img = synth.base_frame(12, 2160, 3840)
    xy, sc = ctx.gftt(img, 10000, 0.01, 5)
    oxy, osc = oracle.gftt(img, 10000, 0.01, 5)
    assert len(xy) == 10000
    assert_same_corners_up_to_ties(xy, sc, oxy, osc, float(osc.max()))
    # size-independent properties: sorted by response, pairwise distance >= min_dist
    assert np.all(np.diff(sc) <= 0)
    from scipy.spatial import cKDTree
    assert len(cKDTree(xy).query_pairs(4.999)) == 0


@pytest.mark.parametrize("quirk", [True, False])
def test_shitomasi_reference_flavour(ctx, synth, quirk):
    img = synth.base_frame(5)
    R, Ro = ctx.shitomasi_response(img, quirk), oracle.shitomasi_response(img, quirk)
    assert np.abs(R - Ro).max() <= 1e-12 * Ro.max()
    col, row, sc = ctx.shitomasi(img, 400, 0.4, quirk)
    ocol, orow, osc = oracle.shitomasi(img, 400, 0.4, quirk)
    assert len(col) == len(ocol)
    assert np.allclose(sc, osc, rtol=1e-12)
    same = (col == ocol) & (row == orow)
    # positions may swap only between exactly-tied scores (std::sort is unstable in the reference)
    assert np.all(same | np.isclose(sc, np.roll(sc, 1), rtol=1e-12) | np.isclose(sc, np.roll(sc, -1), rtol=1e-12))


def test_shitomasi_roi_view_and_small(ctx, synth):
    big = synth.base_frame(7)
    view = big[0:255, 255:510]                      # strided view, like Frame::regionOfInterest
    col, row, sc = ctx.shitomasi(view, 40)
    ocol, orow, osc = oracle.shitomasi(np.ascontiguousarray(view), 40)
    assert np.array_equal(col, ocol) and np.array_equal(row, orow) and np.allclose(sc, osc, rtol=1e-12)
    tiny = np.ascontiguousarray(big[:3, :5])
    assert np.allclose(ctx.shitomasi_response(tiny), oracle.shitomasi_response(tiny))


@pytest.mark.parametrize("thr,nms", [(10, True), (25, False), (40, True)])
def test_fast_bit_exact(ctx, synth, thr, nms):
    img = synth.base_frame(5)
    col, row, sc, tot = ctx.fast(img, thr, nms)
    ocol, orow, osc = oracle.fast(img, thr, nms)
    assert tot == len(ocol) == len(col)
    assert np.array_equal(col, ocol) and np.array_equal(row, orow) and np.array_equal(sc, osc)
    # adapter semantics: first `max` in raster order (OpenCVFASTFeatureExtractor.cpp:11-20)
    col, row, sc, tot = ctx.fast(img, thr, nms, max_feats=40)
    assert np.array_equal(col, ocol[:40]) and np.array_equal(row, orow[:40]) and tot == len(ocol)


def test_fast_small_and_view(ctx, synth):
    big = synth.base_frame(8)
    for view in (big[:7, :9], big[10:131, 20:241], big[:6, :50]):
        col, row, sc, tot = ctx.fast(view, 10, True)
        ocol, orow, osc = oracle.fast(np.ascontiguousarray(view), 10, True)
        assert np.array_equal(col, ocol) and np.array_equal(row, orow) and np.array_equal(sc, osc)


def test_corner_golden_fixture(ctx):
    g = np.load(GOLD / "corners_small.npz")
    img = g["img"]
    xy, sc = ctx.gftt(img, 60, 0.01, 5)
    assert_same_corners_up_to_ties(xy, sc, g["gftt_xy"], sc, float(g["eig"].max()))
    col, row, fsc, _ = ctx.fast(img, 10, True)
    assert np.array_equal(col, g["fast_col"]) and np.array_equal(row, g["fast_row"]) and np.array_equal(fsc, g["fast_score"])
    col, row, ssc = ctx.shitomasi(img, 50)
    assert np.array_equal(col, g["shi_col"]) and np.array_equal(row, g["shi_row"])
    assert np.allclose(ssc, g["shi_score"], rtol=1e-12)


@pytest.mark.parametrize("shape", [(376, 1241), (97, 333)])
def test_batched_resident_responses_match_single_image_calls(ctx, synth, shape):
    """pmv_min_eigen_val_batched_dev / pmv_shitomasi_response_batched_dev (images resident in HBM, one launch for the
    batch) == the single-image host entry points, bit for bit, maxima included."""
    import torch
    h, w = shape
    B = 3
    imgs = np.stack([synth.frame_pair(40 + b, h=h, w=w)[0] for b in range(B)])
    pitch = (w + 15) // 16 * 16
    d = torch.zeros(B, h, pitch, dtype=torch.uint8, device="cuda")
    d[:, :, :w] = torch.from_numpy(imgs).cuda()
    eig = torch.zeros(B, h, w, dtype=torch.float32, device="cuda"); emax = torch.zeros(B, dtype=torch.float32, device="cuda")
    R = torch.zeros(B, h, w, dtype=torch.float64, device="cuda"); rmax = torch.zeros(B, dtype=torch.float64, device="cuda")
    ctx.min_eigen_val_batched_dev(d.data_ptr(), B, h * pitch, h, w, pitch, eig.data_ptr(), emax.data_ptr())
    ctx.shitomasi_response_batched_dev(d.data_ptr(), B, h * pitch, h, w, pitch, R.data_ptr(), rmax.data_ptr())
    ctx.sync()
    for b in range(B):
        e1 = ctx.min_eigen_val(imgs[b])
        r1 = ctx.shitomasi_response(imgs[b])
        assert np.array_equal(eig[b].cpu().numpy(), e1)
        assert np.array_equal(R[b].cpu().numpy(), r1)
        assert float(emax[b]) == float(e1.max()) and float(rmax[b]) == float(r1.max())


@pytest.mark.parametrize("shape,roi,pitch_pad", [((376, 1241), None, 7), ((376, 1241), (256, 60, 255, 255), 7), ((376, 1241), (255, 61, 250, 200), 0),
                                                 ((720, 1280), None, 0)])
def test_gftt_on_resident_image_matches_host_call(ctx, synth, shape, roi, pitch_pad):
    """pmv_gftt_dev reads an image that is already in HBM in place (parent borders included) and leaves the corner
    list on the device: same ordered list and scores as pmv_gftt on the host copy, for whole images and for ROI views,
    with word-aligned and unaligned pitches / ROI origins (fast kernels vs tile kernels)."""
    import torch
    h, w = shape
    img = synth.frame_pair(77, h=h, w=w)[0]
    pitch = w + pitch_pad
    d = torch.zeros(h, pitch, dtype=torch.uint8, device="cuda")
    d[:, :w] = torch.from_numpy(img).cuda()
    nmax = 700
    d_xy = torch.zeros(nmax, 2, dtype=torch.float32, device="cuda"); d_sc = torch.zeros(nmax, dtype=torch.float32, device="cuda")
    n = ctx.gftt_dev(d.data_ptr(), h, w, pitch, nmax, d_xy.data_ptr(), d_sc.data_ptr(), 0.01, 5.0, roi=roi)
    ctx.sync()
    xy, sc = ctx.gftt(img, nmax, 0.01, 5.0, roi=roi)
    assert n == len(xy) and n > 50
    assert np.array_equal(d_xy[:n].cpu().numpy(), xy)
    assert np.array_equal(d_sc[:n].cpu().numpy(), sc)


def test_fast_many_keypoints_bucket_sorted(ctx):
    """More than 8192 keypoints: the raster-order sort goes through buckets of the pixel index (sort.cuh) instead of the
    bitonic network.  Same list, scores included, as cv2 on a noisy 1280 x 1024 frame."""
    import cv2
    rng = np.random.default_rng(4)
    img = cv2.GaussianBlur(rng.integers(0, 256, (1024, 1280), dtype=np.uint8), (0, 0), 1.2)
    col, row, sc, tot = ctx.fast(img, 10, True)
    kp = cv2.FastFeatureDetector_create(10, True).detect(img)
    assert tot == len(kp) == len(col) and tot > 8192
    assert np.array_equal(col, [int(k.pt[0]) for k in kp]) and np.array_equal(row, [int(k.pt[1]) for k in kp])
    assert np.array_equal(sc, [k.response for k in kp])


def test_gftt_plateaus_fall_back_to_the_bitonic_sort(ctx):
    """A large periodic image has thousands of corners with EQUAL responses: the bucket sort of the candidate list would
    degenerate, the selector falls back to the bitonic network, and ties still come out in OpenCV's order (higher address
    first, greaterThanPtr)."""
    import cv2
    tile = np.zeros((16, 16), np.uint8)
    tile[4:12, 4:12] = 200
    tile[6:10, 6:10] = 90
    img = np.tile(tile, (64, 80))          # 1024 x 1280, 5120 identical cells
    xy, sc = ctx.gftt(img, 3000, 0.01, 5)
    want = cv2.goodFeaturesToTrack(img, 3000, 0.01, 5).reshape(-1, 2)
    assert len(xy) == len(want) == 3000
    assert np.array_equal(xy, want)
