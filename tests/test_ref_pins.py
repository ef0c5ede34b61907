"""CPU: the oracle (plain-C restatement, oracle/*.c) against the REFERENCE's own code.

oracle/_ref/libpmv_ref.so holds the reference's translation units compiled unchanged from /root/reference
(oracle/ref_build.py lists them) against the functional OpenCV / Ceres shim in oracle/ref_shim/, with the third-party
OpenCV kernels forwarded to the real cv2 wheel.  These tests pin the restatement to it:
  a6-a9  Frame::computeSpatialGradient / computeHarrisMatrix, ShiTomasiFeatureExtractor (response + extractFeatures)
  a11-12 ProjectionResidual::operator() under real automatic differentiation (ProjectionResidual::Create -> Evaluate)
  a13    CeresBundleAdjustment::apply's parameterisation and write-back (the minimiser behind ceres::Solve is the oracle's
         LM -- Ceres is absent, that part stays unpinned)
  a1/a5/a10  OpenCVLucasKanadeFM / OpenCVGoodFeatureExtractor / OpenCVFASTFeatureExtractor marshalling around real cv2."""
import numpy as np
import pytest

import oracle
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpmv_ref.so neither present nor buildable")


@pytest.fixture(scope="module")
def frames(synth):
    return synth.frame_pair(3)


def test_ref_lists_reference_sources():
    src = ref.lib().ref_sources().decode().split()
    for f in ("Frame.cpp", "ShiTomasiFeatureExtractor.cpp", "ProjectionResidual.cpp", "CeresBundleAdjustment.cpp", "OpenCVLucasKanadeFM.cpp"):
        assert f in src


@pytest.mark.parametrize("small_angle", [False, True])
def test_oracle_residual_and_jacobians_equal_reference_autodiff(synth, small_angle):
    w = synth.ba_window(3, n_poses=6, n_points=300)
    poses = w["poses"].copy()
    if small_angle:                      # |a|^2 <= DBL_EPSILON: AngleAxisRotatePoint's first-order branch
        poses[:, :3] = np.random.default_rng(0).normal(0, 1e-9, (6, 3))
    r, Jc, Jp = ref.residual(poses, w["points"], w["obs"], w["K"], w["cam_idx"], w["pt_idx"])
    ro, Jco, Jpo, _ = oracle.ba_eval(poses, w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"])
    # north_star tolerance: 1e-5 relative; observed: bit-identical (same operation order under Jets)
    assert np.abs(r - ro).max() <= 1e-12 * np.abs(ro).max()
    assert np.abs(Jc - Jco).max() <= 1e-12 * np.abs(Jco).max()
    assert np.abs(Jp - Jpo).max() <= 1e-12 * np.abs(Jpo).max()
    r2, _, _ = ref.residual(poses, w["points"], w["obs"], w["K"], w["cam_idx"], w["pt_idx"], jac=False)   # double instantiation
    assert np.abs(r2 - ro).max() <= 1e-12 * np.abs(ro).max()


@pytest.mark.parametrize("hooks", [True, False])
def test_oracle_shitomasi_response_equals_reference(frames, hooks):
    ref.use_cv2_hooks(hooks)             # True: cv::blur is the real cv2.blur; False: the shim's 9-term sums
    try:
        R = ref.shitomasi_response(frames[0])
    finally:
        ref.use_cv2_hooks(True)
    Ro = oracle.shitomasi_response(frames[0])
    assert not np.isnan(R).any()
    assert np.abs(R - Ro).max() <= 1e-12 * np.abs(Ro).max()


def test_reference_frame_planes(frames):
    """a7 / a8: the gradient planes read the u8 image through schar* (Frame.cpp:65-67), the harris Mat is the 3x3 blur."""
    import cv2
    img = frames[0]
    gx, gy, hm = ref.frame_planes(img)
    s = img.astype(np.int8).astype(np.float64)
    ex = np.zeros_like(s); ey = np.zeros_like(s)
    ex[1:-1, 1:-1] = 0.5 * s[1:-1, 2:] - 0.5 * s[1:-1, :-2]
    ey[1:-1, 1:-1] = 0.5 * s[2:, 1:-1] - 0.5 * s[:-2, 1:-1]
    assert np.array_equal(gx, ex) and np.array_equal(gy, ey)
    assert np.array_equal(hm, cv2.blur(np.stack([ex * ex, ey * ey, ex * ey], -1), (3, 3)))


@pytest.mark.parametrize("max_feats", [40, 400, 5000])
def test_oracle_shitomasi_list_equals_reference(frames, max_feats):
    c, r, s, tr = ref.extract("shitomasi", frames[0], max_feats)
    co, ro, so = oracle.shitomasi(frames[0], max_feats)
    assert len(c) == len(co)
    assert np.array_equal(s, so)                                  # scores (sorted, descending) identical
    same = (c == co) & (r == ro)                                  # positions may differ only inside runs of equal score
    assert all(np.isclose(s[i], s[i - 1]) or (i + 1 < len(s) and np.isclose(s[i], s[i + 1])) for i in np.nonzero(~same)[0])
    assert not tr.any()                                           # Feature() leaves tracked = false (.cpp:24)


def test_reference_shitomasi_roi_is_isolated(frames):
    """A ROI Frame recomputes gradient + harris on the view alone (Frame.cpp:95-117: flags false on the new Frame)."""
    roi = (255, 0, 255, 255)
    x, y, w, h = roi
    c, r, s, _ = ref.extract("shitomasi", frames[0], 40, roi=roi)
    co, ro, so = oracle.shitomasi(np.ascontiguousarray(frames[0][y:y + h, x:x + w]), 40)
    assert np.array_equal(c, co) and np.array_equal(r, ro) and np.array_equal(s, so)


def test_reference_opencv_extractors_and_matcher(frames):
    """The reference's OpenCV* plugins (compiled unchanged) around real cv2 == the oracle's view of them."""
    import cv2
    f0, f1 = frames
    c, r, s, tr = ref.extract("gftt", f0, 400)
    cc = cv2.goodFeaturesToTrack(f0, 400, 0.01, 5).reshape(-1, 2)
    assert np.array_equal(np.stack([c, r], 1), cc.astype(np.int32)) and not s.any() and tr.all()
    roi = (510, 0, 255, 255)
    c, r, _, _ = ref.extract("gftt", f0, 40, roi=roi)
    xy, _ = oracle.gftt(f0, 40, 0.01, 5.0, roi=roi)
    assert np.array_equal(np.stack([c, r], 1), xy.astype(np.int32))
    c, r, s, tr = ref.extract("fast", f0, 500)
    co, ro, so = oracle.fast(f0, 10, True, 500)
    assert np.array_equal(c, co) and np.array_equal(r, ro) and np.array_equal(s, so) and tr.all()
    feats = cc.astype(np.int32)
    corr, sz = ref.match(f0, f1, feats)
    nx, st, _ = oracle.lk_track(f0, f1, feats.astype(np.float32), (32, 32), 4)
    ok = st == 1
    exp = np.concatenate([feats[ok], nx[ok].astype(np.int32)], 1)
    exp = exp[np.lexsort(exp.T[::-1])]
    assert sz == int(ok.sum()) and np.array_equal(corr, exp)


def _pipeline_state(synth, seed, n_poses, n_points):
    import cv2
    w = synth.ba_window(seed, n_poses=n_poses, n_points=n_points)
    nf = n_poses + 1                                           # frame 0 is never a parameter (CeresBundleAdjustment.cpp:22)
    R = np.zeros((nf, 3, 3)); t = np.zeros((nf, 3)); R[0] = np.eye(3)
    for i in range(n_poses):
        R[i + 1] = cv2.Rodrigues(w["poses"][i, :3])[0].T; t[i + 1] = -w["poses"][i, 3:]
    return w, R, t


def test_reference_ba_apply_equals_oracle_solve(synth):
    import cv2
    n = 5
    w, R, t = _pipeline_state(synth, 11, n, 300)
    R2, t2, p2, s = ref.ba_apply(R, t, w["points"], w["cam_idx"] + 1, w["pt_idx"], w["obs"][:, 0], w["obs"][:, 1], w["K"], n, 5, n, impl=0)
    po, xo, so = oracle.ba_solve(w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"], 1.0, 5)
    assert s["iterations"] == so["iterations"] and s["successful_steps"] == so["successful_steps"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-9 * so["final_cost"]       # gate 1e-6; differences: Rodrigues round trip
    Ro = np.stack([cv2.Rodrigues(po[i, :3])[0].T for i in range(n)])
    assert np.abs(Ro - R2[1:]).max() < 1e-9 and np.abs(-po[:, 3:] - t2[1:]).max() < 1e-8
    assert np.abs(xo.astype(np.float32) - p2).max() <= 1e-5                         # Feature3D stores float (Feature3D.cpp:111-116)
    assert np.array_equal(R2[0], np.eye(3))


def test_reference_ba_window_smaller_than_history(synth):
    """bundle_size < frames: only the last bundle_size frames are parameters, earlier ones stay untouched."""
    w, R, t = _pipeline_state(synth, 12, 6, 200)
    R2, t2, p2, s = ref.ba_apply(R, t, w["points"], w["cam_idx"] + 1, w["pt_idx"], w["obs"][:, 0], w["obs"][:, 1], w["K"], 3, 4, 6, impl=0)
    assert np.array_equal(R2[:4], R[:4]) and np.array_equal(t2[:4], t[:4])
    sel = w["cam_idx"] >= 3
    used = np.unique(w["pt_idx"][sel]); remap = -np.ones(200, np.int64); remap[used] = np.arange(len(used))
    po, xo, so = oracle.ba_solve(w["poses"][3:], w["points"][used], w["obs"][sel], w["cam_idx"][sel] - 3, remap[w["pt_idx"][sel]], w["K"], 1.0, 4)
    assert s["iterations"] == so["iterations"] and abs(s["final_cost"] - so["final_cost"]) <= 1e-9 * so["final_cost"]
