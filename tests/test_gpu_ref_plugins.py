"""GPU: the product's drop-in adapters (practical-multi-view_b200/host/pmv_adapters.h -> libpmv_cuda.so), LINKED and RUN,
against the reference's own plugin classes compiled unchanged from /root/reference -- both live in
oracle/_ref/libpmv_ref.so (built on the CPU container by oracle/ref_build.py, shipped prebuilt) and both are driven
through the reference's own interfaces (BaseFeatureExtractor::extractFeatures, BaseFeatureMatcher::matchFeatures,
BaseOptimizer::apply) on identical Frame / OdometryPipeline state.  impl 0 = reference class, impl 1 = Gpu* adapter."""
import numpy as np
import pytest

from oracle import ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpmv_ref.so missing")]


@pytest.fixture(scope="module")
def frames(synth):
    return synth.frame_pair(3)


def test_shitomasi_response_adapter_vs_reference(frames):
    R0 = ref.shitomasi_response(frames[0], impl=0)
    R1 = ref.shitomasi_response(frames[0], impl=1)
    assert np.abs(R1 - R0).max() <= 1e-12 * np.abs(R0).max()


@pytest.mark.parametrize("which,max_feats,roi", [("shitomasi", 400, None), ("shitomasi", 40, (255, 0, 255, 255)),
                                                  ("gftt", 400, None), ("gftt", 40, (510, 0, 255, 255)), ("gftt", 40, (1020, 255, 221, 121)),
                                                  ("fast", 500, None), ("fast", 40, (0, 0, 255, 255))])
def test_extractor_adapter_vs_reference(frames, which, max_feats, roi):
    c0, r0, s0, t0 = ref.extract(which, frames[0], max_feats, roi=roi, impl=0)
    c1, r1, s1, t1 = ref.extract(which, frames[0], max_feats, roi=roi, impl=1)
    assert len(c0) == len(c1) and len(c0) > 0
    assert np.array_equal(t0, t1)                                                    # Feature::tracked as the reference leaves it
    if which == "shitomasi":                                                         # unstable std::sort: ties may permute
        assert np.abs(s1 - s0).max() <= 1e-12 * np.abs(s0).max()
        diff = np.nonzero((c0 != c1) | (r0 != r1))[0]
        assert all(np.isclose(s0[i], s0[i - 1]) or (i + 1 < len(s0) and np.isclose(s0[i], s0[i + 1])) for i in diff)
    else:
        assert np.array_equal(c0, c1) and np.array_equal(r0, r1) and np.array_equal(s0, s1)


def test_matcher_adapter_vs_reference(frames):
    import cv2
    f0, f1 = frames
    feats = cv2.goodFeaturesToTrack(f0, 400, 0.01, 5).reshape(-1, 2).astype(np.int32)
    feats = np.concatenate([feats, [[0, 0], [1240, 375], [3, 370], [620, 1]]]).astype(np.int32)   # border features
    corr0, n0 = ref.match(f0, f1, feats, impl=0)
    corr1, n1 = ref.match(f0, f1, feats, impl=1)
    assert n0 == n1 and corr0.shape == corr1.shape and n0 > 300
    assert np.array_equal(corr0[:, :2], corr1[:, :2])                                # identical status flags
    # Feature(int, int) truncates the tracked float position: 0.01 px of LK tolerance can flip a pixel boundary
    d = np.abs(corr0[:, 2:] - corr1[:, 2:])
    assert d.max() <= 1 and (d > 0).sum() <= max(2, len(d) // 100)


@pytest.mark.parametrize("n_poses,n_points,bundle,iters", [(5, 300, 5, 5), (6, 200, 3, 4), (20, 2000, 20, 5)])
def test_bundle_adjuster_adapter_vs_reference(synth, n_poses, n_points, bundle, iters):
    import cv2
    w = synth.ba_window(11 + n_poses, n_poses=n_poses, n_points=n_points)
    nf = n_poses + 1
    R = np.zeros((nf, 3, 3)); t = np.zeros((nf, 3)); R[0] = np.eye(3)
    for i in range(n_poses):
        R[i + 1] = cv2.Rodrigues(w["poses"][i, :3])[0].T; t[i + 1] = -w["poses"][i, 3:]
    args = (R, t, w["points"], w["cam_idx"] + 1, w["pt_idx"], w["obs"][:, 0], w["obs"][:, 1], w["K"], bundle, iters, n_poses)
    R0, t0, p0, s0 = ref.ba_apply(*args, impl=0)
    R1, t1, p1, s1 = ref.ba_apply(*args, impl=1)
    assert s0["iterations"] == s1["iterations"] and s0["successful_steps"] == s1["successful_steps"]
    assert abs(s0["initial_cost"] - s1["initial_cost"]) <= 1e-9 * s0["initial_cost"]
    assert abs(s0["final_cost"] - s1["final_cost"]) <= 1e-6 * s0["final_cost"]      # north_star gate
    assert np.abs(R0 - R1).max() < 1e-7 and np.abs(t0 - t1).max() < 1e-6
    assert np.abs(p0 - p1).max() <= 1e-4 * max(1.0, np.abs(p0).max())               # float32 storage of Feature3D


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_pnp_solver_adapter_vs_reference(seed):
    """pnpsolver->solvePnP(src, next, R, t): OpenCVEPnPSolver compiled unchanged from the reference (cv2.solvePnPRansac behind
    the shim) vs GpuEPnPSolver, on the same Frame / OdometryPipeline state: same pose, same 3-D points dropped as outliers,
    same next.map size."""
    import cv2
    from harness import pnp_scene
    sc = pnp_scene.scene(4000 + seed)
    n = len(sc["X"])
    pts = sc["X"].copy(); pts[:, 2] *= -1                      # the pipeline stores z < 0 in front of the camera (.cpp:26 flips it back)
    src_cr = np.stack([np.arange(n) % 1241, np.arange(n) // 7], 1)
    next_cr = sc["uv"].astype(np.int32)
    Rg = cv2.Rodrigues(sc["guess_r"])[0]
    out = [ref.pnp_solve(sc["K"], np.eye(3), np.zeros(3), pts, src_cr, next_cr, Rg, sc["guess_t"], impl=i) for i in (0, 1)]
    (R0, t0, k0, n0), (R1, t1, k1, n1) = out
    assert n0 == n1 == n
    assert np.array_equal(k0, k1) and k0.sum() >= n - len(sc["outliers"]) - 6 and not k0[sc["outliers"]].any()
    assert np.abs(R0 - R1).max() < 1e-7 and np.abs(t0 - t1).max() < 1e-6


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_triangulator_adapter_vs_reference(seed):
    """triangulator->triangulate(src, next, R, t): OpenCVFivePointTri compiled unchanged from the reference
    (cv2.findEssentialMat / cv2.recoverPose behind the shim) vs GpuFivePointTri, on the same Frame / OdometryPipeline state:
    same pose and scale, the same correspondences turned into Feature3D, the same world points."""
    from harness import twoview_scene
    sc = twoview_scene.scene(5000 + seed)
    gt0, gt1 = np.zeros(3), np.array([0.05, -0.02, 0.8 + 0.01 * seed])
    out = [ref.triangulate(sc["K"], sc["p1"].astype(np.int32), sc["p2"].astype(np.int32), gt0, gt1, impl=i) for i in (0, 1)]
    (R0, t0, s0, i0, p0), (R1, t1, s1, i1, p1) = out
    assert s0 == s1 == pytest.approx(np.linalg.norm(gt1))
    assert np.array_equal(i0, i1) and (i0 >= 0).sum() > 0.5 * len(i0)
    assert np.abs(R0 - R1).max() < 1e-6 and np.abs(t0 - t1).max() < 1e-6
    assert np.abs(p0 - p1).max() <= 1e-4 * np.abs(p0).max()                          # float32 storage of Feature3D
    assert np.abs(R0 - sc["R"]).max() < 3e-2
