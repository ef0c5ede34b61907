"""GPU: pmv_find_essential_mat / pmv_recover_pose / pmv_five_point_pose (fivept.cu: 256 samples at a time, sequential
accept rule replayed in order, cheirality vote over the CTA) == cv2.findEssentialMat(RANSAC, 0.99, 1) + cv2.recoverPose as
OpenCVFivePointTri.cpp:25-27 calls them, and == the CPU build of the same arithmetic."""
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
INF = float("inf")


def test_find_essential_mat_matches_opencv(ctx):
    from harness import twoview_scene
    tight, total = 0, 60
    for seed in range(total):
        sc = twoview_scene.scene(6000 + seed)
        Ecv, mcv = cv2.findEssentialMat(sc["p1"], sc["p2"], sc["K"], cv2.RANSAC, 0.99, 1.0)
        E, m = ctx.find_essential_mat(sc["p1"], sc["p2"], sc["K"])
        assert E is not None and np.array_equal(m, mcv.ravel()), seed
        d = np.abs(E - Ecv).max()
        assert d < 1e-3, (seed, d)
        tight += d < 1e-7
    assert tight >= 0.8 * total


def test_recover_pose_matches_opencv(ctx):
    from harness import twoview_scene
    for seed in range(30):
        sc = twoview_scene.scene(7000 + seed)
        Ecv, mcv = cv2.findEssentialMat(sc["p1"], sc["p2"], sc["K"], cv2.RANSAC, 0.99, 1.0)
        for dist in (INF, 25.0):
            ncv, Rcv, tcv, m2cv, tricv = cv2.recoverPose(Ecv, sc["p1"], sc["p2"], sc["K"], distanceThresh=dist, mask=mcv.copy())
            good, R, t, m, tri = ctx.recover_pose(Ecv, sc["p1"], sc["p2"], sc["K"], dist, mcv)
            assert good == ncv and np.array_equal(m, (m2cv.ravel() != 0).astype(np.uint8))
            assert np.abs(R - Rcv).max() < 1e-12 and np.abs(t - tcv.ravel()).max() < 1e-12
            ok = m != 0
            q = tri[:3, ok] / tri[3, ok]; qcv = tricv[:3, ok] / tricv[3, ok]
            assert (np.abs(q - qcv) / np.abs(qcv).max(0)).max() < 1e-8
        # without a mask every correspondence votes
        ncv, _, _, m2cv, _ = cv2.recoverPose(Ecv, sc["p1"], sc["p2"], sc["K"], distanceThresh=INF)
        good, _, _, m, _ = ctx.recover_pose(Ecv, sc["p1"], sc["p2"], sc["K"])
        assert good == ncv and np.array_equal(m != 0, m2cv.ravel() != 0)


def test_fused_call_equals_the_two_calls_and_the_cpu_build(ctx):
    """One launch for both OpenCV calls vs the two entry points vs the serial CPU driver of fivept_math.cuh."""
    from harness import twoview_scene
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import test_fivept_host as th
    host = th.build_host()
    for seed in range(20):
        sc = twoview_scene.scene(8000 + seed)
        r = ctx.five_point_pose(sc["p1"], sc["p2"], sc["K"])
        E, m = ctx.find_essential_mat(sc["p1"], sc["p2"], sc["K"])
        assert np.array_equal(r["E"], E) and np.array_equal(r["ransac_mask"], m) and r["n_inliers"] == int(m.sum())
        good, R, t, m2, tri = ctx.recover_pose(E, sc["p1"], sc["p2"], sc["K"], INF, m)
        assert good == r["n_good"] and np.array_equal(m2, r["mask"])
        assert np.array_equal(R, r["R"]) and np.array_equal(t, r["t"]) and np.array_equal(tri, r["tri"])
        hg, hE, hm, _ = th.host_find_essential(host, sc)
        assert hg == r["n_inliers"] and np.array_equal(hm, m)
        assert np.abs(hE - E).max() < 1e-6                      # same arithmetic; the device contracts a*b+c into FMAs
        g2, hR, ht, hm2, _ = th.host_recover_pose(host, hE, sc, hm)
        assert g2 == good and np.array_equal(hm2, m2) and np.abs(hR - R).max() < 1e-6


def test_no_model_and_small_inputs(ctx, pmv):
    rng = np.random.default_rng(0)
    K = np.array([[700.0, 0, 600], [0, 700, 180], [0, 0, 1]])
    p1 = rng.uniform(0, 1200, (5, 2)); p2 = rng.uniform(0, 370, (5, 2))
    with pytest.raises(pmv.PmvError):
        ctx.find_essential_mat(p1, p2, K)                       # five points: OpenCV stacks every root, recoverPose would throw
    # pure noise: whatever cv2 decides (usually a handful of chance inliers), the same here
    p1 = np.trunc(rng.uniform(0, 1200, (80, 2))); p2 = np.trunc(rng.uniform(0, 370, (80, 2)))
    Ecv, mcv = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.99, 1.0)
    E, m = ctx.find_essential_mat(p1, p2, K)
    assert (E is None) == (Ecv is None)
    if E is not None:
        assert np.array_equal(m, mcv.ravel())
