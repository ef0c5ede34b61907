"""GPU: pmv_pnp_ransac (pnp.cu: hypotheses eight at a time, sequential accept rule replayed in order, refinement in the
same launch) == cv2.solvePnPRansac as OpenCVEPnPSolver.cpp:34-35 calls it, and == the CPU build of the same arithmetic."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _cv(sc):
    ok, r, t, inl = cv2.solvePnPRansac(sc["X"], sc["uv"], sc["K"], None, sc["guess_r"].reshape(3, 1).copy(), sc["guess_t"].reshape(3, 1).copy(),
                                       True, 100, 8.0, 0.99)
    return ok, r.ravel(), t.ravel(), (inl.ravel() if inl is not None else np.zeros(0, np.int32))


def test_pnp_ransac_matches_opencv(ctx):
    from harness import pnp_scene
    same, total = 0, 80
    for seed in range(total):
        sc = pnp_scene.scene(2000 + seed)
        ok, r, t, inl = _cv(sc)
        g_ok, gr, gt, ginl = ctx.pnp_ransac(sc["X"], sc["uv"], sc["K"], sc["guess_r"], sc["guess_t"], True, 100, 8.0, 0.99)
        assert ok and g_ok
        if np.array_equal(ginl, inl):
            same += 1
            assert np.abs(gr - r).max() < 1e-7 and np.abs(gt - t).max() < 1e-6      # same inliers -> same minimiser
        else:                                                                       # another 5-point hypothesis won
            assert len(set(ginl.tolist()) ^ set(inl.tolist())) <= max(6, len(inl) // 25)
            assert np.abs(gr - r).max() < 1e-3 and np.abs(gt - t).max() < 2e-2
        # the pose explains the scene: outliers rejected, truth recovered to the noise level
        assert not set(sc["outliers"].tolist()) & set(ginl.tolist())
        assert np.abs(gr - sc["rvec"]).max() < 5e-3 and np.abs(gt - sc["tvec"]).max() < 0.2
    assert same >= 0.95 * total


def test_pnp_ransac_equals_cpu_build_of_the_same_arithmetic(ctx):
    """Device schedule (8 hypotheses per batch, CTA-wide reductions) vs the serial CPU driver of pnp_math.cuh."""
    from harness import pnp_scene
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_pnp_host import build_host, host_ransac
    lib = build_host()
    for seed in range(20):
        sc = pnp_scene.scene(3000 + seed)
        good, r, t, inl = host_ransac(lib, sc)
        g_ok, gr, gt, ginl = ctx.pnp_ransac(sc["X"], sc["uv"], sc["K"], sc["guess_r"], sc["guess_t"], True, 100, 8.0, 0.99)
        assert g_ok and np.array_equal(ginl, inl)
        assert np.abs(gr - r).max() < 1e-9 and np.abs(gt - t).max() < 1e-8


def test_pnp_ransac_no_model_and_errors(ctx, pmv):
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (50, 3)).astype(np.float32); X[:, 2] += 10
    uv = rng.uniform(0, 1000, (50, 2)).astype(np.float32)            # unrelated observations: no consensus
    K = np.array([[700., 0, 600], [0, 700., 180], [0, 0, 1]])
    ok, r, t, inl = ctx.pnp_ransac(X, uv, K, np.array([0.1, 0.2, 0.3]), np.array([1., 2., 3.]), True, 100, 1.0, 0.99)
    cok, cr, ct, cinl = cv2.solvePnPRansac(X, uv, K, None, np.array([[0.1], [0.2], [0.3]]), np.array([[1.], [2.], [3.]]), True, 100, 1.0, 0.99)
    assert ok == bool(cok)
    if not ok:
        assert np.array_equal(r, [0.1, 0.2, 0.3]) and np.array_equal(t, [1., 2., 3.]) and len(inl) == 0     # pose untouched
    with pytest.raises(pmv.PmvError):
        ctx.pnp_ransac(X[:4], uv[:4], K)
