"""CPU: the arithmetic of the GPU two-view initialiser (practical-multi-view_b200/csrc/fivept_math.cuh, __host__ __device__)
compiled with g++ and pinned to the OpenCV kernels the reference calls (cv2 4.13: findEssentialMat(RANSAC, 0.99, 1),
recoverPose -- OpenCVFivePointTri.cpp:25-27).  No GPU needed; the device schedule (fivept.cu) is compared with both in
test_gpu_fivept.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = Path(__file__).resolve().parent.parent
INF = float("inf")


def build_host():
    out = ROOT / "tests" / "_build" / "libfivept_host.so"
    out.parent.mkdir(exist_ok=True)
    src = ROOT / "tests" / "fivept_host_harness.cpp"
    hdrs = [ROOT / "practical-multi-view_b200" / "csrc" / h for h in ("fivept_math.cuh", "pnp_math.cuh")]
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", str(out), str(src), "-lm"], check=True)
    lib = C.CDLL(str(out))
    vp = C.c_void_p
    lib.fivept_host_models.restype = C.c_int
    lib.fivept_host_models.argtypes = [vp, vp, vp]
    lib.fivept_host_find_essential.restype = C.c_int
    lib.fivept_host_find_essential.argtypes = [vp, vp, C.c_int, vp, C.c_double, C.c_double, C.c_int, vp, vp, vp]
    lib.fivept_host_recover_pose.restype = C.c_int
    lib.fivept_host_recover_pose.argtypes = [vp, vp, vp, C.c_int, vp, C.c_double, vp, vp, vp, vp]
    return lib


@pytest.fixture(scope="module")
def host():
    return build_host()


def host_find_essential(lib, sc, prob=0.99, thr=1.0, iters=1000):
    p1, p2, K = sc["p1"], sc["p2"], np.ascontiguousarray(sc["K"])
    E = np.zeros(9); m = np.zeros(len(p1), np.uint8); ev = C.c_int(0)
    good = lib.fivept_host_find_essential(p1.ctypes.data, p2.ctypes.data, len(p1), K.ctypes.data, prob, thr, iters, E.ctypes.data,
                                          m.ctypes.data, C.byref(ev))
    return good, E.reshape(3, 3), m, ev.value


def host_recover_pose(lib, E, sc, mask, dist=INF):
    p1, p2, K = sc["p1"], sc["p2"], np.ascontiguousarray(sc["K"])
    E = np.ascontiguousarray(E, np.float64)
    R = np.zeros(9); t = np.zeros(3); m = np.ascontiguousarray(mask, np.uint8).reshape(-1).copy(); tri = np.zeros((4, len(p1)))
    good = lib.fivept_host_recover_pose(E.ctypes.data, p1.ctypes.data, p2.ctypes.data, len(p1), K.ctypes.data, dist, R.ctypes.data,
                                        t.ctypes.data, m.ctypes.data, tri.ctypes.data)
    return good, R.reshape(3, 3), t, m, tri


def test_minimal_sample_models_match_opencv(host):
    """cv2.findEssentialMat on exactly five points returns every model of EMEstimatorCallback::runKernel stacked: same
    count, same ORDER (null-space completion vectors and Durand-Kerner dynamics restated), same values."""
    from harness import twoview_scene
    tight = 0
    for seed in range(200):
        sc = twoview_scene.scene(seed, n=5, outlier_share=0, integer=False)
        K = sc["K"]
        Ecv, _ = cv2.findEssentialMat(sc["p1"], sc["p2"], K, cv2.RANSAC, 0.99, 1.0)
        f = np.array([K[0, 0], K[1, 1]]); c = K[:2, 2]
        x1 = np.ascontiguousarray((sc["p1"] - c) / f); x2 = np.ascontiguousarray((sc["p2"] - c) / f)
        E = np.zeros((10, 9))
        n = host.fivept_host_models(x1.ctypes.data, x2.ctypes.data, E.ctypes.data)
        ncv = 0 if Ecv is None else Ecv.shape[0] // 3
        assert n == ncv, (seed, n, ncv)
        if n == 0:
            continue
        d = np.abs(Ecv.reshape(-1, 9) - E[:n]).max()
        assert d < 2e-3, (seed, d)          # ill-conditioned roots: both sides satisfy the constraints equally well
        tight += d < 1e-7
        h1 = np.c_[x1, np.ones(5)]; h2 = np.c_[x2, np.ones(5)]
        for m in range(n):
            Em = E[m].reshape(3, 3)
            assert np.abs(np.einsum("ij,jk,ik->i", h2, Em, h1)).max() < 1e-12
            assert np.abs(2 * Em @ Em.T @ Em - np.trace(Em @ Em.T) * Em).max() < 1e-3
    assert tight >= 180


def test_find_essential_mat_matches_opencv(host):
    from harness import twoview_scene
    tight = 0
    for seed in range(120):
        sc = twoview_scene.scene(3000 + seed)
        Ecv, mcv = cv2.findEssentialMat(sc["p1"], sc["p2"], sc["K"], cv2.RANSAC, 0.99, 1.0)
        good, E, m, _ = host_find_essential(host, sc)
        assert np.array_equal(m, mcv.ravel()), seed
        assert good == int(mcv.sum())
        d = np.abs(E - Ecv).max()
        assert d < 1e-3, seed               # the winning minimal sample can be ill-conditioned: 1e-16 noise x its condition number
        tight += d < 1e-7
        assert m[sc["outliers"]].mean() < 0.4 < m.mean()      # a displaced point stays an inlier only along its epipolar line
    assert tight >= 100


def test_recover_pose_matches_opencv(host):
    from harness import twoview_scene
    for seed in range(60):
        sc = twoview_scene.scene(4000 + seed)
        Ecv, mcv = cv2.findEssentialMat(sc["p1"], sc["p2"], sc["K"], cv2.RANSAC, 0.99, 1.0)
        ncv, Rcv, tcv, m2cv, tricv = cv2.recoverPose(Ecv, sc["p1"], sc["p2"], sc["K"], distanceThresh=INF, mask=mcv.copy())
        good, R, t, m, tri = host_recover_pose(host, Ecv, sc, mcv)
        assert good == ncv and np.array_equal(m, (m2cv.ravel() != 0).astype(np.uint8))
        assert np.abs(R - Rcv).max() < 1e-12 and np.abs(t - tcv.ravel()).max() < 1e-12
        ok = m != 0
        q = tri[:3, ok] / tri[3, ok]; qcv = tricv[:3, ok] / tricv[3, ok]
        assert (np.abs(q - qcv) / np.abs(qcv).max(0)).max() < 1e-8
        # and the pose is the scene's (up to the noise)
        assert np.abs(R - sc["R"]).max() < 3e-2 and float(t @ sc["t"]) > 0.95
        # a finite distance threshold filters far points the same way
        ncv50, _, _, m50cv, _ = cv2.recoverPose(Ecv, sc["p1"], sc["p2"], sc["K"], distanceThresh=25.0, mask=mcv.copy())
        g50, _, _, m50, _ = host_recover_pose(host, Ecv, sc, mcv, 25.0)
        assert g50 == ncv50 and np.array_equal(m50, (m50cv.ravel() != 0).astype(np.uint8))
