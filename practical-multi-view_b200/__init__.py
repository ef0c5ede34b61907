"""practical-multi-view_b200 -- B200 (sm_100a) VO hot path behind the reference's plugin API.

This Python module is only the *test/bench binding* of the product: ``libpmv_cuda.so``
(hand-written CUDA behind the C ABI in ``include/pmv_cuda.h``) plus the C++ adapters in
``host/`` that subclass the reference's ``Base*`` plugin interfaces.  Import as::

    import pmv_b200            # via the repo-root shim (the directory name has a hyphen)

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the
calls raise ``PmvError``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libpmv_cuda.so"
HEADER_PATH = _HERE.parent / "include" / "pmv_cuda.h"

PMV_OK = 0
LK_USE_INITIAL_FLOW = 4
LK_GET_MIN_EIGENVALS = 8


class PmvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pmv error {code}: {msg}")
        self.code = code


_lib = None

_u8p, _i16p, _i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int32)
_f32p, _f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
_vp, _int, _sz, _dbl = C.c_void_p, C.c_int, C.c_size_t, C.c_double

# name -> (restype, argtypes); mirrors include/pmv_cuda.h one to one
_SIGS = {
    "pmv_version": (C.c_char_p, []),
    "pmv_create": (_vp, [_int]),
    "pmv_destroy": (None, [_vp]),
    "pmv_set_stream": (_int, [_vp, _vp]),
    "pmv_sync": (_int, [_vp]),
    "pmv_last_error": (C.c_char_p, [_vp]),
    "pmv_launch_count": (C.c_uint64, [_vp]),
    "pmv_profile_enable": (_int, [_vp, _int]),
    "pmv_profile_collect": (_int, [_vp, _int, _f64p, _i32p]),
    "pmv_probe_fp64": (_int, [_vp, _f64p, _f64p]),
    "pmv_pyr_levels": (_int, [_int] * 5),
    "pmv_pyramid_build": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _vp, _sz, _i32p]),
    "pmv_scharr": (_int, [_vp, _vp, _int, _int, _int, _vp]),
    "pmv_lk_track": (_int, [_vp, _vp, _vp, _int, _int, _int, _vp, _int, _int, _int, _int, _int, _dbl,
                            _int, _dbl, _vp, _vp, _vp]),
    "pmv_lk_track_batched": (_int, [_vp, _vp, _vp, _int, _sz, _int, _int, _int, _vp, _int, _int, _int,
                                    _int, _int, _dbl, _int, _dbl, _vp, _vp, _vp]),
    "pmv_lk_track_batched_dev": (_int, [_vp, _vp, _vp, _int, _sz, _int, _int, _int, _vp, _int, _int, _int,
                                        _int, _int, _dbl, _int, _dbl, _vp, _vp, _vp]),
    "pmv_tracker_create": (_vp, [_vp, _int, _int, _int, _int, _int, _int, _int, _int, _int, C.c_double, C.c_double, _int]),
    "pmv_tracker_destroy": (None, [_vp]),
    "pmv_tracker_init": (_int, [_vp, _vp, _int, _i32p]),
    "pmv_tracker_add_frame": (_int, [_vp, _vp, _int, _i32p, _i32p, _i32p, _vp, _vp, _int]),
    "pmv_tracker_features": (_int, [_vp, _vp, _int, _i32p]),
    "pmv_pnp_ransac": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _vp, _int, _int, C.c_float, C.c_double, _vp, _i32p]),
    "pmv_kitti_parse_poses": (_int, [C.c_char_p, _int, _vp, _vp, _int, _i32p]),
    "pmv_kitti_parse_calibration": (_int, [C.c_char_p, _int, _vp]),
    "pmv_kitti_error_report": (_int, [_vp, _vp, _int, _vp, _vp, _int, _int, C.c_double, _vp, C.c_char_p, _int]),
    "pmv_find_essential_mat": (_int, [_vp, _vp, _vp, _int, _vp, C.c_double, C.c_double, _int, _vp, _vp, _i32p]),
    "pmv_recover_pose": (_int, [_vp, _vp, _vp, _vp, _int, _vp, C.c_double, _vp, _vp, _vp, _vp, _i32p]),
    "pmv_five_point_pose": (_int, [_vp, _vp, _vp, _int, _vp, C.c_double, C.c_double, _int, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _i32p, _i32p]),
    "pmv_min_eigen_val_batched_dev": (_int, [_vp, _vp, _int, _sz, _int, _int, _int, _vp, _vp]),
    "pmv_shitomasi_response_batched_dev": (_int, [_vp, _vp, _int, _sz, _int, _int, _int, _int, _vp, _vp]),
    "pmv_min_eigen_val": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp]),
    "pmv_gftt": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _dbl, _dbl, _int, _int,
                        _vp, _vp, _i32p]),
    "pmv_ba_index_observations": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32p]),
    "pmv_gftt_dev": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _int, _int, _dbl, _dbl, _vp, _vp, _i32p]),
    "pmv_shitomasi_response": (_int, [_vp, _vp, _int, _int, _int, _int, _vp]),
    "pmv_shitomasi": (_int, [_vp, _vp, _int, _int, _int, _int, _dbl, _int, _vp, _vp, _vp, _i32p]),
    "pmv_fast": (_int, [_vp, _vp, _int, _int, _int, _int, _int, _int, _vp, _vp, _vp, _i32p, _i32p]),
    "pmv_ba_eval": (_int, [_vp] * 6 + [_int] * 3 + [_vp, _dbl] + [_vp] * 4),
    "pmv_ba_solve": (_int, [_vp] * 6 + [_int] * 3 + [_vp, _dbl, _int, _vp]),
    "pmv_ba_solve_batched": (_int, [_vp] * 7 + [_int] * 4 + [_vp, _dbl, _int, _vp]),
    "pmv_ba_problem_create": (_vp, [_vp] * 7 + [_int] * 4 + [_vp, _dbl, _int, _int]),
    "pmv_ba_problem_reset": (_int, [_vp, _vp, _vp]),
    "pmv_ba_problem_solve": (_int, [_vp, _int]),
    "pmv_ba_problem_download": (_int, [_vp, _vp, _vp, _vp]),
    "pmv_ba_problem_device_bytes": (_sz, [_vp]),
    "pmv_ba_problem_destroy": (None, [_vp]),
    "pmv_comm_unique_id": (_int, [_vp]),
    "pmv_comm_init": (_int, [_vp, _int, _int, _vp]),
    "pmv_comm_destroy": (_int, [_vp]),
}


class BASummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("iterations", C.c_int),
                ("successful_steps", C.c_int), ("termination", C.c_int), ("final_radius", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def load_library() -> C.CDLL:
    """dlopen libpmv_cuda.so and bind every symbol of the header.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PmvError(-100, f"{LIB_PATH} is not built -- run __graft_entry__.build(); "
                             "there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGS)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def ba_index_observations(obs, cam_idx, pt_idx, Nc, Np, obs_off=None):
    """Host side of BA problem creation (no GPU): the observation list in device order.  Returns a dict with pt_off, cam_off,
    cam, pt, win, obs, cam_obs and route (0 taken in place, 1 sorted window by window, 2 general)."""
    lib = load_library()
    obs = np.ascontiguousarray(obs, np.float64); cam_idx = np.ascontiguousarray(cam_idx, np.int32); pt_idx = np.ascontiguousarray(pt_idx, np.int32)
    No = len(cam_idx)
    W = 1 if obs_off is None else len(obs_off) - 1
    off = None if obs_off is None else np.ascontiguousarray(obs_off, np.int32)
    out = {"pt_off": np.zeros(W * Np + 1, np.int32), "cam_off": np.zeros(W * Nc + 1, np.int32), "cam": np.zeros(No, np.int32),
           "pt": np.zeros(No, np.int32), "win": np.zeros(No, np.int32), "obs": np.zeros((No, 2)), "cam_obs": np.zeros(No, np.int32)}
    route = C.c_int(-1)
    rc = lib.pmv_ba_index_observations(_ptr(obs), _ptr(cam_idx), _ptr(pt_idx), _ptr(off), W, Nc, Np, No, _ptr(out["pt_off"]),
                                       _ptr(out["cam_off"]), _ptr(out["cam"]), _ptr(out["pt"]), _ptr(out["win"]), _ptr(out["obs"]),
                                       _ptr(out["cam_obs"]), C.byref(route))
    if rc != 0:
        raise ValueError(f"pmv_ba_index_observations failed with status {rc}")
    out["route"] = route.value
    return out


def kitti_parse_poses(path, stop=1 << 30):
    """OdometryPipeline::parsePoses (OdometryPipeline.cpp:525-593).  Returns (R (n,3,3), t (n,3)).  Host only."""
    lib = load_library()
    n = C.c_int(0)
    rc = lib.pmv_kitti_parse_poses(str(path).encode(), int(stop), None, None, 0, C.byref(n))
    if rc != 0:
        raise PmvError(rc, "Unable to open pose file")
    R = np.zeros((n.value, 9)); t = np.zeros((n.value, 3))
    lib.pmv_kitti_parse_poses(str(path).encode(), int(stop), _ptr(R), _ptr(t), n.value, C.byref(n))
    return R.reshape(-1, 3, 3), t


def kitti_parse_calibration(path, num_calib=0, K=None):
    """OdometryPipeline::parseCalibration (:595-653).  Returns K (3,3).  Host only."""
    lib = load_library()
    Km = np.zeros(9) if K is None else np.array(K, np.float64).reshape(9).copy()
    rc = lib.pmv_kitti_parse_calibration(str(path).encode(), int(num_calib), _ptr(Km))
    if rc != 0:
        raise PmvError(rc, "Unable to open calibration file")
    return Km.reshape(3, 3)


def kitti_error_report(R, t, gt_R, gt_t, init_offset=0, runtime=0.0):
    """The error report of OdometryPipeline::run (:272-300).  Returns (stats dict, text of error_path).  Host only."""
    lib = load_library()
    R = np.ascontiguousarray(R, np.float64).reshape(-1, 9); t = np.ascontiguousarray(t, np.float64).reshape(-1, 3)
    gR = np.ascontiguousarray(gt_R, np.float64).reshape(-1, 9); gt = np.ascontiguousarray(gt_t, np.float64).reshape(-1, 3)
    stats = np.zeros(8); text = C.create_string_buffer(1024)
    rc = lib.pmv_kitti_error_report(_ptr(R), _ptr(t), len(R), _ptr(gR), _ptr(gt), len(gR), int(init_offset), float(runtime), _ptr(stats),
                                    text, 1024)
    if rc != 0:
        raise PmvError(rc, "pmv_kitti_error_report: bad argument")
    keys = ["R_total", "R_min", "R_max", "R_std", "t_total", "t_min", "t_max", "t_std"]
    return dict(zip(keys, stats.tolist())), text.value.decode()


class Context:
    """Owns one ``pmv_ctx`` (device buffers + a CUDA stream)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        self.h = self.lib.pmv_create(device)
        if not self.h:
            raise PmvError(-101, "pmv_create failed: no usable CUDA device (no CPU fallback)")
        if stream:
            self.lib.pmv_set_stream(self.h, C.c_void_p(stream))

    def close(self):
        if getattr(self, "h", None):
            self.lib.pmv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int):
        if rc != PMV_OK:
            raise PmvError(rc, self.lib.pmv_last_error(self.h).decode())

    def set_stream(self, stream: int | None):
        self._chk(self.lib.pmv_set_stream(self.h, C.c_void_p(stream or 0)))

    def sync(self):
        self._chk(self.lib.pmv_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.pmv_launch_count(self.h))

    PHASES = ("pyramid", "lk", "response", "select", "fast", "ba", "pyr_l0", "p7")

    def profile(self, on: bool):
        self._chk(self.lib.pmv_profile_enable(self.h, int(on)))

    def profile_collect(self):
        """{phase: (ms_sum, groups)} of CUDA-event time since the last collect."""
        ms = (C.c_double * 8)()
        cnt = (C.c_int32 * 8)()
        self._chk(self.lib.pmv_profile_collect(self.h, 8, ms, cnt))
        return {p: (ms[i], cnt[i]) for i, p in enumerate(self.PHASES) if cnt[i]}

    def probe_fp64(self):
        """Measured (DFMA, DMMA) fp64 TFLOP/s of this device."""
        a, b = C.c_double(0), C.c_double(0)
        self._chk(self.lib.pmv_probe_fp64(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ------------------------------------------------------------------ pyramid
    def pyramid_build(self, img: np.ndarray, win=(21, 21), max_level=3):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        rows, cols = img.shape
        L = self.lib.pmv_pyr_levels(rows, cols, win[0], win[1], max_level)
        shapes, r, c = [], rows, cols
        for _ in range(L):
            r, c = (r + 1) // 2, (c + 1) // 2
            shapes.append((r, c))
        out = np.empty(max(1, sum(a * b for a, b in shapes)), np.uint8)
        nl = C.c_int(0)
        self._chk(self.lib.pmv_pyramid_build(self.h, _ptr(img), rows, cols, img.strides[0], win[0], win[1],
                                             max_level, _ptr(out), out.size, C.byref(nl)))
        levels, o = [], 0
        for (r, c) in shapes[:nl.value]:
            levels.append(out[o:o + r * c].reshape(r, c).copy())
            o += r * c
        return levels

    def scharr(self, img: np.ndarray) -> np.ndarray:
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        rows, cols = img.shape
        out = np.empty((rows, cols, 2), np.int16)
        self._chk(self.lib.pmv_scharr(self.h, _ptr(img), rows, cols, img.strides[0], _ptr(out)))
        return out

    # ------------------------------------------------------------------ Lucas-Kanade
    def lk_track(self, prev, nxt, pts, win=(21, 21), max_level=3, max_count=30, eps=0.01, flags=0,
                 min_eig=1e-4, init=None):
        assert prev.dtype == np.uint8 and nxt.dtype == np.uint8 and prev.shape == nxt.shape
        assert prev.strides == nxt.strides and prev.strides[1] == 1
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        n = len(pts)
        out = np.zeros((n, 2), np.float32) if init is None else np.ascontiguousarray(init, np.float32).reshape(-1, 2).copy()
        st = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        rows, cols = prev.shape
        self._chk(self.lib.pmv_lk_track(self.h, _ptr(prev), _ptr(nxt), rows, cols, prev.strides[0], _ptr(pts), n,
                                        win[0], win[1], max_level, max_count, eps, flags, min_eig,
                                        _ptr(out), _ptr(st), _ptr(err)))
        return out, st, err

    def lk_track_batched(self, prev, nxt, pts, win=(21, 21), max_level=3, max_count=30, eps=0.01, flags=0,
                         min_eig=1e-4, out=None):
        """prev/nxt: (B, H, W) u8 C-contiguous host arrays; pts: (B, n, 2) float32."""
        B, rows, cols = prev.shape
        assert prev.flags.c_contiguous and nxt.flags.c_contiguous and pts.dtype == np.float32
        n = pts.shape[1]
        if out is None:
            out = (np.zeros((B, n, 2), np.float32), np.zeros((B, n), np.uint8), np.zeros((B, n), np.float32))
        nx, st, err = out
        self._chk(self.lib.pmv_lk_track_batched(self.h, _ptr(prev), _ptr(nxt), B, rows * cols, rows, cols, cols,
                                                _ptr(pts), n, win[0], win[1], max_level, max_count, eps, flags,
                                                min_eig, _ptr(nx), _ptr(st), _ptr(err)))
        return nx, st, err

    def lk_track_batched_dev(self, d_prev: int, d_nxt: int, B, img_stride, rows, cols, step, d_pts: int, n,
                             d_next: int, d_status: int, d_err: int, win=(21, 21), max_level=3, max_count=30,
                             eps=0.01, flags=0, min_eig=1e-4):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on the context stream."""
        self._chk(self.lib.pmv_lk_track_batched_dev(self.h, _ptr(d_prev), _ptr(d_nxt), B, img_stride, rows, cols,
                                                    step, _ptr(d_pts), n, win[0], win[1], max_level, max_count,
                                                    eps, flags, min_eig, _ptr(d_next), _ptr(d_status), _ptr(d_err)))

    def pnp_ransac(self, obj, img, K, rvec=None, tvec=None, use_guess=True, iterations=100, reproj_err=8.0, confidence=0.99):
        """== cv2.solvePnPRansac(obj, img, K, None, rvec, tvec, use_guess, iterations, reproj_err, confidence).
        Returns (ok, rvec (3,), tvec (3,), inlier indices)."""
        obj = np.ascontiguousarray(obj, np.float32).reshape(-1, 3); img = np.ascontiguousarray(img, np.float32).reshape(-1, 2)
        n = len(obj)
        assert len(img) == n
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        r = np.zeros(3) if rvec is None else np.array(rvec, np.float64).reshape(3).copy()
        t = np.zeros(3) if tvec is None else np.array(tvec, np.float64).reshape(3).copy()
        mask = np.zeros(max(n, 1), np.uint8)
        ni = C.c_int(0)
        self._chk(self.lib.pmv_pnp_ransac(self.h, _ptr(obj), _ptr(img), n, _ptr(K), _ptr(r), _ptr(t), int(use_guess), iterations,
                                          reproj_err, confidence, _ptr(mask), C.byref(ni)))
        return ni.value > 0, r, t, np.nonzero(mask[:n])[0].astype(np.int32)

    @staticmethod
    def _pairs(p1, p2):
        p1 = np.ascontiguousarray(p1, np.float64).reshape(-1, 2); p2 = np.ascontiguousarray(p2, np.float64).reshape(-1, 2)
        assert len(p1) == len(p2)
        return p1, p2

    def find_essential_mat(self, p1, p2, K, prob=0.99, threshold=1.0, max_iters=1000):
        """== cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, prob, threshold, maxIters).  Returns (E (3,3) or None, mask (n,))."""
        p1, p2 = self._pairs(p1, p2)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        E = np.zeros(9); mask = np.zeros(max(len(p1), 1), np.uint8); ni = C.c_int(0)
        self._chk(self.lib.pmv_find_essential_mat(self.h, _ptr(p1), _ptr(p2), len(p1), _ptr(K), prob, threshold, max_iters, _ptr(E),
                                                  _ptr(mask), C.byref(ni)))
        return (E.reshape(3, 3) if ni.value > 0 else None), mask[:len(p1)]

    def recover_pose(self, E, p1, p2, K, distance_thresh=float("inf"), mask=None):
        """== cv2.recoverPose(E, p1, p2, K, distanceThresh, mask).  Returns (n_good, R, t, mask, tri (4, n))."""
        p1, p2 = self._pairs(p1, p2)
        n = len(p1)
        K = np.ascontiguousarray(K, np.float64).reshape(9); E = np.ascontiguousarray(E, np.float64).reshape(9)
        R = np.zeros(9); t = np.zeros(3); tri = np.zeros((4, n)); ng = C.c_int(0)
        m = np.ones(n, np.uint8) if mask is None else (np.asarray(mask).reshape(-1) != 0).astype(np.uint8)
        self._chk(self.lib.pmv_recover_pose(self.h, _ptr(E), _ptr(p1), _ptr(p2), n, _ptr(K), distance_thresh, _ptr(R), _ptr(t), _ptr(m),
                                            _ptr(tri), C.byref(ng)))
        return ng.value, R.reshape(3, 3), t, m, tri

    def five_point_pose(self, p1, p2, K, prob=0.99, threshold=1.0, max_iters=1000, distance_thresh=float("inf")):
        """findEssentialMat + recoverPose in one launch (OpenCVFivePointTri.cpp:25-27).
        Returns dict(E, R, t, ransac_mask, mask, tri, n_inliers, n_good); E is None when no model was found."""
        p1, p2 = self._pairs(p1, p2)
        n = len(p1)
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        E = np.zeros(9); R = np.zeros(9); t = np.zeros(3); tri = np.zeros((4, n))
        rm = np.zeros(max(n, 1), np.uint8); m = np.zeros(max(n, 1), np.uint8); ni = C.c_int(0); ng = C.c_int(0)
        self._chk(self.lib.pmv_five_point_pose(self.h, _ptr(p1), _ptr(p2), n, _ptr(K), prob, threshold, max_iters, distance_thresh,
                                               _ptr(E), _ptr(R), _ptr(t), _ptr(rm), _ptr(m), _ptr(tri), C.byref(ni), C.byref(ng)))
        ok = ni.value > 0
        return dict(E=E.reshape(3, 3) if ok else None, R=R.reshape(3, 3), t=t, ransac_mask=rm[:n], mask=m[:n], tri=tri,
                    n_inliers=ni.value, n_good=ng.value)

    def tracker(self, rows, cols, win=(32, 32), max_level=4, capacity=4096, min_tracked=400, tracked_tol=150, grid=255,
                quality=0.01, min_dist=5.0, neighbor_dist=5):
        """Device-resident front end (pmv_tracker_*): OdometryPipeline::addFrame without the host round trips."""
        return Tracker(self, rows, cols, win, max_level, capacity, min_tracked, tracked_tol, grid, quality, min_dist, neighbor_dist)

    # ------------------------------------------------------------------ corner detectors
    @staticmethod
    def _roi(img, roi):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        R, Cc = img.shape
        return roi if roi is not None else (0, 0, Cc, R)

    def min_eigen_val(self, img, roi=None):
        x, y, w, h = self._roi(img, roi)
        out = np.empty((h, w), np.float32)
        self._chk(self.lib.pmv_min_eigen_val(self.h, _ptr(img), img.shape[0], img.shape[1], img.strides[0],
                                             x, y, w, h, _ptr(out)))
        return out

    def min_eigen_val_batched_dev(self, d_imgs: int, batch, img_stride, rows, cols, step, d_eig: int, d_max: int):
        """Device pointers (ints); asynchronous.  d_eig: batch*rows*cols float32, d_max: batch float32."""
        self._chk(self.lib.pmv_min_eigen_val_batched_dev(self.h, _ptr(d_imgs), batch, img_stride, rows, cols, step, _ptr(d_eig), _ptr(d_max)))

    def shitomasi_response_batched_dev(self, d_imgs: int, batch, img_stride, rows, cols, step, d_R: int, d_max: int, signed_quirk=True):
        self._chk(self.lib.pmv_shitomasi_response_batched_dev(self.h, _ptr(d_imgs), batch, img_stride, rows, cols, step,
                                                              int(signed_quirk), _ptr(d_R), _ptr(d_max)))

    def gftt(self, img, max_corners, quality=0.01, min_dist=5.0, roi=None, block_size=3, ksize=3):
        x, y, w, h = self._roi(img, roi)
        cap = max_corners if max_corners > 0 else w * h
        xy = np.zeros((max(cap, 1), 2), np.float32)
        sc = np.zeros(max(cap, 1), np.float32)
        n = C.c_int(0)
        self._chk(self.lib.pmv_gftt(self.h, _ptr(img), img.shape[0], img.shape[1], img.strides[0], x, y, w, h,
                                    max_corners, quality, min_dist, block_size, ksize, _ptr(xy), _ptr(sc), C.byref(n)))
        return xy[:n.value], sc[:n.value]

    def gftt_dev(self, d_img: int, rows, cols, step, max_corners, d_xy: int, d_score: int = 0, quality=0.01, min_dist=5.0, roi=None):
        """pmv_gftt_dev: resident image (device pointer as int) in, corner list to device buffers; returns the count."""
        x, y, w, h = roi if roi is not None else (0, 0, cols, rows)
        n = C.c_int(0)
        self._chk(self.lib.pmv_gftt_dev(self.h, _ptr(d_img), rows, cols, step, x, y, w, h, max_corners, quality, min_dist,
                                        _ptr(d_xy), _ptr(d_score), C.byref(n)))
        return n.value

    def shitomasi_response(self, img, signed_quirk=True):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        out = np.empty(img.shape, np.float64)
        self._chk(self.lib.pmv_shitomasi_response(self.h, _ptr(img), img.shape[0], img.shape[1], img.strides[0],
                                                  int(signed_quirk), _ptr(out)))
        return out

    def shitomasi(self, img, max_feats, quality=0.4, signed_quirk=True):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        cap = max(max_feats, 1)
        col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32); sc = np.zeros(cap, np.float64)
        n = C.c_int(0)
        self._chk(self.lib.pmv_shitomasi(self.h, _ptr(img), img.shape[0], img.shape[1], img.strides[0], max_feats,
                                         quality, int(signed_quirk), _ptr(col), _ptr(row), _ptr(sc), C.byref(n)))
        return col[:n.value], row[:n.value], sc[:n.value]

    def fast(self, img, threshold=10, nonmax=True, max_feats=None):
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        cap = img.size if max_feats is None else max_feats
        col = np.zeros(max(cap, 1), np.int32); row = np.zeros(max(cap, 1), np.int32); sc = np.zeros(max(cap, 1), np.float32)
        n, tot = C.c_int(0), C.c_int(0)
        self._chk(self.lib.pmv_fast(self.h, _ptr(img), img.shape[0], img.shape[1], img.strides[0], threshold,
                                    int(nonmax), cap, _ptr(col), _ptr(row), _ptr(sc), C.byref(n), C.byref(tot)))
        return col[:n.value], row[:n.value], sc[:n.value], tot.value

    # ------------------------------------------------------------------ bundle adjustment
    @staticmethod
    def _ba_args(poses, points, obs, cam_idx, pt_idx, K):
        return (np.array(poses, np.float64, order="C"), np.array(points, np.float64, order="C"),
                np.ascontiguousarray(obs, np.float64), np.ascontiguousarray(cam_idx, np.int32),
                np.ascontiguousarray(pt_idx, np.int32), np.ascontiguousarray(K, np.float64).ravel())

    def ba_eval(self, poses, points, obs, cam_idx, pt_idx, K, huber_delta=1.0):
        poses, points, obs, cam_idx, pt_idx, K = self._ba_args(poses, points, obs, cam_idx, pt_idx, K)
        n = len(cam_idx)
        r = np.zeros((n, 2)); Jc = np.zeros((n, 2, 6)); Jp = np.zeros((n, 2, 3)); cost = np.zeros(1)
        self._chk(self.lib.pmv_ba_eval(self.h, _ptr(poses), _ptr(points), _ptr(obs), _ptr(cam_idx), _ptr(pt_idx),
                                       len(poses), len(points), n, _ptr(K), huber_delta, _ptr(r), _ptr(Jc), _ptr(Jp),
                                       _ptr(cost)))
        return r, Jc, Jp, float(cost[0])

    def ba_solve(self, poses, points, obs, cam_idx, pt_idx, K, huber_delta=1.0, max_iters=5):
        poses, points, obs, cam_idx, pt_idx, K = self._ba_args(poses, points, obs, cam_idx, pt_idx, K)
        s = BASummary()
        self._chk(self.lib.pmv_ba_solve(self.h, _ptr(poses), _ptr(points), _ptr(obs), _ptr(cam_idx), _ptr(pt_idx),
                                        len(poses), len(points), len(cam_idx), _ptr(K), huber_delta, max_iters,
                                        C.byref(s)))
        return poses, points, s.as_dict()

    def ba_solve_batched(self, poses, points, obs, cam_idx, pt_idx, obs_off, K, huber_delta=1.0, max_iters=5):
        poses, points, obs, cam_idx, pt_idx, K = self._ba_args(poses, points, obs, cam_idx, pt_idx, K)
        obs_off = np.ascontiguousarray(obs_off, np.int32)
        W, Nc, _ = poses.shape
        sums = (BASummary * W)()
        self._chk(self.lib.pmv_ba_solve_batched(self.h, _ptr(poses), _ptr(points), _ptr(obs), _ptr(cam_idx),
                                                _ptr(pt_idx), _ptr(obs_off), W, Nc, points.shape[1], len(cam_idx),
                                                _ptr(K), huber_delta, max_iters, sums))
        return poses, points, [x.as_dict() for x in sums]

    def ba_problem(self, poses, points, obs, cam_idx, pt_idx, K, huber_delta=1.0, obs_off=None, rank=0, nranks=1):
        return BAProblem(self, poses, points, obs, cam_idx, pt_idx, K, huber_delta, obs_off, rank, nranks)

    # ------------------------------------------------------------------ NCCL (sharded BA)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.pmv_comm_unique_id(buf)
        if rc != PMV_OK:
            raise PmvError(rc, "pmv_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        self._chk(self.lib.pmv_comm_init(self.h, nranks, rank, C.c_char_p(uid)))

    def comm_destroy(self):
        self._chk(self.lib.pmv_comm_destroy(self.h))


class Tracker:
    """pmv_tracker handle: frames in, (features, n_tracked, extracted) out; pyramids and tracks stay on the device."""

    def __init__(self, ctx, rows, cols, win, max_level, capacity, min_tracked, tracked_tol, grid, quality, min_dist, neighbor_dist):
        self.ctx, self.rows, self.cols, self.cap = ctx, rows, cols, capacity
        self.h = ctx.lib.pmv_tracker_create(ctx.h, rows, cols, win[0], win[1], max_level, capacity, min_tracked, tracked_tol, grid,
                                            quality, min_dist, neighbor_dist)
        if not self.h:
            raise PmvError(-1, ctx.lib.pmv_last_error(ctx.h).decode())
        self._xy = np.zeros((capacity, 2), np.int32)
        self._pi = np.zeros(capacity, np.int32)

    def _img(self, img):
        assert img.dtype == np.uint8 and img.shape == (self.rows, self.cols) and img.strides[1] == 1
        return img

    def init(self, img):
        n = C.c_int(0)
        self.ctx._chk(self.ctx.lib.pmv_tracker_init(self.h, _ptr(self._img(img)), img.strides[0], C.byref(n)))
        return self.features()

    def features(self):
        n = C.c_int(0)
        self.ctx._chk(self.ctx.lib.pmv_tracker_features(self.h, _ptr(self._xy), self.cap, C.byref(n)))
        return self._xy[:n.value].copy()

    def add_frame(self, img):
        """Returns (features (n, 2) int32 [column, row], prev_index (n,), n_tracked, extracted)."""
        nt, nf, ex = C.c_int(0), C.c_int(0), C.c_int(0)
        self.ctx._chk(self.ctx.lib.pmv_tracker_add_frame(self.h, _ptr(self._img(img)), img.strides[0], C.byref(nt), C.byref(nf),
                                                         C.byref(ex), _ptr(self._xy), _ptr(self._pi), self.cap))
        n = min(nf.value, self.cap)
        return self._xy[:n].copy(), self._pi[:n].copy(), nt.value, bool(ex.value)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.pmv_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BAProblem:
    """Device-resident bundle-adjustment problem (pmv_ba_problem_*)."""

    def __init__(self, ctx, poses, points, obs, cam_idx, pt_idx, K, huber_delta, obs_off, rank, nranks):
        self.ctx = ctx
        poses, points, obs, cam_idx, pt_idx, K = Context._ba_args(poses, points, obs, cam_idx, pt_idx, K)
        if poses.ndim == 2:
            poses, points = poses[None], points[None]
        self.W, self.Nc, _ = poses.shape
        self.Np = points.shape[1]
        off = None if obs_off is None else np.ascontiguousarray(obs_off, np.int32)
        self.h = ctx.lib.pmv_ba_problem_create(ctx.h, _ptr(poses), _ptr(points), _ptr(obs), _ptr(cam_idx), _ptr(pt_idx),
                                               _ptr(off), self.W, self.Nc, self.Np, len(cam_idx), _ptr(K), huber_delta,
                                               rank, nranks)
        if not self.h:
            raise PmvError(-1, ctx.lib.pmv_last_error(ctx.h).decode())

    def reset(self):
        self.ctx._chk(self.ctx.lib.pmv_ba_problem_reset(self.h, None, None))

    def solve(self, max_iters):
        self.ctx._chk(self.ctx.lib.pmv_ba_problem_solve(self.h, max_iters))

    def download(self):
        poses = np.zeros((self.W, self.Nc, 6)); points = np.zeros((self.W, self.Np, 3))
        sums = (BASummary * self.W)()
        self.ctx._chk(self.ctx.lib.pmv_ba_problem_download(self.h, _ptr(poses), _ptr(points), sums))
        return poses, points, [x.as_dict() for x in sums]

    @property
    def device_bytes(self):
        return int(self.ctx.lib.pmv_ba_problem_device_bytes(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.pmv_ba_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
