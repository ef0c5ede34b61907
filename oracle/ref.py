"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of ``oracle/_ref/libpmv_ref.so``.

That library is the REFERENCE's own code (its .cpp files compiled unchanged from /root/reference, see
``oracle/ref_build.py``) plus, side by side, the product's drop-in adapters (``impl=1``).  The third-party kernels
the reference delegates to are reached through hooks that this module points at the real ``cv2`` wheel, so
``impl=0`` = reference source + real OpenCV arithmetic (+ the restated Ceres minimiser for the BA)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import ref_build

_u8p, _f32p, _f64p, _i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)

_LK = C.CFUNCTYPE(C.c_int, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _u8p, _f32p)
_GFTT = C.CFUNCTYPE(C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                    C.c_int, _f32p, _i32p)
_FAST = C.CFUNCTYPE(C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _i32p)
_BLUR = C.CFUNCTYPE(C.c_int, _f64p, C.c_int, C.c_int, C.c_int, _f64p)
_PNP = C.CFUNCTYPE(C.c_int, _f32p, _f32p, C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_float, C.c_double, _i32p, _i32p)


_FINDE = C.CFUNCTYPE(C.c_int, _f64p, _f64p, C.c_int, _f64p, C.c_int, C.c_double, C.c_double, _f64p, _u8p)
_RECOVER = C.CFUNCTYPE(C.c_int, _f64p, _f64p, _f64p, C.c_int, _f64p, C.c_double, _f64p, _f64p, _u8p, _f64p)


class Hooks(C.Structure):
    _fields_ = [("lk", _LK), ("gftt", _GFTT), ("fast", _FAST), ("blur3", _BLUR), ("pnp_ransac", _PNP), ("find_essential", _FINDE),
                ("recover_pose", _RECOVER)]


def _img(ptr, rows, cols, step):
    a = np.ctypeslib.as_array(ptr, shape=(rows * step,))
    return np.lib.stride_tricks.as_strided(a, (rows, cols), (step, 1))


def _cv2_hooks():
    import cv2

    def lk(prev, nxt, rows, cols, sp, sn, pts, n, ww, wh, ml, out, st, err):
        try:
            a = np.ascontiguousarray(_img(prev, rows, cols, sp)); b = np.ascontiguousarray(_img(nxt, rows, cols, sn))
            p = np.ctypeslib.as_array(pts, shape=(n, 2)).copy().reshape(-1, 1, 2)
            nx, s, e = cv2.calcOpticalFlowPyrLK(a, b, p, None, winSize=(ww, wh), maxLevel=ml)
            np.ctypeslib.as_array(out, shape=(n, 2))[:] = nx.reshape(-1, 2)
            np.ctypeslib.as_array(st, shape=(n,))[:] = s.ravel()
            np.ctypeslib.as_array(err, shape=(n,))[:] = e.ravel()
            return 0
        except Exception:
            return -1

    def gftt(base, fr, fc, step, x, y, w, h, maxc, q, md, cap, xy, n_out):
        try:
            full = _img(base, fr, fc, step)
            if (x, y, w, h) == (0, 0, fc, fr):
                c = cv2.goodFeaturesToTrack(np.ascontiguousarray(full), maxc, q, md)
                c = np.zeros((0, 2), np.float32) if c is None else c.reshape(-1, 2)
            else:
                # a numpy view loses cv::Mat ROI parentage (SURVEY 8c caveat): C++ Sobel on a sub-Mat reads the parent
                # beyond the ROI edge.  Real cv2 kernels for the response (Sobel on the parent, crop, boxFilter on the
                # crop), the pinned oracle for the selection stage (cv2 exposes no entry point for it).
                from . import gftt_select
                f = np.ascontiguousarray(full)
                sc = 1.0 / (4 * 3 * 255.0)
                dx = cv2.Sobel(f, cv2.CV_32F, 1, 0, ksize=3, scale=sc)[y:y + h, x:x + w]
                dy = cv2.Sobel(f, cv2.CV_32F, 0, 1, ksize=3, scale=sc)[y:y + h, x:x + w]
                cov = [cv2.boxFilter(np.ascontiguousarray(m), cv2.CV_32F, (3, 3), normalize=False) for m in (dx * dx, dx * dy, dy * dy)]
                a, b, cc = cov[0] * 0.5, cov[1], cov[2] * 0.5
                eig = (a + cc) - np.sqrt((a - cc) * (a - cc) + b * b)
                c, _ = gftt_select(eig.astype(np.float32), maxc, q, md)
            n = min(len(c), cap)
            np.ctypeslib.as_array(xy, shape=(max(cap, 1), 2))[:n] = c[:n]
            n_out[0] = n
            return 0
        except Exception:
            return -1

    def fast(img, rows, cols, step, thr, nonmax, cap, xy, resp, n_out):
        try:
            d = cv2.FastFeatureDetector_create(threshold=thr, nonmaxSuppression=bool(nonmax))
            kp = d.detect(np.ascontiguousarray(_img(img, rows, cols, step)))
            n = min(len(kp), cap)
            o = np.ctypeslib.as_array(xy, shape=(max(cap, 1), 2)); r = np.ctypeslib.as_array(resp, shape=(max(cap, 1),))
            for i in range(n):
                o[i] = kp[i].pt; r[i] = kp[i].response
            n_out[0] = n
            return 0
        except Exception:
            return -1

    def blur3(src, rows, cols, cn, dst):
        try:
            a = np.ctypeslib.as_array(src, shape=(rows, cols, cn)).copy()
            np.ctypeslib.as_array(dst, shape=(rows, cols, cn))[:] = cv2.blur(a, (3, 3)).reshape(rows, cols, cn)
            return 0
        except Exception:
            return -1

    def pnp(obj, img, n, K, rv, tv, guess, iters, reproj, conf, inl, ninl):
        try:
            o = np.ctypeslib.as_array(obj, shape=(n, 3)).copy(); p = np.ctypeslib.as_array(img, shape=(n, 2)).copy()
            Km = np.ctypeslib.as_array(K, shape=(3, 3)).copy()
            r = np.ctypeslib.as_array(rv, shape=(3,)); t = np.ctypeslib.as_array(tv, shape=(3,))
            ok, r2, t2, il = cv2.solvePnPRansac(o, p, Km, None, r.copy().reshape(3, 1), t.copy().reshape(3, 1), bool(guess), iters, reproj, conf)
            r[:] = r2.ravel(); t[:] = t2.ravel()
            il = np.zeros(0, np.int32) if il is None else il.ravel()
            np.ctypeslib.as_array(inl, shape=(max(n, 1),))[:len(il)] = il
            ninl[0] = len(il)
            return int(ok)
        except Exception:
            return 0

    def find_essential(p1, p2, n, K, method, prob, thr, E, mask):
        try:
            a = np.ctypeslib.as_array(p1, shape=(n, 2)).copy(); b = np.ctypeslib.as_array(p2, shape=(n, 2)).copy()
            Km = np.ctypeslib.as_array(K, shape=(3, 3)).copy()
            Em, m = cv2.findEssentialMat(a, b, Km, method, prob, thr)
            np.ctypeslib.as_array(mask, shape=(n,))[:] = 0 if m is None else m.ravel()
            if Em is None:
                return 0
            np.ctypeslib.as_array(E, shape=(90,))[:Em.size] = Em.ravel()
            return Em.shape[0]
        except Exception:
            return -1

    def recover_pose(E, p1, p2, n, K, dist, R, t, mask, tri):
        try:
            a = np.ctypeslib.as_array(p1, shape=(n, 2)).copy(); b = np.ctypeslib.as_array(p2, shape=(n, 2)).copy()
            Km = np.ctypeslib.as_array(K, shape=(3, 3)).copy(); Em = np.ctypeslib.as_array(E, shape=(3, 3)).copy()
            m = np.ctypeslib.as_array(mask, shape=(n,))
            good, Rm, tm, m2, tr = cv2.recoverPose(Em, a, b, Km, distanceThresh=dist, mask=m.copy().reshape(n, 1))
            np.ctypeslib.as_array(R, shape=(9,))[:] = Rm.ravel(); np.ctypeslib.as_array(t, shape=(3,))[:] = tm.ravel()
            m[:] = m2.ravel(); np.ctypeslib.as_array(tri, shape=(4, n))[:] = tr
            return int(good)
        except Exception:
            return -1

    return Hooks(_LK(lk), _GFTT(gftt), _FAST(fast), _BLUR(blur3), _PNP(pnp), _FINDE(find_essential), _RECOVER(recover_pose))


_lib = None
_hooks_keepalive = None


def available() -> bool:
    return ref_build.build() is not None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        p = ref_build.build()
        if p is None:
            raise RuntimeError("oracle/_ref/libpmv_ref.so is neither present nor buildable here (needs /root/reference + g++)")
        _lib = C.CDLL(str(p))
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_sources.restype = C.c_char_p
        use_cv2_hooks(True)
    return _lib


def use_cv2_hooks(on: bool = True):
    """on: the shim's cv:: kernels are the real cv2 ones; off: the plain-C oracle / 9-term blur."""
    global _hooks_keepalive
    l = lib() if _lib is None else _lib
    if on:
        _hooks_keepalive = _cv2_hooks()
        l.ref_set_hooks(C.byref(_hooks_keepalive))
    else:
        l.ref_set_hooks(None)
        _hooks_keepalive = None


def _chk(rc):
    if rc < 0:
        raise RuntimeError("libpmv_ref: " + lib().ref_last_error().decode())
    return rc


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def shitomasi_response(img, impl=0):
    img = np.ascontiguousarray(img, np.uint8); r, c = img.shape
    out = np.empty((r, c), np.float64)
    _chk(lib().ref_shitomasi_response(impl, _p(img, C.c_uint8), r, c, c, _p(out, C.c_double)))
    return out


def frame_planes(img):
    """(grad_x, grad_y, harris[rows, cols, 3]) of the reference Frame (Frame.cpp:58-86, 119-138)."""
    img = np.ascontiguousarray(img, np.uint8); r, c = img.shape
    gx = np.empty((r, c)); gy = np.empty((r, c)); hm = np.empty((r, c, 3))
    _chk(lib().ref_frame_planes(_p(img, C.c_uint8), r, c, c, _p(gx, C.c_double), _p(gy, C.c_double), _p(hm, C.c_double)))
    return gx, gy, hm


EXTRACTORS = {"shitomasi": 0, "gftt": 1, "fast": 2}


def extract(which, img, max_feats, roi=None, impl=0):
    """extractor->extractFeatures(frame or ROI view, max): (col, row, score, tracked) arrays."""
    img = np.ascontiguousarray(img, np.uint8); r, c = img.shape
    cap = max(max_feats, 1)
    col = np.zeros(cap, np.int32); row = np.zeros(cap, np.int32); sc = np.zeros(cap, np.float64); tr = np.zeros(cap, np.int32)
    roi_a = None if roi is None else np.asarray(roi, np.int32)
    n = _chk(lib().ref_extract(EXTRACTORS[which], impl, _p(img, C.c_uint8), r, c, c, None if roi_a is None else _p(roi_a, C.c_int),
                               max_feats, _p(col, C.c_int), _p(row, C.c_int), _p(sc, C.c_double), _p(tr, C.c_int), cap))
    return col[:n], row[:n], sc[:n], tr[:n]


def match(prev, nxt, feats_cr, impl=0):
    """matcher->matchFeatures(src, next): correspondences (n, 4) = (src col, src row, next col, next row), sorted; next.map size."""
    prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8); r, c = prev.shape
    f = np.ascontiguousarray(feats_cr, np.int32).reshape(-1, 2); n = len(f)
    corr = np.zeros((max(n, 1), 4), np.int32); sz = C.c_int(0)
    k = _chk(lib().ref_match(impl, _p(prev, C.c_uint8), _p(nxt, C.c_uint8), r, c, c, _p(f, C.c_int), n, _p(corr, C.c_int), max(n, 1), C.byref(sz)))
    corr = corr[:k]
    return corr[np.lexsort(corr.T[::-1])], sz.value


def residual(poses, points, obs, K, cam_idx=None, pt_idx=None, jac=True):
    """ProjectionResidual::Create(p2d, camera)->Evaluate for n observations: r (n,2), J_pose (n,2,6), J_pt (n,2,3)."""
    poses = np.ascontiguousarray(poses, np.float64); points = np.ascontiguousarray(points, np.float64)
    obs = np.ascontiguousarray(obs, np.float64).reshape(-1, 2); K = np.ascontiguousarray(K, np.float64).ravel()
    n = len(obs)
    ci = None if cam_idx is None else np.ascontiguousarray(cam_idx, np.int32)
    pi = None if pt_idx is None else np.ascontiguousarray(pt_idx, np.int32)
    r = np.zeros((n, 2)); Jc = np.zeros((n, 2, 6)); Jp = np.zeros((n, 2, 3))
    _chk(lib().ref_residual(_p(poses, C.c_double), _p(points, C.c_double), _p(obs, C.c_double), _p(K, C.c_double), n,
                            None if ci is None else _p(ci, C.c_int), None if pi is None else _p(pi, C.c_int),
                            _p(r, C.c_double), _p(Jc, C.c_double) if jac else None, _p(Jp, C.c_double) if jac else None))
    return r, Jc, Jp


def ba_apply(R, t, points, obs_frame, obs_point, obs_col, obs_row, K, bundle_size, ba_iterations, apply_frame, impl=0):
    """optimizer->apply(frame) on a synthetic OdometryPipeline state.  Returns (R, t, points float32, summary dict)."""
    R = np.array(R, np.float64, order="C").reshape(-1, 9); t = np.array(t, np.float64, order="C").reshape(-1, 3)
    pts = np.array(points, np.float32, order="C").reshape(-1, 3); K = np.ascontiguousarray(K, np.float64).ravel()
    of, op, oc, orow = (np.ascontiguousarray(a, np.int32) for a in (obs_frame, obs_point, obs_col, obs_row))
    s = np.zeros(4)
    _chk(lib().ref_ba_apply(impl, len(R), bundle_size, ba_iterations, _p(K, C.c_double), _p(R, C.c_double), _p(t, C.c_double), len(pts),
                            _p(pts, C.c_float), len(of), _p(of, C.c_int), _p(op, C.c_int), _p(oc, C.c_int), _p(orow, C.c_int), apply_frame,
                            _p(s, C.c_double)))
    return R.reshape(-1, 3, 3), t, pts, {"initial_cost": s[0], "final_cost": s[1], "iterations": int(s[2]), "successful_steps": int(s[3])}


def pnp_solve(K, R1, t1, points, src_cr, next_cr, R_guess, t_guess, impl=0):
    """pnpsolver->solvePnP(src, next, R, t) on a synthetic pipeline state (impl 0: OpenCVEPnPSolver compiled from the
    reference, cv2.solvePnPRansac behind the shim; impl 1: GpuEPnPSolver).  Returns (R, t, kept mask, len(next.map))."""
    K = np.ascontiguousarray(K, np.float64).ravel(); R1 = np.ascontiguousarray(R1, np.float64).ravel(); t1 = np.ascontiguousarray(t1, np.float64).ravel()
    pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    s_cr = np.ascontiguousarray(src_cr, np.int32).reshape(-1, 2); n_cr = np.ascontiguousarray(next_cr, np.int32).reshape(-1, 2)
    R = np.array(R_guess, np.float64, order="C").reshape(9).copy(); t = np.array(t_guess, np.float64, order="C").reshape(3).copy()
    kept = np.zeros(len(pts), np.int32)
    n_next = lib().ref_pnp_solve(impl, _p(K, C.c_double), _p(R1, C.c_double), _p(t1, C.c_double), len(pts), _p(pts, C.c_float), _p(s_cr, C.c_int),
                                 _p(n_cr, C.c_int), _p(R, C.c_double), _p(t, C.c_double), _p(kept, C.c_int))
    _chk(min(n_next, 0))
    return R.reshape(3, 3), t, kept.astype(bool), n_next


def triangulate(K, src_cr, next_cr, gt0, gt1, impl=0):
    """triangulator->triangulate(src, next, R, t) on a synthetic pipeline state (impl 0: OpenCVFivePointTri compiled from
    the reference, cv2.findEssentialMat / recoverPose behind the shim; impl 1: GpuFivePointTri).  Returns
    (R, t, scale, f3d_index per input pair, world points of OdometryPipeline::feats3d)."""
    K = np.ascontiguousarray(K, np.float64).ravel()
    s_cr = np.ascontiguousarray(src_cr, np.int32).reshape(-1, 2); n_cr = np.ascontiguousarray(next_cr, np.int32).reshape(-1, 2)
    g0 = np.ascontiguousarray(gt0, np.float64).ravel(); g1 = np.ascontiguousarray(gt1, np.float64).ravel()
    n = len(s_cr)
    R = np.zeros(9); t = np.zeros(3); scale = C.c_double(0); idx = np.zeros(n, np.int32); pts = np.zeros((n, 3), np.float32)
    k = lib().ref_triangulate(impl, _p(K, C.c_double), n, _p(s_cr, C.c_int), _p(n_cr, C.c_int), _p(g0, C.c_double), _p(g1, C.c_double),
                              _p(R, C.c_double), _p(t, C.c_double), C.byref(scale), _p(idx, C.c_int), _p(pts, C.c_float))
    _chk(min(k, 0))
    return R.reshape(3, 3), t, scale.value, idx, pts[:k]
