"""CPU oracle for the VO hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``practical-multi-view_b200/``) never does: it fails loudly without its CUDA library.

Two kinds of checker live here:

* ``oracle.cv2_ref`` -- thin calls into the in-image ``cv2`` 4.13.0 wheel, i.e. the
  *actual* third-party OpenCV kernels the reference delegates to
  (``OpenCVLucasKanadeFM.cpp:15``, ``OpenCVGoodFeatureExtractor.cpp:7``,
  ``OpenCVFASTFeatureExtractor.cpp:8``).  ``cpu_baseline.kind == "reference"``.
* ``liborc`` (this module) -- a plain-C restatement (``oracle/*.c``) of the same
  algorithms plus the reference's own ``ShiTomasiFeatureExtractor`` / ``Frame``
  gradient code and the Ceres LM + Schur bundle adjuster (no Ceres in the image:
  BA parity is "unpinned").  ``cpu_baseline.kind == "port"``.

The reference executable itself cannot be built here (needs C++ OpenCV, Ceres, dlib;
none present, no network), so there is no ``oracle/_ref``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = sorted(_HERE.glob("pmv_oracle_*.c"))
_LIB = _HERE / "_build" / "libpmv_oracle.so"


def build(force: bool = False) -> Path:
    """Compile oracle/*.c into oracle/_build/libpmv_oracle.so (gcc, no deps)."""
    _LIB.parent.mkdir(exist_ok=True)
    if not force and _LIB.exists() and all(_LIB.stat().st_mtime >= s.stat().st_mtime for s in _SRC):
        return _LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fvisibility=hidden",
           "-Wall", "-o", str(_LIB)] + [str(s) for s in _SRC] + ["-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_u8, _i16, _i32, _f32, _f64 = C.c_uint8, C.c_int16, C.c_int32, C.c_float, C.c_double


# ----------------------------------------------------------------------------- pyramid / LK
def pyr_down(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    r, c = img.shape
    out = np.empty(((r + 1) // 2, (c + 1) // 2), np.uint8)
    lib().orc_pyr_down(_p(img, _u8), r, c, c, _p(out, _u8), out.shape[1])
    return out


def pyr_levels(rows, cols, win_w, win_h, max_level) -> int:
    return lib().orc_pyr_levels(rows, cols, win_w, win_h, max_level)


def scharr(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    r, c = img.shape
    out = np.empty((r, c, 2), np.int16)
    lib().orc_scharr(_p(img, _u8), r, c, c, _p(out, _i16))
    return out


def lk_track(prev, nxt, pts, win=(21, 21), max_level=3, max_count=30, eps=0.01, flags=0,
             min_eig=1e-4, init=None):
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = len(pts)
    out = np.zeros((n, 2), np.float32) if init is None else np.ascontiguousarray(init, np.float32).copy()
    st = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    fn = lib().orc_lk_track
    fn.argtypes = [C.POINTER(_u8), C.POINTER(_u8), C.c_int, C.c_int, C.c_int, C.POINTER(_f32), C.c_int,
                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                   C.POINTER(_f32), C.POINTER(_u8), C.POINTER(_f32)]
    fn.restype = C.c_int
    r, c = prev.shape
    fn(_p(prev, _u8), _p(nxt, _u8), r, c, c, _p(pts, _f32), n, win[0], win[1], max_level, max_count,
       eps, flags, min_eig, _p(out, _f32), _p(st, _u8), _p(err, _f32))
    return out, st, err
