"""oracle/ref_build.py -- TEST INFRASTRUCTURE ONLY: the recipe that builds ``oracle/_ref/libpmv_ref.so``.

The reference's own translation units are compiled **unchanged, from where they lie under /root/reference** (never
copied into this repository) against the functional OpenCV / Ceres / dlib shim in ``oracle/ref_shim/``; the product's
drop-in adapters (``practical-multi-view_b200/host/pmv_adapters.h``) are compiled into the same object against the
same shim and linked to ``libpmv_cuda.so``; ``oracle/ref_harness.cpp`` exposes both through a C ABI.

``oracle/_ref/`` is git-ignored (no reference-derived binary enters the history) but not gpurun-ignored: the prebuilt
.so travels to the GPU box, where /root/reference does not exist.  ``build()`` is a no-op there when the .so is present.
"""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF = Path("/root/reference")
OUT = HERE / "_ref"
LIB = OUT / "libpmv_ref.so"
PKG = ROOT / "practical-multi-view_b200"

REF_SOURCES = ["Feature.cpp", "Feature3D.cpp", "Frame.cpp", "ShiTomasiFeatureExtractor.cpp", "ProjectionResidual.cpp",
               "CeresBundleAdjustment.cpp", "OpenCVGoodFeatureExtractor.cpp", "OpenCVFASTFeatureExtractor.cpp",
               "OpenCVLucasKanadeFM.cpp", "OpenCVEPnPSolver.cpp", "OpenCVFivePointTri.cpp"]
OWN_SOURCES = [HERE / "ref_harness.cpp", HERE / "ref_shim" / "shim_impl.cpp"]
ORACLE_C = sorted(HERE.glob("pmv_oracle_*.c"))


def available() -> bool:
    return LIB.exists()


def can_build() -> bool:
    return REF.exists() and shutil.which("g++") is not None and (PKG / "libpmv_cuda.so").exists()


def _deps():
    d = [REF / s for s in REF_SOURCES] + OWN_SOURCES + ORACLE_C + [PKG / "host" / "pmv_adapters.h", ROOT / "include" / "pmv_cuda.h"]
    d += list((HERE / "ref_shim").rglob("*.h")) + list((HERE / "ref_shim").rglob("*.hpp")) + list((REF / "include").glob("*.h"))
    return d


def build(force: bool = False) -> Path | None:
    """Returns the library path, or None when it neither exists nor can be built here."""
    if not can_build():
        return LIB if LIB.exists() else None
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= p.stat().st_mtime for p in _deps()):
        return LIB
    OUT.mkdir(exist_ok=True)
    inc = ["-I", str(HERE / "ref_shim"), "-I", str(REF / "include"), "-I", str(ROOT / "include"), "-I", str(PKG / "host")]
    objs = []
    for c in ORACLE_C:                      # the plain-C oracle is the fallback behind the cv:: hooks and the LM behind ceres::Solve
        o = OUT / (c.stem + ".o")
        subprocess.run(["gcc", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-c", str(c), "-o", str(o)], check=True)
        objs.append(o)
    for s in [REF / x for x in REF_SOURCES] + OWN_SOURCES:
        o = OUT / (("ref_" if s.parent == REF else "own_") + s.stem + ".o")
        subprocess.run(["g++", "-std=c++11", "-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-w", *inc, "-c", str(s), "-o", str(o)],
                       check=True)
        objs.append(o)
    subprocess.run(["g++", "-shared", "-fopenmp", "-o", str(LIB), *map(str, objs), "-L", str(PKG), "-lpmv_cuda",
                    "-Wl,-rpath,$ORIGIN/../../practical-multi-view_b200", "-lm"], check=True)
    for o in objs:
        o.unlink()
    return LIB


if __name__ == "__main__":
    print(build(force=True))
