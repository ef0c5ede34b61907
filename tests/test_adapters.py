"""CPU: the C++ drop-in adapters (practical-multi-view_b200/host/pmv_adapters.h) must compile against the
REFERENCE's own plugin headers (Base*.h, Frame.h, Feature.h, OdometryPipeline.h).  The image has no C++
OpenCV / dlib / Ceres; oracle/ref_shim/ is a small functional stand-in for the slice the reference uses.  This
test is the syntax check (runs only where /root/reference is mounted); the adapters are also LINKED and RUN
against the reference's own classes through oracle/_ref/libpmv_ref.so (tests/test_ref_pins.py on CPU for the
reference side, tests/test_gpu_ref_plugins.py on the GPU box for reference-vs-adapter)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/include")


@pytest.mark.skipif(not REF.exists() or shutil.which("g++") is None, reason="reference headers / g++ not present")
def test_adapters_compile_against_reference_headers(tmp_path):
    tu = tmp_path / "tu.cpp"
    tu.write_text('#include "pmv_adapters.h"\n'
                  "BaseFeatureMatcher* m() { return new GpuLucasKanadeFM(); }\n"
                  "BaseFeatureExtractor* e1() { return new GpuGoodFeatureExtractor(); }\n"
                  "BaseFeatureExtractor* e2() { return new GpuShiTomasiFeatureExtractor(); }\n"
                  "BaseFeatureExtractor* e3() { return new GpuFASTFeatureExtractor(); }\n"
                  "BaseOptimizer* b(OdometryPipeline* p) { return new GpuBundleAdjustment(p); }\n"
                  "BasePnPSolver* s(OdometryPipeline* p) { return new GpuEPnPSolver(p); }\n")
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-I", str(ROOT / "oracle" / "ref_shim"), "-I", str(REF),
           "-I", str(ROOT / "include"), "-I", str(ROOT / "practical-multi-view_b200" / "host"), str(tu)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_header_is_plain_c(tmp_path):
    """include/pmv_cuda.h must be consumable from C (cgo / ctypes / JNI style bindings)."""
    if shutil.which("gcc") is None:
        pytest.skip("gcc not present")
    tu = tmp_path / "tu.c"
    tu.write_text('#include "pmv_cuda.h"\nint main(void) { return pmv_pyr_levels(376, 1241, 21, 21, 3) == 3 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-pedantic", "-I", str(ROOT / "include"), str(tu)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
