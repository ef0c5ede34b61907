"""CPU: the C-ABI library builds, loads and exports every symbol include/pmv_cuda.h declares."""
import ctypes
import re

import pytest


def _declared_symbols(header_text):
    return sorted(set(re.findall(r"PMV_API\s+[\w\s\*]+?\b(pmv_\w+)\s*\(", header_text)))


def test_header_symbols_exported(pmv):
    import __graft_entry__ as g
    g.build()
    lib = pmv.load_library()
    declared = _declared_symbols(pmv.HEADER_PATH.read_text())
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in pmv_cuda.h but not exported"
    # the python binding table covers the header one to one
    assert sorted(pmv.exported_symbols()) == declared


def test_pyr_levels_host_arithmetic(pmv):
    lib = pmv.load_library()
    # probes quoted in SURVEY.md Appendix A.1
    assert lib.pmv_pyr_levels(376, 1241, 32, 32, 4) == 3
    assert lib.pmv_pyr_levels(376, 1241, 21, 21, 4) == 4
    assert lib.pmv_pyr_levels(376, 1241, 21, 21, 3) == 3
    assert lib.pmv_pyr_levels(2160, 3840, 21, 21, 3) == 3
    assert lib.pmv_pyr_levels(376, 1241, 21, 21, 0) == 0


def test_no_gpu_fails_loudly(pmv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pmv.PmvError):
        pmv.Context(0)


def test_null_context_rejected(pmv):
    lib = pmv.load_library()
    assert lib.pmv_sync(None) != 0
    assert lib.pmv_launch_count(None) == 0
    assert b"null" in lib.pmv_last_error(None)
