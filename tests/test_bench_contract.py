"""The reference arm of bench.py runs on host cores only, so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_json_contract():
    env = dict(os.environ, PMV_BENCH_REF_PAIRS="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "frame_pairs_tracked_per_s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it."""
    import re
    pkg = ROOT / "practical-multi-view_b200"
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.+\s*oracle)", re.M)
    offenders = [str(p.relative_to(ROOT)) for p in pkg.rglob("*.py") if pat.search(p.read_text())]
    offenders += [str(p.relative_to(ROOT)) for p in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h"))
                  if "oracle/" in p.read_text() and "#include" in "".join(l for l in p.read_text().splitlines(True) if "oracle/" in l)]
    assert offenders == []
