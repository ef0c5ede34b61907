"""Build libpmv_cuda.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "build"
LIB = HERE / "libpmv_cuda.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3,-fopenmp", "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found: libpmv_cuda.so cannot be built (there is no CPU fallback)")


def _stale(out: Path, deps) -> bool:
    return (not out.exists()) or any(out.stat().st_mtime < Path(d).stat().st_mtime for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = sorted(CSRC.glob("*.cu"))
    hdrs = sorted(CSRC.glob("*.cuh")) + sorted((HERE.parent / "include").glob("*.h"))
    OBJ.mkdir(exist_ok=True)
    cc = nvcc()

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and not _stale(obj, [src] + hdrs):
            return obj, ""
        cmd = [cc] + NVCC_FLAGS + ["-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        (OBJ / (src.stem + ".ptxas.txt")).write_text(r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    if force or _stale(LIB, objs):
        cmd = [cc, "-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-lcudart_static", "-ldl", "-lpthread", "-lrt", "-lgomp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
