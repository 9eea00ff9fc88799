"""In-tree build of libknnsvc_b200.so (nvcc, sm_100a only).

    python -m knn_svc_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libknnsvc_b200.so"
SOURCES = ["capi.cu", "rows.cu", "knn_filter_sm100.cu", "knn_select.cu", "post.cu", "concat_cost_sm100.cu", "weight_fit.cu", "harmonic.cu",
           "pool_ops.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "knnsvc_b200.h"]
    nvcc = _nvcc()
    extra = os.environ.get("KNNSVC_NVCC_EXTRA", "").split()     # e.g. -DKNNSVC_K5_PROFILE (debugging aids)

    def compile_one(src: str):
        s, o = CSRC / src, OBJ / (src + ".o")
        if force or _stale(o, [s] + headers):
            r = subprocess.run([nvcc, *NVCC_FLAGS, *extra, "-c", str(s), "-o", str(o)], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            (OBJ / (src + ".ptxas.txt")).write_text(r.stderr)
            if verbose:
                print(r.stderr)
        return o

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", str(LIB), *map(str, objs)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
