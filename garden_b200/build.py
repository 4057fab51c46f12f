"""Builds libgarden_sceneprep.so (the product: hand-written sm_100a kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box. The library has no dependency on
torch or Python; `garden_b200.binding` loads it with ctypes.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "libgarden_sceneprep.so"
SOURCES = ["staging.cu", "cull.cu", "sort.cu", "emit.cu", "merge.cu", "exchange.cu", "viewsetup.cu", "visible.cu", "selftest.cu", "next.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # parity kernels spell every rounding explicitly (__fmul_rn/__fmaf_rn/...); -fmad=false is belt and braces
    "-fmad=false", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libgarden_sceneprep.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [ROOT.parent / "include" / "garden_sceneprep.h"])
    return newest > LIB.stat().st_mtime


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    obj_dir = ROOT.parent / "build"
    obj_dir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        if not (CSRC / src).exists():
            continue
        obj = obj_dir / (src + ".o")
        # GSP_NVCC_EXTRA: extra flags for experiments (e.g. "-DGSP_SORT_BLOCKS_PER_SM=4"); never set in normal builds
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("GSP_NVCC_EXTRA", "").split(), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, proc in procs:
        out, _ = proc.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    tmp = LIB.with_suffix(".so.tmp")
    # (NCCL is resolved with dlopen at run time — exchange.cu — so the library itself only needs libdl)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *objs, "-ldl"]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
