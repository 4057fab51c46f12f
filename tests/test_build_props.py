"""CPU-side checks of what nvcc produced (no device needed): the library carries sm_100a SASS only, the hot kernels keep the
register / shared-memory budget their occupancy is designed around, nothing spills, and the packed-FP32 path is in the SASS."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="CUDA toolkit not on PATH")


def test_library_is_sm_100a_only(sceneprep_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(sceneprep_lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, f"expected sm_100a cubins only, found {archs}"


def test_hot_kernel_budgets(tmp_path):
    """kCull: <= 64 registers (8 blocks x 128 threads per SM), no spills, <= 27.3 KB static shared per block (8 blocks fit
    the 227 KB of an SM with their 1 KB reservations); kPrepass: <= 64 registers, no spills; kSortPass: 3 blocks per SM
    (<= 85 registers); FFMA2 (packed FP32 pairs) present in kCull's SASS."""
    from garden_b200.build import CSRC, NVCC_FLAGS
    info = {}
    for name in ("cull.cu", "sort.cu"):
        cubin = tmp_path / (name + ".cubin")
        res = subprocess.run(["nvcc", *[f for f in NVCC_FLAGS if f not in ("-Xcompiler", "-fPIC")], "-Xptxas", "-v", "-cubin",
                              str(CSRC / name), "-o", str(cubin)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        cur = None
        for ln in res.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\w+)'", ln)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and cur:
                info.setdefault(cur, {})["spill"] = int(m.group(1)) + int(m.group(2))
            m = re.search(r"Used (\d+) registers.*?(\d+) bytes smem", ln)
            if m and cur:
                info.setdefault(cur, {}).update(regs=int(m.group(1)), smem=int(m.group(2)))
    culls = {k: v for k, v in info.items() if "kCull" in k}
    assert len(culls) >= 8, f"expected one kCull instantiation per view count, got {sorted(culls)}"
    for k, v in culls.items():
        assert v["regs"] <= 64 and v["spill"] == 0 and v["smem"] <= (227 * 1024) // 8 - 1024, (k, v)
    pre = {k: v for k, v in info.items() if "kPrepass" in k}
    assert len(pre) >= 8, f"expected one kPrepass instantiation per view count, got {sorted(pre)}"
    for k, v in pre.items():
        assert v["regs"] <= 64 and v["spill"] == 0, (k, v)
    sort = next(v for k, v in info.items() if "kSortPass" in k)
    assert sort["regs"] <= 85 and sort["smem"] <= 48 * 1024, sort
    sass = subprocess.run(["cuobjdump", "-sass", str(tmp_path / "cull.cu.cubin")], capture_output=True, text=True).stdout
    assert "FFMA2" in sass and "FMUL2" in sass, "the chain product is expected to use packed FP32 pairs"
