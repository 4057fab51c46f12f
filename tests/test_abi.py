"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every declared symbol."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(sceneprep_lib):
    from garden_b200.binding import SYMBOLS, load_library
    lib = load_library()
    header = (ROOT / "include" / "garden_sceneprep.h").read_text()
    declared = set(re.findall(r"\b(gsp_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/garden_sceneprep.h but not exported"
    assert declared == set(SYMBOLS), f"binding and header disagree: {declared ^ set(SYMBOLS)}"
    assert b"sm_100a" in lib.gsp_version()


def test_view_setup_helpers_match_oracle(sceneprep_lib, oracle_built):
    """gsp_frustum_planes / gsp_view_from_viewproj (host code in the product library, no device needed) against the oracle's
    Frustum(viewProj) restatement, which tests/test_oracle_vs_ref.py pins against the reference's math library."""
    import numpy as np
    import reflib
    from garden_b200 import views as V
    from garden_b200.binding import load_library
    from garden_b200.layout import VIEW_DTYPE
    lib = load_library()
    o = reflib.Oracle()
    views, vps = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    rng = np.random.default_rng(2)
    mats = [np.asarray(vp, dtype=np.float32).reshape(16) for vp in vps] + [rng.standard_normal(16).astype(np.float32) * 7 for _ in range(50)]
    for i, m in enumerate(mats):
        got = np.zeros((6, 4), np.float32)
        lib.gsp_frustum_planes(m.ctypes.data, got.ctypes.data)
        want = o.frustum_planes(m)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"matrix {i}"
    # the filled gsp_view equals what the Python view builder packs for the same matrix
    for v in range(views.size):
        out = np.zeros(1, dtype=VIEW_DTYPE)
        off = np.ascontiguousarray(views[v]["cameraOffset"], dtype=np.float32)
        assert lib.gsp_view_from_viewproj(mats[v].ctypes.data, off.ctypes.data, int(views[v]["shadowPass"]), out.ctypes.data) == 0
        assert out[0]["planeCount"] == 6 and out[0]["uiPlaneCount"] == 0 and out[0]["shadowPass"] == views[v]["shadowPass"]
        assert np.array_equal(out[0]["cameraOffset"], off)
        assert np.array_equal(out[0]["planes"].view(np.uint32), o.frustum_planes(mats[v]).view(np.uint32))


def test_no_cpu_fallback(sceneprep_lib, monkeypatch, tmp_path):
    """The product path fails loudly: without a CUDA device gsp_create reports GSP_ERR_CUDA (no CPU fallback), and without the
    built library the binding raises instead of routing anywhere else. Nothing under garden_b200/ loads anything from oracle/."""
    import garden_b200.binding as B
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        try:
            B.ScenePrep(0)
            raise AssertionError("gsp_create succeeded without a CUDA device")
        except B.ScenePrepError as e:
            assert e.code == B.GSP_ERR_CUDA and "no CPU fallback" in str(e)
    monkeypatch.setattr(B, "LIB_PATH", tmp_path / "missing.so")
    monkeypatch.setattr(B, "_lib", None)
    try:
        B.load_library()
        raised = False
    except FileNotFoundError as e:
        raised = "no CPU fallback" in str(e)
    assert raised, "a missing libgarden_sceneprep.so must raise"
    # the product never imports, links or opens the checkers
    for path in (ROOT / "garden_b200").rglob("*"):
        if path.suffix not in (".py", ".cu", ".cuh", ".h"):
            continue
        for ln in path.read_text().splitlines():
            low = ln.lower()
            if "oracle" in low or "reflib" in low:
                assert not any(tok in low for tok in ("import ", "#include", "cdll", "dlopen")), f"{path}: {ln.strip()}"


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 and as C++ with no other include (plain pointers and sizes only)."""
    import subprocess
    src = tmp_path / "use.c"
    src.write_text('#include "garden_sceneprep.h"\nint main(void) { gsp_view v; gsp_record r; (void)v; (void)r; return sizeof(gsp_record) == 64 ? 0 : 1; }\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        res = subprocess.run([cc, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", f"-I{ROOT / 'include'}", "-x",
                              "c" if cc == "gcc" else "c++", str(src)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    exe = tmp_path / "use"
    subprocess.run(["gcc", "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0, "gsp_record must be 64 bytes (UnsortedMesh / SortedMesh, mesh.hpp:191-205)"


def test_cpp_host_links_the_library(sceneprep_lib, tmp_path):
    """A C++ host (what the reference is) links libgarden_sceneprep.so directly: view setup works without a device, and
    gsp_create either succeeds (GPU box) or reports GSP_ERR_CUDA with a message (here) — never a silent fallback."""
    import subprocess
    src = tmp_path / "host.cpp"
    src.write_text(r'''
#include "garden_sceneprep.h"
#include <cstdio>
#include <cstring>
int main()
{
    float vp[16] = { 1, 0, 0, 0,  0, 2, 0, 0,  0, 0, 0.5f, 1,  3, 4, 5, 1 }, off[4] = { 1, 2, 3, 0 };
    gsp_view view;
    if (gsp_view_from_viewproj(vp, off, 2, &view) != GSP_OK || view.planeCount != 6 || view.shadowPass != 2) return 10;
    // Frustum(viewProj), frustum.hpp:53-60: plane 0 = t.c3 + t.c0 with t = transpose(viewProj)
    if (view.planes[0][0] != vp[3] + vp[0] || view.planes[0][3] != vp[15] + vp[12] || view.planes[4][2] != vp[10]) return 11;
    gsp_context* ctx = nullptr;
    int rc = gsp_create(0, &ctx);
    if (rc == GSP_OK) { std::printf("created\n"); gsp_destroy(ctx); return 0; }
    if (rc != GSP_ERR_CUDA || ctx != nullptr) return 12;
    const char* msg = gsp_last_error(nullptr);
    if (!msg || !std::strstr(msg, "no CPU fallback")) return 13;
    std::printf("no device: %s\n", msg);
    return 0;
}
''')
    exe = tmp_path / "host"
    lib_dir = ROOT / "garden_b200"
    subprocess.run(["g++", "-std=c++17", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{lib_dir}", "-lgarden_sceneprep",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
