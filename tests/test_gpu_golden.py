"""GPU: the CUDA path (through the C ABI) against the golden vectors produced by the reference's own code — draw lists in exact
order, counts, and MeshRenderComponent::isVisible after gsp_writeback_visible. Nothing here needs /root/reference."""
import numpy as np
import pytest

from common import GoldenCase, assert_frames_equal, golden_cases, gpu_frame

pytestmark = pytest.mark.gpu
CASES = golden_cases()


@pytest.mark.parametrize("path", CASES, ids=[p.stem for p in CASES])
def test_cuda_matches_reference_golden(sceneprep_lib, path):
    from garden_b200.binding import ScenePrep
    case = GoldenCase(path)
    sp = ScenePrep(0)
    case.stage(sp)
    sp.run()
    pools_aos = [(p.copy(), int(m[1])) for p, m in zip(case.pools, case.meta)]
    got = gpu_frame(sp, case.views, pools_aos)
    want = case.frames
    for g, w in zip(got, want):
        if "visible" in g:
            for k, m in enumerate(case.meta):
                active = int(m[3]) > 0 and bool(m[4]) and int(m[2]) > 0  # pools the main view skipped keep their bytes
                if active:
                    assert np.array_equal(g["visible"][k], w["visible"][k]), f"{case.name}: isVisible pool {k}"
                else:
                    assert (g["visible"][k] == 0xFF).all()
            del g["visible"]
        w.pop("visible", None)
    assert_frames_equal(got, want, case.render_types, case.name, canonicalise_got=False)
    sp.close()
