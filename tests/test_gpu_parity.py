"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit, on seeded scenes."""
import numpy as np
import pytest

from garden_b200 import scenes, views as V
from garden_b200.layout import RT_COLOR, RT_OIT, RT_OPAQUE, RT_TRANSLUCENT, RT_UI, make_views

from common import OracleRun, aos_inputs, compare_gpu_to_oracle

pytestmark = pytest.mark.gpu


def _run_both(scene, views, strides=None, ready=None, draw_ready=None):
    from garden_b200.binding import ScenePrep
    t, pools = aos_inputs(scene, strides)
    rts = [p.render_type for p in scene.pools]
    orun = OracleRun((t, t.dtype.itemsize, t.size), [(m, m.dtype.itemsize, m.size) for m in pools], rts, views,
                     scene.camera_pos, ready=ready, draw_ready=draw_ready)
    sp = ScenePrep(0)
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    sp.set_pool_count(len(pools))
    for k, m in enumerate(pools):
        sp.set_mesh_pool(k, rts[k], m, m.dtype.itemsize, m.size, m.size,
                         True if draw_ready is None else draw_ready[k], None if ready is None else ready[k])
    sp.set_views(views, scene.camera_pos)
    sp.run()
    compare_gpu_to_oracle(sp, orun, rts, views, scene.name)
    return sp, orun, t, pools


def test_c1_flat_single_view(oracle_built, sceneprep_lib):
    scene = scenes.config_scene("C1")
    views, _ = V.perspective_views([(0.4, -0.05)], 1.2, 16 / 9, 0.01)
    sp, orun, _, _ = _run_both(scene, views)
    assert orun.views[0]["unsorted"][0][1] > 500  # the frustum sees a real share of the scene


def test_depth4_camera_and_cascades(oracle_built, sceneprep_lib):
    scene = scenes.config_scene("C2", n=200_000)
    scene.camera_pos = np.array([5.0, 2.0, -3.0], np.float32)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    _run_both(scene, views)


def test_depth8_opaque_and_translucent(oracle_built, sceneprep_lib):
    scene = scenes.config_scene("C3", n=150_000)
    views, _ = V.perspective_views([(1.1, 0.05)], 1.3, 16 / 9, 0.01)
    _run_both(scene, views, strides=[48, 64])
