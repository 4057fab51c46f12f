"""GPU parity of the SURVEY.md §8f rows, through the C ABI, against the oracle:
  f1 gsp_emit_instances[_device]   mvp = (float4x4)(viewProj * f32x4x4(bakedModel, (0,0,0,1))) per record, in draw order
  f3 gsp_set_active                TransformComponent::setActive on the staged hierarchy, then a frame that filters on it"""
import numpy as np
import pytest

import reflib
from garden_b200 import scenes, views as V
from common import OracleRun, aos_inputs, compare_gpu_to_oracle
from edge_scenes import mixed_scene

pytestmark = pytest.mark.gpu


def _stage(sp, scene, views, t, pools):
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    sp.set_pool_count(len(pools))
    for k, m in enumerate(pools):
        sp.set_mesh_pool(k, scene.pools[k].render_type, m, m.dtype.itemsize, m.size)
    sp.set_views(views, scene.camera_pos)


def test_instances_match_oracle(sceneprep_lib, oracle_built):
    import torch
    from garden_b200.binding import ScenePrep
    scene = scenes.config_scene("C3", n=60000)  # opaque buffer + shared translucent list
    scene.camera_pos = np.array([1.0, 0.5, -2.0], np.float32)
    views, vps = V.perspective_views([(1.1, 0.05), (0.2, -0.1)], 1.3, 16 / 9, 0.01)
    t, pools = aos_inputs(scene)
    sp = ScenePrep(0)
    _stage(sp, scene, views, t, pools)
    sp.run()
    o = reflib.Oracle()
    checked = 0
    for v in range(views.size):
        vp = np.asarray(vps[v], dtype=np.float32).reshape(16)
        lists = [(0, b, sp.get_unsorted(v, b)[0]) for b in range(sp.unsorted_buffer_count(v))] + [(1, 0, sp.get_sorted(v, 0)[0])]
        for kind, b, rec in lists:
            want = o.instance_mvp(vp, rec)
            # host instance buffer with a 96-byte BaseInstanceData-like stride, mvp at offset 16
            got = sp.emit_instances(v, kind, b, vp, rec.size, stride=96, mvp_offset=16)
            assert got.shape == (rec.size, 96)
            g = np.ascontiguousarray(got[:, 16:80]).view(np.float32).reshape(-1, 16)
            assert np.array_equal(g.view(np.uint32), want.view(np.uint32)), f"view {v} kind {kind} buffer {b}: mvp bits"
            assert not got[:, :16].any() and not got[:, 80:].any(), "bytes outside the mvp field must not be touched"
            checked += rec.size
    assert checked > 1000
    # device variant: enqueued right after gsp_run_async (draw count read on the device), packed 64-byte instances;
    # the capacity clamps what is written
    sp.run_async()
    vp = np.asarray(vps[0], dtype=np.float32).reshape(16)
    cap = 100000
    dst = torch.full((cap * 16,), float("nan"), dtype=torch.float32, device="cuda")
    sp.emit_instances_device(0, 0, 0, vp, dst.data_ptr(), cap)
    small = torch.full((8 * 16,), float("nan"), dtype=torch.float32, device="cuda")
    sp.emit_instances_device(0, 0, 0, vp, small.data_ptr(), 7)
    sp.sync()
    torch.cuda.synchronize()
    rec = sp.get_unsorted(0, 0)[0]
    want = o.instance_mvp(vp, rec)
    got = dst.cpu().numpy().reshape(cap, 16)
    assert np.array_equal(got[:rec.size].view(np.uint32), want.view(np.uint32))
    assert np.isnan(got[rec.size:]).all(), "nothing past the draw count may be written"
    sm = small.cpu().numpy().reshape(8, 16)
    assert np.array_equal(sm[:7].view(np.uint32), want[:7].view(np.uint32)) and np.isnan(sm[7]).all()
    # argument checks
    with pytest.raises(Exception):
        sp.emit_instances(0, 0, 0, vp, 4, stride=72)       # stride not a multiple of 16
    with pytest.raises(Exception):
        sp.emit_instances(0, 0, 0, vp, 4, stride=64, mvp_offset=16)  # mvp does not fit the stride
    sp.close()


def test_set_active_then_frame_matches_oracle(sceneprep_lib, oracle_built):
    """Random setActive sequences on the device vs the oracle on the host bytes; after each batch a frame is prepared on both
    sides (the cull filter reads isActive(), mesh.cpp:150-155 / transform.hpp:110) and compared bit for bit."""
    from garden_b200.binding import ScenePrep, ScenePrepError
    scene = mixed_scene(seed=9, n=4000, max_depth=10, with_ui=False, with_ready=False)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    t, pools = aos_inputs(scene)
    rts = [p.render_type for p in scene.pools]
    sp = ScenePrep(0)
    _stage(sp, scene, views, t, pools)
    o = reflib.Oracle()
    host = t.copy()                       # the oracle's pool
    raw = host.view(np.uint8).reshape(host.size, host.dtype.itemsize)
    mirror = t.copy()                     # receives gsp_writeback_active
    mraw = mirror.view(np.uint8).reshape(mirror.size, mirror.dtype.itemsize)
    ents = host["entity"].copy()
    live = np.nonzero(ents)[0]
    rng = np.random.default_rng(3)
    for step in range(6):
        pick = rng.choice(live, size=int(rng.integers(1, 200)), replace=True)
        active = step % 2 == 1
        assert o.set_active(raw, host.dtype.itemsize, host.size, ents[pick], active) == 0
        sp.set_active(ents[pick], active)
        sp.writeback_active(mirror, mirror.dtype.itemsize)
        assert np.array_equal(mraw[:, 72:74], raw[:, 72:74]), f"step {step}: selfActive / ancestorsActive"
        sp.run()
        orun = OracleRun((host, host.dtype.itemsize, host.size), [(m, m.dtype.itemsize, m.size) for m in pools], rts, views,
                         scene.camera_pos)
        compare_gpu_to_oracle(sp, orun, rts, views, f"after setActive batch {step}")
    assert 0 < int(raw[live, 73].sum()) < live.size
    # an id without a TransformComponent is reported (Manager::get would throw), the valid ids are still applied
    with pytest.raises(ScenePrepError):
        sp.set_active(np.array([int(ents.max()) + 5], np.uint32), False)
    sp.close()
