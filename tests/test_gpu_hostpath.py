"""Host boundary of the C ABI: where the staging kernels read the caller's AoS bytes from (pageable copy, pinned memory read
in place over PCIe, device pointers, unaligned bases, wide derived components) and how isVisible travels back
(full vs changed-slots-only). Every variant must give the oracle's lists bit for bit."""
import numpy as np
import pytest

from garden_b200 import scenes, views as V
from garden_b200.layout import RT_OPAQUE

from common import OracleRun, aos_inputs, compare_gpu_to_oracle

pytestmark = pytest.mark.gpu


def _scene():
    scene = scenes.config_scene("C3", n=70_001)
    scene.camera_pos = np.array([2.0, 1.0, -4.0], np.float32)
    views, _ = V.perspective_views([(0.7, -0.05)], 1.3, 16 / 9, 0.01)
    return scene, views


def _stage_and_run(sp, t, pools, rts, views, cam):
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    sp.set_pool_count(len(pools))
    for k, m in enumerate(pools):
        sp.set_mesh_pool(k, rts[k], m, m.dtype.itemsize, m.size)
    sp.set_views(views, cam)
    sp.run()


def _oracle(t, pools, rts, views, cam):
    return OracleRun((t, t.dtype.itemsize, t.size), [(m, m.dtype.itemsize, m.size) for m in pools], rts, views, cam)


@pytest.mark.parametrize("source", ["pageable", "pinned", "pinned_unaligned", "device"])
def test_upload_sources(oracle_built, sceneprep_lib, source):
    import torch
    from garden_b200.binding import ScenePrep, pin_host, unpin_host
    scene, views = _scene()
    t, pools = aos_inputs(scene, strides=[48, 112])
    rts = [p.render_type for p in scene.pools]
    orun = _oracle(t, pools, rts, views, scene.camera_pos)
    keep = []
    if source == "pinned":
        for a in [t] + pools:
            pin_host(a)
        tt, pp = t, pools
    elif source == "pinned_unaligned":
        # a pool whose base is only 4-byte aligned (the 32-bit load path of the tile loader)
        def shifted(a):
            raw = torch.empty(a.nbytes + 64, dtype=torch.uint8, pin_memory=True)
            keep.append(raw)
            view = raw.numpy()[4:4 + a.nbytes].view(a.dtype)
            view[...] = a
            return view
        tt, pp = shifted(t), [shifted(m) for m in pools]
    elif source == "device":
        def dev(a):
            d = torch.from_numpy(a.view(np.uint8).copy()).cuda()
            keep.append(d)
            return d.data_ptr()
        tt, pp = dev(t), [dev(m) for m in pools]
    else:
        tt, pp = t, pools
    sp = ScenePrep(0)
    sp.set_transforms(tt, t.dtype.itemsize, t.size)
    sp.set_pool_count(len(pools))
    for k, m in enumerate(pools):
        sp.set_mesh_pool(k, rts[k], pp[k], m.dtype.itemsize, m.size)
    sp.set_views(views, scene.camera_pos)
    sp.run()
    sp.fetch_all()
    compare_gpu_to_oracle(sp, orun, rts, views, source)
    sp.close()
    if source == "pinned":
        for a in [t] + pools:
            unpin_host(a)


def test_visible_writeback_full_and_delta(oracle_built, sceneprep_lib):
    """The camera turns every frame; the pool is staged once. After each frame the host bytes must equal the oracle's
    isVisible whether they travel as a full bit mask or as the list of changed slots."""
    from garden_b200.binding import ScenePrep
    scene, _ = _scene()
    t, pools = aos_inputs(scene, strides=[48, 64])
    rts = [p.render_type for p in scene.pools]
    rng = np.random.default_rng(5)
    for m in pools:  # garbage in the host bytes: 0, 1 and values that are neither
        m["isVisible"] = rng.integers(0, 4, m.size).astype(np.uint8) * 85
    def clone(m):  # raw byte copy (ndarray.copy() of a padded struct dtype need not preserve the padding bytes)
        return np.frombuffer(bytearray(m.tobytes()), dtype=m.dtype)
    full = [clone(m) for m in pools]
    delta = [clone(m) for m in pools]
    sp_full, sp_delta = ScenePrep(0), ScenePrep(0)
    changed_counts = []
    for frame in range(4):
        views, _ = V.perspective_views([(0.7 + 0.15 * frame, -0.05)], 1.3, 16 / 9, 0.01)
        if frame == 0:
            _stage_and_run(sp_full, t, full, rts, views, scene.camera_pos)
            _stage_and_run(sp_delta, t, delta, rts, views, scene.camera_pos)
        else:
            for sp in (sp_full, sp_delta):
                sp.set_views(views, scene.camera_pos)
                sp.run()
        orun = OracleRun((t, t.dtype.itemsize, t.size), [(m, m.dtype.itemsize, m.size) for m in pools], rts, views,
                         scene.camera_pos)
        total_changed = 0
        for k in range(len(pools)):
            sp_full.writeback_visible(k, full[k], full[k].dtype.itemsize)
            total_changed += sp_delta.writeback_visible_delta(k, delta[k], delta[k].dtype.itemsize)
            want = orun.views[0]["visible"][k]
            assert np.array_equal(full[k]["isVisible"], want), f"frame {frame} pool {k}: full write-back differs"
            assert np.array_equal(delta[k]["isVisible"], want), f"frame {frame} pool {k}: delta write-back differs"
            # nothing but the isVisible byte may change
            for got in (full[k], delta[k]):
                a = got.view(np.uint8).reshape(-1, got.dtype.itemsize).copy()
                b = pools[k].view(np.uint8).reshape(-1, got.dtype.itemsize).copy()
                a[:, 15] = 0; b[:, 15] = 0
                assert np.array_equal(a, b)
        changed_counts.append(total_changed)
    # frame 0 rewrites the garbage; later frames move only what the turning camera changed
    assert changed_counts[0] > changed_counts[1] > 0, changed_counts
    # mixing the two: a full write-back keeps the device's picture of the host bytes in step
    views, _ = V.perspective_views([(2.0, 0.0)], 1.3, 16 / 9, 0.01)
    sp_delta.set_views(views, scene.camera_pos)
    sp_delta.run()
    for k in range(len(pools)):
        sp_delta.writeback_visible(k, delta[k], delta[k].dtype.itemsize)
        assert sp_delta.writeback_visible_delta(k, delta[k], delta[k].dtype.itemsize) == 0
    sp_full.close(); sp_delta.close()


def test_wide_component_stride(oracle_built, sceneprep_lib):
    """Derived mesh components are larger than MeshRenderComponent (getMeshComponentSize(), mesh.hpp:60-147): strides far
    past the tile loader's 256-slot budget."""
    from garden_b200.binding import ScenePrep
    scene = scenes.config_scene("C2", n=9_000)
    scene.camera_pos = np.array([1.0, 0.5, -2.0], np.float32)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    t, pools = aos_inputs(scene, strides=[1024])
    rts = [p.render_type for p in scene.pools]
    orun = _oracle(t, pools, rts, views, scene.camera_pos)
    sp = ScenePrep(0)
    _stage_and_run(sp, t, pools, rts, views, scene.camera_pos)
    compare_gpu_to_oracle(sp, orun, rts, views, "stride 1024")
    sp.close()


def test_sync_and_writeback_follow_the_frame_state(oracle_built, sceneprep_lib):
    """gsp_sync only validates results of a frame enqueued after the last change (no stale counters / segments), and a pool
    the new frame's main view did not process keeps its isVisible bytes (the reference would not touch them, mesh.cpp:426,482)."""
    from garden_b200.binding import GSP_ERR_STATE, ScenePrep, ScenePrepError
    scene = scenes.config_scene("C2", n=20_000)
    t, pools = aos_inputs(scene)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    sp = ScenePrep(0)
    with pytest.raises(ScenePrepError) as e:
        sp.sync()  # nothing enqueued yet
    assert e.value.code == GSP_ERR_STATE
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    sp.set_pool_count(1)
    sp.set_mesh_pool(0, RT_OPAQUE, pools[0], 48, pools[0].size)
    sp.set_views(views, scene.camera_pos)
    sp.run()
    m = pools[0]
    sp.writeback_visible(0, m, 48)
    seen = int(m["isVisible"].sum())
    assert seen > 0
    # a change invalidates the frame: sync without a new gsp_run_async must not resurrect the old results
    sp.set_views(views[:4], scene.camera_pos)  # shadow passes only
    with pytest.raises(ScenePrepError) as e:
        sp.sync()
    assert e.value.code == GSP_ERR_STATE
    with pytest.raises(ScenePrepError):
        sp.get_unsorted(0, 0)
    sp.run()
    m["isVisible"][:] = 7
    sp.writeback_visible(0, m, 48)        # no main view in this frame: bytes untouched
    assert (m["isVisible"] == 7).all()
    assert sp.writeback_visible_delta(0, m, 48) == 0 and (m["isVisible"] == 7).all()
    # the pool is shadow-ready only: the main view of a full frame skips it, the cascades keep their lists
    sp.set_views(views, scene.camera_pos)
    sp.set_pool_view_mask(0, 0b01111)
    sp.run()
    assert sp.get_unsorted(4, 0)[1] == 0 and sum(sp.get_unsorted(v, 0)[1] for v in range(4)) > 0
    sp.writeback_visible(0, m, 48)
    assert (m["isVisible"] == 7).all()
    sp.set_pool_view_mask(0, 0xFFFFFFFF)
    sp.run()
    sp.writeback_visible(0, m, 48)
    assert int(m["isVisible"].sum()) == seen
    sp.close()


def test_fetched_lists_survive_the_staging_of_the_next_frame(oracle_built, sceneprep_lib):
    """gsp_fetch_all_async takes a snapshot: the list getters keep serving frame k while frame k+1's inputs are staged (the
    pipelined end-to-end loop of bench.py), until the next run or a layout change."""
    from garden_b200.binding import GSP_ERR_STATE, ScenePrep, ScenePrepError
    scene = scenes.config_scene("C2", n=30_000)
    t, pools = aos_inputs(scene)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    sp = ScenePrep(0)

    def stage(tr, vw):
        sp.set_transforms(tr, tr.dtype.itemsize, tr.size)
        sp.set_pool_count(1)
        sp.set_mesh_pool(0, RT_OPAQUE, pools[0], 48, pools[0].size)
        sp.set_views(vw, scene.camera_pos)

    stage(t, views)
    sp.run()
    before = [sp.get_unsorted(v, 0)[0].copy() for v in range(views.size)]
    assert sum(b.size for b in before) > 0
    # frame k+1 = the same scene moved: staged while frame k's snapshot is still being read
    moved = t.copy()
    moved["position"][:, 0] += 3.0
    sp.run()
    sp.fetch_all_async()
    stage(moved, views)
    during = [sp.get_unsorted(v, 0)[0].copy() for v in range(views.size)]
    for a, b in zip(before, during):
        assert a.tobytes() == b.tobytes(), "the snapshot must still be frame k"
    assert sp.get_sorted(0, 0)[1] == 0
    with pytest.raises(ScenePrepError) as e:
        sp.sync()  # ... but nothing of the NEW inputs has run yet
    assert e.value.code == GSP_ERR_STATE
    sp.run()
    after = [sp.get_unsorted(v, 0)[0] for v in range(views.size)]
    assert any(a.tobytes() != b.tobytes() for a, b in zip(before, after)), "the next run replaces the snapshot"
    # a layout change ends the snapshot at once
    sp.fetch_all_async()
    sp.set_views(views[:4], scene.camera_pos)
    with pytest.raises(ScenePrepError) as e:
        sp.get_unsorted(0, 0)
    assert e.value.code == GSP_ERR_STATE
    sp.close()
