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


# ---- f2: TransformSystem::animateAsync on the device ---------------------------------------------------------------------------
ROT_TOL = 4e-6  # |component| difference allowed on the slerp branch for unit quaternions (device acosf / sinf vs host libm)


def _compare_trs(got: np.ndarray, want: np.ndarray, what: str):
    """got / want: [occupancy, >= 74] transform bytes. Position, scale (incl. their lane W) and the active flags bit for bit,
    rotation within ROT_TOL. Returns the largest rotation difference seen."""
    assert np.array_equal(got[:, 16:48], want[:, 16:48]), f"{what}: position / scale bytes differ"
    assert np.array_equal(got[:, 72:74], want[:, 72:74]), f"{what}: selfActive / ancestorsActive differ"
    rg, rw = got[:, 48:64].copy().view(np.float32), want[:, 48:64].copy().view(np.float32)
    diff = np.abs(rg.astype(np.float64) - rw.astype(np.float64))
    assert diff.max() <= ROT_TOL, f"{what}: rotation differs by {diff.max():.3e} (tolerance {ROT_TOL})"
    return float(diff.max()), int((rg.view(np.uint32) != rw.view(np.uint32)).sum())


def test_animate_matches_reference_golden(sceneprep_lib):
    """tests/golden/frows/f2_animate.npz: the reference's own animateAsync, step by step. lerp (position, scale), the lerp
    branch of slerp and the active flags are exact; the slerp branch is within ROT_TOL."""
    from pathlib import Path
    from garden_b200.binding import ScenePrep
    g = np.load(Path(__file__).resolve().parent / "golden" / "frows" / "f2_animate.npz")
    t = g["transforms"].copy()
    occ, stride = t.shape
    sp = ScenePrep(0)
    sp.set_transforms(t, stride, occ)
    worst, inexact = 0.0, 0
    for step in range(int(g["steps"][0])):
        sp.animate(g[f"ids{step}"], g[f"flags{step}"], g[f"a{step}"], g[f"b{step}"], g[f"t{step}"])
        sp.writeback_trs(t, stride)
        sp.writeback_active(t, stride)
        want = t.copy()
        want[:, 16:64] = g[f"after{step}"][:, :48]
        want[:, 72:74] = g[f"after{step}"][:, 48:50]
        d, n = _compare_trs(t, want, f"golden step {step}")
        worst, inexact = max(worst, d), inexact + n
        t[:, 48:64] = want[:, 48:64]  # continue from the reference's bytes
        sp.set_transforms(t, stride, occ)
    print(f"f2 golden: largest rotation difference {worst:.3e}, {inexact} rotation components not bit-identical")
    sp.close()


def test_animate_matches_oracle_and_feeds_the_frame(sceneprep_lib, oracle_built):
    """Random keyframe pairs (both slerp branches, both hemispheres, t at the rounding boundary of isActive) against the pinned
    oracle; then a frame over the animated pool: transforms that only had position / scale / flags animated are bit-exact, so
    with rotation animation switched off the whole frame must equal the oracle's frame over the oracle-animated bytes."""
    from garden_b200.binding import ScenePrep
    scene = mixed_scene(seed=33, n=3000, max_depth=8, with_ui=False, with_ready=False)
    views, _ = V.perspective_views([(0.7, -0.05)], 1.3, 16 / 9, 0.01)
    t, pools = aos_inputs(scene)
    stride, occ = t.dtype.itemsize, t.size
    raw = t.view(np.uint8).reshape(occ, stride)
    ents = t["entity"].copy()
    live = np.nonzero(ents)[0]
    rng = np.random.default_rng(9)
    o = reflib.Oracle()
    sp = ScenePrep(0)
    _stage(sp, scene, views, t, pools)
    worst = 0.0
    for step in range(5):
        n = 400
        pick = rng.choice(live, size=n, replace=False)
        flags = rng.integers(0, 64, n).astype(np.uint8)
        if step >= 3:
            flags &= np.uint8(0xFB)  # no rotation animation: everything stays bit-exact
        fa = rng.uniform(-5, 5, (n, 10)).astype(np.float32); fb = rng.uniform(-5, 5, (n, 10)).astype(np.float32)
        fa[:, 3:6] = np.abs(fa[:, 3:6]) * 0.2 + 0.5; fb[:, 3:6] = np.abs(fb[:, 3:6]) * 0.2 + 0.5
        for f in (fa, fb):
            f[:, 6:] /= np.linalg.norm(f[:, 6:], axis=1, keepdims=True).astype(np.float32)
        near = rng.random(n) < 0.25
        fb[near, 6:] = fa[near, 6:] * np.where(rng.random(near.sum()) < 0.5, 1.0, -1.0)[:, None].astype(np.float32)
        fb[near, 6] = np.nextafter(fb[near, 6], np.float32(2.0))
        tt = rng.random(n).astype(np.float32)
        tt[:8] = [0.0, 1.0, 0.5, 0.49999997, 0.50000006, 0.25, 0.75, 1.0]
        want = raw.copy()
        assert o.animate(want, stride, occ, ents[pick], flags, fa, fb, tt) == 0
        sp.animate(ents[pick], flags, fa, fb, tt)
        got = raw.copy()
        sp.writeback_trs(got, stride)
        sp.writeback_active(got, stride)
        d, _ = _compare_trs(got, want, f"step {step}")
        worst = max(worst, d)
        if step >= 3:
            assert np.array_equal(got[:, 48:64], want[:, 48:64])
        raw[:] = want
        # keep both sides on the same bytes (the rotation tolerance must not accumulate): restage from the oracle's result
        sp.set_transforms(t, stride, occ)
        if step >= 3:
            sp.animate(ents[pick], flags, fa, fb, tt)  # idempotent for these flags: same inputs, same bytes
    sp.set_views(views, scene.camera_pos)
    sp.run()
    rts = [p.render_type for p in scene.pools]
    orun = OracleRun((t, stride, occ), [(m, m.dtype.itemsize, m.size) for m in pools], rts, views, scene.camera_pos)
    compare_gpu_to_oracle(sp, orun, rts, views, "frame after gsp_animate")
    print(f"f2 oracle: largest rotation difference {worst:.3e}")
    sp.close()


def test_animate_rejects_unknown_entities(sceneprep_lib):
    from garden_b200.binding import GSP_ERR_INVALID, ScenePrep, ScenePrepError
    scene = scenes.config_scene("C2", n=1000)
    t, pools = aos_inputs(scene)
    sp = ScenePrep(0)
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    ids = np.array([1, 5000000], np.uint32)
    fr = np.zeros((2, 10), np.float32); fr[:, 9] = 1.0
    with pytest.raises(ScenePrepError) as e:
        sp.animate(ids, np.array([1, 1], np.uint8), fr, fr, np.zeros(2, np.float32))
    assert e.value.code == GSP_ERR_INVALID
    sp.close()
