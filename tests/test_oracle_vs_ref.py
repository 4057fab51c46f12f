"""Pins the oracle: oracle/sceneprep_oracle.c (plain-C restatement) against the reference's OWN translation units
(oracle/_ref/libgarden_ref_parity.so = mesh.cpp + transform.cpp + thread-pool.cpp + ecsm + math, built by oracle/Makefile with
-ffp-contract=off) — bit for bit, through the reference's real ECS memory.

Needs /root/reference at build time (this container); skipped where oracle/_ref is absent. The committed golden vectors in
tests/golden/ (generated from the same reference build by tests/golden/make_golden.py) cover the boxes without it.
"""
import numpy as np
import pytest

import reflib
from garden_b200 import scenes, views as V
from garden_b200.layout import RT_TRANSLUCENT, RT_UI

from common import OracleRun, assert_frames_equal, ref_frame
from edge_scenes import few_planes_views, mixed_scene, mixed_views

pytestmark = pytest.mark.skipif(not reflib.ref_available("parity"), reason="oracle/_ref not built (no /root/reference)")


def _oracle_on_ref_memory(ref, scene, views):
    taddr, tstride, tocc = ref.transform_pool()
    pools = [ref.mesh_pool(k) for k in range(len(scene.pools))]
    rts = [p.render_type for p in scene.pools]
    ready = [ref.pool_ready_counts(k) for k in range(len(scene.pools))]
    draw_ready = [p.draw_ready for p in scene.pools]
    return OracleRun((taddr, tstride, tocc), [(a, s, o) for a, s, o, c in pools], rts, views, scene.camera_pos,
                     ready=ready, draw_ready=draw_ready, counts=[c for a, s, o, c in pools])


def _check(scene, views, threads=-1, mutate=None):
    rts = [p.render_type for p in scene.pools]
    with reflib.RefEngine("parity", threads=threads) as ref:
        ref.load_scene(scene)
        if mutate is not None:
            mutate(ref)
        got = ref_frame(ref, views)
        want = _oracle_on_ref_memory(ref, scene, views)
        assert_frames_equal(got, want.views, rts, scene.name)
    return got


def test_primitives_match_reference(oracle_built):
    """calcModel chains (transform.hpp:197-214) and Frustum(viewProj) (frustum.hpp:51-61), value by value."""
    scene = mixed_scene(seed=3, n=1500, max_depth=20, with_ui=False, with_ready=False)
    o = reflib.Oracle()
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        taddr, tstride, tocc = ref.transform_pool()
        o.set_transforms(taddr, tstride, tocc)
        tbytes = ref.transform_bytes().reshape(tocc, tstride)
        ents = tbytes[:, 0:4].copy().view(np.uint32).reshape(-1)
        cam = np.array([0.25, -3.0, 8.5], np.float32)
        checked = 0
        for slot in range(tocc):
            if ents[slot] == 0:
                continue
            a = ref.calc_model(int(ents[slot]) - 1, cam)
            b = o.calc_model(slot, cam)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"calcModel bits differ at slot {slot}"
            checked += 1
        assert checked > 1000
        rng = np.random.default_rng(5)
        for _ in range(50):
            vp = rng.standard_normal((4, 4)).astype(np.float32)
            assert np.array_equal(ref.frustum_planes(vp).view(np.uint32), o.frustum_planes(vp).view(np.uint32))


def test_c1_flat(oracle_built):
    scene = scenes.config_scene("C1")
    views, _ = V.perspective_views([(0.4, -0.05)], 1.2, 16 / 9, 0.01)
    got = _check(scene, views)
    assert got[0]["unsorted"][0][1] > 500


def test_c2_depth4_five_views(oracle_built):
    scene = scenes.config_scene("C2", n=40_000)
    scene.camera_pos = np.array([5.0, 2.0, -3.0], np.float32)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    _check(scene, views)


def test_c3_depth8_opaque_translucent(oracle_built):
    scene = scenes.config_scene("C3", n=30_000)
    for k, p in enumerate(scene.pools):
        p.stride = 48 + 16 * (k % 3)
    views, _ = V.perspective_views([(1.1, 0.05)], 1.3, 16 / 9, 0.01)
    _check(scene, views)


@pytest.mark.parametrize("threads", [-1, 0])
def test_mixed_scene_all_branches(oracle_built, threads):
    scene = mixed_scene(seed=7, single_translucent=threads == 0)
    got = _check(scene, mixed_views(), threads=threads)
    main = got[-1]
    assert main["ui"][1] > 0 and main["trans"][1] > 0 and all(u[1] > 0 for u in main["unsorted"])


def test_freed_slots_and_few_planes(oracle_built):
    """destroy() leaves free slots (entity == 0) inside the pools (linear-pool.hpp); frusta with 1 and 4 planes."""
    scene = mixed_scene(seed=11, n=2500, with_ui=False)
    leaves = np.ones(scene.entity_count, bool)
    leaves[scene.parent[scene.parent >= 0]] = False
    victims = np.nonzero(leaves)[0][::7].astype(np.uint32)

    def mutate(ref):
        ref.destroy_entities(victims)
    _check(scene, few_planes_views(), mutate=mutate)


def test_draw_ready_false_and_empty_pool(oracle_built):
    scene = mixed_scene(seed=13, n=1200, with_ui=False, with_ready=False)
    scene.pools[0].draw_ready = False
    empty = scene.pools[2]
    scene.pools[2] = scenes.PoolDesc(empty.render_type, empty.entity_index[:0], empty.aabb[:0], empty.enabled[:0], None,
                                     empty.stride)
    _check(scene, mixed_views(with_ui=False))


def test_animated_update(oracle_built):
    """C3's per-frame TRS rewrite (plain stores, transform.hpp:74-104) followed by another frame."""
    scene = scenes.config_scene("C3", n=10_000)
    for k, p in enumerate(scene.pools):
        p.stride = 48 + 16 * (k % 3)
    views, _ = V.perspective_views([(0.2, 0.0)], 1.3, 16 / 9, 0.01)
    rts = [p.render_type for p in scene.pools]
    with reflib.RefEngine("parity") as ref:
        ref.load_scene(scene)
        for frame in range(3):
            sel, p, r, s = scenes.animate_trs(scene, frame, 99)
            ref.update_trs(sel, p, r, s)
            got = ref_frame(ref, views)
            want = _oracle_on_ref_memory(ref, scene, views)
            assert_frames_equal(got, want.views, rts, f"frame {frame}")


# ---- SURVEY.md §8f rows -------------------------------------------------------------------------------------------------
def test_instance_mvp_matches_reference(oracle_built):
    """f1: (float4x4)(viewProj * f32x4x4(bakedModel, (0,0,0,1))) by the reference's math library vs the restatement, bit for
    bit, on the draw lists of every view of a C2-shaped scene (perspective and orthographic viewProj)."""
    scene = scenes.config_scene("C2", n=4000)
    scene.camera_pos = np.array([5.0, 2.0, -3.0], np.float32)
    views, vps = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    o = reflib.Oracle()
    total = 0
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        for v in range(views.size):
            ref.prepare(views[v])
            rec = ref.get_unsorted(0)[0]
            vp = np.asarray(vps[v], dtype=np.float32).reshape(16)
            a, b = ref.instance_mvp(vp, rec), o.instance_mvp(vp, rec)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"view {v}: mvp bits differ"
            total += rec.size
    assert total > 500


def test_set_active_matches_reference(oracle_built):
    """f3: TransformComponent::setActive sequences through the real ECS vs the restatement on a copy of the pool bytes:
    selfActive / ancestorsActive of every transform after every call."""
    scene = mixed_scene(seed=5, n=1200, max_depth=12, with_ui=False, with_ready=False)
    rng = np.random.default_rng(17)
    o = reflib.Oracle()
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        _, tstride, tocc = ref.transform_pool()
        mine = ref.transform_bytes().reshape(tocc, tstride).copy()
        ents = mine[:, 0:4].copy().view(np.uint32).reshape(-1)
        live = np.nonzero(ents)[0]
        for step in range(12):
            pick = rng.choice(live, size=int(rng.integers(1, 40)), replace=True)  # duplicates included
            active = bool(step % 3 == 2) if step < 8 else bool(rng.integers(0, 2))
            ref.set_active(ents[pick] - 1, active)          # harness takes 0-based creation indices == id - 1
            assert o.set_active(mine, tstride, tocc, ents[pick], active) == 0
            theirs = ref.transform_bytes().reshape(tocc, tstride)
            assert np.array_equal(theirs[:, 72:74], mine[:, 72:74]), f"step {step}: active flags differ"
        assert 0 < int(mine[live, 73].sum()) < live.size  # both states occur


@pytest.mark.parametrize("seed", list(range(20, 60)))
def test_fuzz_mixed_scenes(oracle_built, seed):
    """Differential fuzzing of the pin: random forests (depth up to 3..40), random filter states, render-type buckets, ready
    counts, camera yaw and thread mode per seed — the oracle must reproduce the reference's lists bit for bit every time."""
    rng = np.random.default_rng(seed)
    threads = 0 if seed % 3 == 0 else -1
    scene = mixed_scene(seed=seed, n=int(rng.integers(300, 2500)), max_depth=int(rng.integers(3, 40)),
                        with_ui=bool(seed % 2), with_ready=bool((seed // 2) % 2), box_half=float(rng.uniform(20.0, 120.0)),
                        single_translucent=threads == 0)
    scene.camera_pos = rng.uniform(-30.0, 30.0, 3).astype(np.float32)
    _check(scene, mixed_views(yaw=float(rng.uniform(-3.0, 3.0)), with_ui=bool(seed % 2)), threads=threads)


def test_animate_matches_reference(oracle_built):
    """f2 (oracle side): TransformSystem::animateAsync through the real system vs the restatement on a copy of the pool bytes —
    lerp of position / scale, slerp of rotation (both branches, both hemispheres), isActive switching at t = 0.5, bit for bit
    (the host libm serves both sides, as it serves the reference)."""
    scene = mixed_scene(seed=31, n=900, max_depth=8, with_ui=False, with_ready=False)
    rng = np.random.default_rng(5)
    o = reflib.Oracle()
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        _, tstride, tocc = ref.transform_pool()
        mine = ref.transform_bytes().reshape(tocc, tstride).copy()
        ents = mine[:, 0:4].copy().view(np.uint32).reshape(-1)
        live = np.nonzero(ents)[0]
        for step in range(6):
            n = 150
            pick = rng.choice(live, size=n, replace=False)
            flags = rng.integers(0, 64, n).astype(np.uint8)
            fa = rng.uniform(-5, 5, (n, 10)).astype(np.float32); fb = rng.uniform(-5, 5, (n, 10)).astype(np.float32)
            for f in (fa, fb):  # unit quaternions; some pairs nearly parallel (lerp branch), some in opposite hemispheres
                f[:, 6:] /= np.linalg.norm(f[:, 6:], axis=1, keepdims=True).astype(np.float32)
            near = rng.random(n) < 0.25
            fb[near, 6:] = fa[near, 6:] * np.where(rng.random(near.sum()) < 0.5, 1.0, -1.0)[:, None].astype(np.float32)
            fb[near, 6] = np.nextafter(fb[near, 6], np.float32(2.0))
            t = rng.random(n).astype(np.float32)
            t[:8] = [0.0, 1.0, 0.5, 0.49999997, 0.50000006, 0.25, 0.75, 1.0]
            ref.animate(ents[pick] - 1, flags, fa, fb, t)
            assert o.animate(mine, tstride, tocc, ents[pick], flags, fa, fb, t) == 0
            theirs = ref.transform_bytes().reshape(tocc, tstride)
            for lo, hi, what in ((16, 28, "position"), (32, 44, "scale"), (48, 64, "rotation"), (72, 74, "active flags")):
                assert np.array_equal(theirs[:, lo:hi], mine[:, lo:hi]), f"step {step}: {what} differs"
            assert np.array_equal(theirs[:, 28:32], mine[:, 28:32]) and np.array_equal(theirs[:, 44:48], mine[:, 44:48]), "lane W"
