#!/usr/bin/env python
"""Generates the golden vectors in tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref/libgarden_ref_parity.so:
the unmodified cfnptr/garden translation units built by oracle/Makefile with -O3 -DNDEBUG -march=haswell -ffp-contract=off).

Run here (needs /root/reference for the _ref build):   python tests/golden/make_golden.py
The reference's own tests hold no vectors for this path (calcModel / Frustum / isBehindFrustum / key / sort are
"parity unpinned" upstream), so these files are the pin: inputs are the raw bytes of the reference's live ECS pools
(LinearPool<TransformComponent>, LinearPool<mesh component>) after building the scene through createEntity / add<> /
setParent / setActive / destroy, outputs are the reference's draw lists per view in canonical order
(key, then bufferIndex, then componentOffset; OIT buffers by componentOffset), its counts and MeshRenderComponent::isVisible.

Two fields the path never reads are zeroed in the stored input bytes so the files are reproducible: TransformComponent::uid
(random, transform.hpp:37) and TransformComponent::childs (a heap pointer, transform.hpp:56).
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import reflib  # noqa: E402
from common import canonical, canonical_oit, ref_frame  # noqa: E402
from edge_scenes import few_planes_views, mixed_scene, mixed_views  # noqa: E402
from garden_b200 import scenes, views as V  # noqa: E402
from garden_b200.layout import RT_OIT, RT_TRANSLUCENT, RT_UI  # noqa: E402


def cases():
    s = scenes.config_scene("C1")
    yield "c1_flat_10k", s, V.perspective_views([(0.4, -0.05)], 1.2, 16 / 9, 0.01)[0], None
    s = scenes.config_scene("C2", n=3000)
    s.camera_pos = np.array([5.0, 2.0, -3.0], np.float32)
    yield "c2_depth4_5views", s, V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))[0], None
    s = scenes.config_scene("C3", n=3000)
    for k, p in enumerate(s.pools):
        p.stride = 48 + 16 * (k % 3)
    yield "c3_depth8_opaque_translucent", s, V.perspective_views([(1.1, 0.05)], 1.3, 16 / 9, 0.01)[0], None
    yield "mixed_all_branches", mixed_scene(seed=7, n=2000), mixed_views(), None
    s = mixed_scene(seed=11, n=1500, with_ui=False)
    leaves = np.ones(s.entity_count, bool)
    leaves[s.parent[s.parent >= 0]] = False
    yield "freed_slots_few_planes", s, few_planes_views(), np.nonzero(leaves)[0][::7].astype(np.uint32)
    # isDrawReady(shadowPass) differs between the main pass and the shadow passes (instance.cpp:61-113, label.cpp:262-265):
    # pool 0 has no shadow pipeline, pool 2's base pipeline is not ready yet but its shadow pipeline is, pool 1 (translucent,
    # shared list) is shadow-only, pool 4 main-only
    s = mixed_scene(seed=13, n=1800)
    s.pools[0].draw_ready_shadow = False
    s.pools[2].draw_ready = False; s.pools[2].draw_ready_shadow = True
    s.pools[1].draw_ready = False; s.pools[1].draw_ready_shadow = True
    s.pools[4].draw_ready_shadow = False
    yield "shadow_readiness_differs", s, mixed_views(), None


def main():
    if not reflib.ref_available("parity"):
        raise SystemExit("oracle/_ref/libgarden_ref_parity.so is missing: run `make -C oracle ref` where /root/reference exists")
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name, scene, views, victims in cases():
        if only and name not in only:
            continue
        rts = [p.render_type for p in scene.pools]
        unsorted_types = [rt for rt in rts if rt not in (RT_TRANSLUCENT, RT_UI)]
        with reflib.RefEngine("parity", threads=-1) as ref:
            ref.load_scene(scene)
            if victims is not None:
                ref.destroy_entities(victims)
            out = {}
            taddr, tstride, tocc = ref.transform_pool()
            tbytes = ref.transform_bytes().reshape(tocc, tstride).copy()
            tbytes[:, 8:16] = 0    # uid
            tbytes[:, 64:72] = 0   # childs
            out["transforms"] = tbytes
            meta = []
            for k in range(len(scene.pools)):
                addr, stride, occ, count = ref.mesh_pool(k)
                mb = ref.pool_bytes(k).reshape(occ, stride).copy() if occ else np.zeros((0, stride), np.uint8)
                mb[:, 15] = 0  # isVisible is an output
                out[f"pool{k}"] = mb
                rc = ref.pool_ready_counts(k)
                out[f"ready{k}"] = np.zeros(0, np.uint8) if rc is None else np.pad(rc, (0, max(0, occ - rc.size)), constant_values=1)[:occ]
                pd = scene.pools[k]
                shadow_ready = pd.draw_ready if pd.draw_ready_shadow is None else pd.draw_ready_shadow
                meta.append([rts[k], stride, occ, count, 1 if pd.draw_ready else 0, 0 if rc is None else 1,
                             1 if shadow_ready else 0])
            out["pool_meta"] = np.array(meta, dtype=np.uint32)
            out["views"] = views
            out["camera_pos"] = np.asarray(scene.camera_pos, np.float32)
            frame = ref_frame(ref, views)
            total = 0
            for v, res in enumerate(frame):
                for b, (rec, draw, inst) in enumerate(res["unsorted"]):
                    rec = canonical_oit(rec) if unsorted_types[b] == RT_OIT else canonical(rec, False)
                    out[f"v{v}_unsorted{b}"] = rec
                    out[f"v{v}_unsorted{b}_counts"] = np.array([draw, inst], np.uint32)
                    total += draw
                out[f"v{v}_sorted_counts"] = np.array(res["sorted_counts"], np.uint32).reshape(-1, 2)
                out[f"v{v}_trans"] = canonical(res["trans"][0], True)
                out[f"v{v}_ui"] = canonical(res["ui"][0], True)
                total += res["trans"][1] + res["ui"][1]
                if "visible" in res:
                    for k, vis in enumerate(res["visible"]):
                        out[f"v{v}_visible{k}"] = vis
        path = HERE / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(f"{name}: {scene.entity_count} entities, {views.size} views, {total} records -> {path.stat().st_size / 1024:.0f} KiB")


def f_rows():
    """Golden vectors of the SURVEY.md 8f rows, from the reference itself -> tests/golden/frows/f_rows.npz:
      f1  per view: viewProj, the reference's draw list (first 400 records) and (float4x4)(viewProj * model) per record;
      f3  transform pool bytes, a sequence of TransformComponent::setActive calls, and bytes 72/73 after every call."""
    out = {}
    scene = scenes.config_scene("C2", n=4000)
    scene.camera_pos = np.array([5.0, 2.0, -3.0], np.float32)
    views, vps = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        for v in range(views.size):
            ref.prepare(views[v])
            rec = canonical(ref.get_unsorted(0)[0], False)[:400]
            vp = np.asarray(vps[v], dtype=np.float32).reshape(16)
            out[f"f1_vp{v}"] = vp
            out[f"f1_records{v}"] = rec
            out[f"f1_mvp{v}"] = ref.instance_mvp(vp, rec)
    out["f1_views"] = np.array([views.size], np.uint32)
    scene = mixed_scene(seed=5, n=1200, max_depth=12, with_ui=False, with_ready=False)
    rng = np.random.default_rng(23)
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        _, tstride, tocc = ref.transform_pool()
        tbytes = ref.transform_bytes().reshape(tocc, tstride).copy()
        tbytes[:, 8:16] = 0    # uid
        tbytes[:, 64:72] = 0   # childs
        out["f3_transforms"] = tbytes
        ents = tbytes[:, 0:4].copy().view(np.uint32).reshape(-1)
        live = np.nonzero(ents)[0]
        steps = 10
        for step in range(steps):
            pick = rng.choice(live, size=int(rng.integers(1, 40)), replace=True)
            active = bool(step % 3 == 2) if step < 6 else bool(rng.integers(0, 2))
            ref.set_active(ents[pick] - 1, active)
            out[f"f3_ids{step}"] = ents[pick].astype(np.uint32)
            out[f"f3_active{step}"] = np.array([1 if active else 0], np.uint8)
            out[f"f3_after{step}"] = ref.transform_bytes().reshape(tocc, tstride)[:, 72:74].copy()
        out["f3_steps"] = np.array([steps], np.uint32)
    (HERE / "frows").mkdir(exist_ok=True)
    path = HERE / "frows" / "f_rows.npz"
    np.savez_compressed(path, **out)
    print(f"f_rows -> {path.stat().st_size / 1024:.0f} KiB")


def f2_animate():
    """Golden vectors of TransformSystem::animateAsync from the reference itself -> frows/f2_animate.npz:
    transform pool bytes, per step the animated entity ids / flags / frames / t, and the pool's TRS + active bytes after it."""
    out = {}
    scene = mixed_scene(seed=31, n=900, max_depth=8, with_ui=False, with_ready=False)
    rng = np.random.default_rng(77)
    with reflib.RefEngine("parity", threads=0) as ref:
        ref.load_scene(scene)
        _, tstride, tocc = ref.transform_pool()
        tbytes = ref.transform_bytes().reshape(tocc, tstride).copy()
        tbytes[:, 8:16] = 0    # uid
        tbytes[:, 64:72] = 0   # childs
        out["transforms"] = tbytes
        ents = tbytes[:, 0:4].copy().view(np.uint32).reshape(-1)
        live = np.nonzero(ents)[0]
        steps = 4
        for step in range(steps):
            n = 120
            pick = rng.choice(live, size=n, replace=False)
            flags = rng.integers(0, 64, n).astype(np.uint8)
            fa = rng.uniform(-5, 5, (n, 10)).astype(np.float32); fb = rng.uniform(-5, 5, (n, 10)).astype(np.float32)
            for f in (fa, fb):
                f[:, 6:] /= np.linalg.norm(f[:, 6:], axis=1, keepdims=True).astype(np.float32)
            near = rng.random(n) < 0.25
            fb[near, 6:] = fa[near, 6:] * np.where(rng.random(near.sum()) < 0.5, 1.0, -1.0)[:, None].astype(np.float32)
            t = rng.random(n).astype(np.float32)
            t[:6] = [0.0, 1.0, 0.5, 0.49999997, 0.50000006, 0.75]
            ref.animate(ents[pick] - 1, flags, fa, fb, t)
            after = ref.transform_bytes().reshape(tocc, tstride)
            out[f"ids{step}"], out[f"flags{step}"], out[f"a{step}"], out[f"b{step}"], out[f"t{step}"] = ents[pick].astype(np.uint32), flags, fa, fb, t
            out[f"after{step}"] = np.concatenate([after[:, 16:64], after[:, 72:74]], axis=1).copy()
        out["steps"] = np.array([steps], np.uint32)
    path = HERE / "frows" / "f2_animate.npz"
    np.savez_compressed(path, **out)
    print(f"f2_animate -> {path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()  # (names on the command line: only those cases)
    if len(sys.argv) == 1:
        f_rows()
        f2_animate()
