"""Seeded scenes that exercise every branch the reference takes on the scene-preparation path (test infrastructure).

The filters of prepareUnsortedMeshes / prepareSortedMeshes (source/system/render/mesh.cpp:140-155,217-232):
free pool slots (entity == 0), isEnabled == false, degenerate AABBs, entities without a TransformComponent,
inactive entities and inactive ancestors, modelWithAncestors == false; plus every MeshRenderType bucket
(unsorted buffers, OIT left unsorted, the shared translucent list, the UI list with its own frustum and 2-D key),
getReadyMeshesAsync overrides (ready counts 0..3), isDrawReady() == false and empty pools.
"""
from __future__ import annotations

import numpy as np

from garden_b200 import scenes, views as V
from garden_b200.layout import RT_COLOR, RT_OIT, RT_OPAQUE, RT_REFRACTED, RT_TRANSLUCENT, RT_TRANS_DEPTH, RT_UI, make_views
from garden_b200.scenes import PoolDesc, SceneDesc, hash_u01


def ref_stride(pool_index: int) -> int:
    """Component size of pool k in oracle/ref_harness.cpp (HarnessMeshComponent<K>)."""
    return 48 + 16 * (pool_index % 3)


def random_forest(seed: int, n: int, max_depth: int) -> np.ndarray:
    """Random trees: entity i picks a parent among the 40 entities before it (or none), depth capped."""
    u = hash_u01(seed, 200, np.arange(n, dtype=np.uint64))
    w = hash_u01(seed, 201, np.arange(n, dtype=np.uint64))
    parent = np.full(n, -1, np.int32)
    depth = np.zeros(n, np.int32)
    for i in range(1, n):
        if u[i] < 0.25:
            continue
        p = i - 1 - int(w[i] * min(i, 40))
        if depth[p] + 1 > max_depth:
            continue
        parent[i] = p
        depth[i] = depth[p] + 1
    return parent


def mixed_scene(seed: int = 7, n: int = 3000, max_depth: int = 12, with_ui: bool = True, with_ready: bool = True,
                box_half: float = 60.0, single_translucent: bool = False) -> SceneDesc:
    """single_translucent: the reference's NON-threaded path (asyncPreparing == false) writes every translucent system's
    records at transSortedMeshes[0..] (mesh.cpp:209-211: `meshes = combinedMeshes`, no drawOffset), so a second
    translucent system overwrites the first there. The default threaded path appends correctly (mesh.cpp:257-259) and
    is the behaviour the oracle and the CUDA path follow; single-threaded reference runs use one translucent system."""
    parent = random_forest(seed, n, max_depth)
    box = (-box_half, -8.0, -box_half, box_half, 8.0, box_half)
    pos, rot, scl = scenes.random_trs(seed, n, box, parent >= 0, local_extent=2.5, scale_range=(0.6, 1.4))
    idx = np.arange(n, dtype=np.uint64)
    tflags = np.full(n, 3, np.uint8)
    # ~4% of the LEAF entities have no TransformComponent (a parent without one makes the reference throw)
    is_parent = np.zeros(n, bool)
    is_parent[parent[parent >= 0]] = True
    no_t = (hash_u01(seed, 210, idx) < 0.04) & ~is_parent
    tflags[no_t] = 0
    parent = parent.copy()
    parent[no_t] = -1
    # ~10% keep modelWithAncestors == false (transform.hpp:200)
    tflags[(hash_u01(seed, 211, idx) < 0.10) & ~no_t] = 1
    # a few axis-aligned / identity rotations and unit scales (exact zeros in the rotation matrix)
    ident = hash_u01(seed, 212, idx) < 0.05
    rot[ident] = np.array([0.0, 0.0, 0.0, 1.0], np.float32)
    half_turn = hash_u01(seed, 213, idx) < 0.03
    rot[half_turn] = np.array([0.0, 1.0, 0.0, 0.0], np.float32)
    unit = hash_u01(seed, 214, idx) < 0.05
    scl[unit] = 1.0
    neg = hash_u01(seed, 215, idx) < 0.03
    scl[neg, 0] *= -1.0
    inactive = np.nonzero((hash_u01(seed, 216, idx) < 0.03) & ~no_t)[0].astype(np.uint32)

    bucket = hash_u01(seed, 220, idx)
    kinds = [RT_OPAQUE, RT_TRANSLUCENT, RT_COLOR, RT_OIT, RT_REFRACTED if single_translucent else RT_TRANSLUCENT]
    edges = [0.0, 0.35, 0.55, 0.70, 0.80, 0.92]
    if with_ui:
        kinds.append(RT_UI)
        edges.append(1.0)
    else:
        edges[-1] = 1.0
    pools = []
    for k, rt in enumerate(kinds):
        sel = np.nonzero((bucket >= edges[k]) & (bucket < edges[k + 1]))[0].astype(np.uint32)
        m = sel.size
        j = sel.astype(np.uint64)
        half = np.stack([0.2 + 1.3 * hash_u01(seed, 230 + c, j) for c in range(3)], axis=1).astype(np.float32)
        centre = np.stack([hash_u01(seed, 233 + c, j) - np.float32(0.5) for c in range(3)], axis=1).astype(np.float32)
        aabb = np.concatenate([centre - half, centre + half], axis=1).astype(np.float32)
        deg = hash_u01(seed, 236, j)
        aabb[deg < 0.02, 3:] = aabb[deg < 0.02, :3]                  # zero size: skipped (mesh.cpp:140-142)
        flat = (deg >= 0.02) & (deg < 0.04)
        aabb[flat, 4] = aabb[flat, 1]                                # zero in one axis only: NOT skipped
        inv = (deg >= 0.04) & (deg < 0.05)
        aabb[inv] = aabb[inv][:, [3, 4, 5, 0, 1, 2]]                 # inverted box: skipped
        enabled = (hash_u01(seed, 237, j) >= 0.05).astype(np.uint8)
        ready = None
        if with_ready and k in (1, 2):
            ready = (hash_u01(seed, 238, j) * 4.0).astype(np.uint8)  # 0..3 instances, 0 = not ready
        pools.append(PoolDesc(rt, sel, aabb, enabled, ready, ref_stride(k)))
    scene = SceneDesc(pos, rot, scl, parent, tflags, pools, inactive,
                      np.array([3.5, -1.25, 2.0], np.float32), name=f"mixed{seed}")
    return scene


def mixed_views(yaw: float = 0.4, with_ui: bool = True):
    """Cascades + main view; the main view carries a UI frustum (an orthographic box around the origin)."""
    views, vps = V.camera_and_cascades(yaw, -0.15, 1.1, 16 / 9, 0.05, 60.0, (0.1, 0.3, 1.0))
    if with_ui:
        ui_vp = V.ortho_rev_z((-40.0, 40.0), (-6.0, 6.0), (-50.0, 50.0))
        ui = V.frustum_planes(ui_vp)
        views["uiPlanes"][-1] = ui
        views["uiPlaneCount"][-1] = 6
    return views


def few_planes_views():
    """A main view whose frustum has fewer than 6 planes (Frustum::setPlaneCount, frustum.hpp:71-76)."""
    views, _ = V.perspective_views([(0.2, 0.0), (2.0, 0.3)], 1.0, 1.5, 0.1)
    views["planeCount"][0] = 4
    views["planeCount"][1] = 1
    return views
