"""GPU-side differential fuzzing: the CUDA path (through the C ABI) against the oracle on random mixed scenes — the same
generator and parameter ranges as tests/test_oracle_vs_ref.py::test_fuzz_mixed_scenes uses for oracle-vs-reference —
plus seeds at the world extents of C4 / C5 (+-1600 / +-3200) and scenes whose boxes hug the frustum planes, which is
where a conservative classifier (garden_b200/csrc/cull.cu: prepass sphere bound, per-view band) could go wrong."""
import numpy as np
import pytest

from garden_b200 import scenes, views as V
from garden_b200.layout import RT_OPAQUE, RT_TRANSLUCENT
from garden_b200.scenes import PoolDesc, SceneDesc

from common import OracleRun, aos_inputs, compare_gpu_to_oracle
from edge_scenes import mixed_scene, mixed_views

pytestmark = pytest.mark.gpu


def _run_and_compare(scene, views, tag, ready=None):
    from garden_b200.binding import ScenePrep
    t, pools = aos_inputs(scene)
    rts = [p.render_type for p in scene.pools]
    sp = ScenePrep(0)
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    sp.set_pool_count(len(pools))
    for k, m in enumerate(pools):
        sp.set_mesh_pool(k, rts[k], m, m.dtype.itemsize, m.size, ready_counts=None if ready is None else ready[k])
    sp.set_views(views, scene.camera_pos)
    sp.run()
    orun = OracleRun((t, t.dtype.itemsize, t.size), [(m, m.dtype.itemsize, m.size) for m in pools], rts, views, scene.camera_pos,
                     ready=ready)
    compare_gpu_to_oracle(sp, orun, rts, views, tag)
    # isVisible of the main view, every slot
    main = [v for v in range(views.size) if int(views[v]["shadowPass"]) < 0]
    if main:
        for k, m in enumerate(pools):
            if m.size == 0:
                continue
            raw = m.view(np.uint8).reshape(-1, m.dtype.itemsize)
            raw[:, 15] = 0xFF
            sp.writeback_visible(k, m, m.dtype.itemsize)
            want = orun.views[main[-1]]["visible"][k]
            sel = want != 0xFF
            assert np.array_equal(raw[sel, 15], want[sel]), f"{tag} pool {k}: isVisible differs"
    total = sp.last_visible_total()
    sp.close()
    return total


@pytest.mark.parametrize("seed", list(range(20, 60)))
def test_fuzz_cuda_vs_oracle(sceneprep_lib, oracle_built, seed):
    rng = np.random.default_rng(seed)
    scene = mixed_scene(seed=seed, n=int(rng.integers(300, 2500)), max_depth=int(rng.integers(3, 40)),
                        with_ui=bool(seed % 2), with_ready=bool((seed // 2) % 2), box_half=float(rng.uniform(20.0, 120.0)))
    scene.camera_pos = rng.uniform(-30.0, 30.0, 3).astype(np.float32)
    views = mixed_views(yaw=float(rng.uniform(-3.0, 3.0)), with_ui=bool(seed % 2))
    _run_and_compare(scene, views, f"fuzz seed {seed}", ready=[p.ready for p in scene.pools])


@pytest.mark.parametrize("seed", list(range(60, 76)))
def test_fuzz_large_extents(sceneprep_lib, oracle_built, seed):
    """World extents of C4 (+-1600) and C5 (+-3200), camera far from the origin: magnitudes where the error band of the
    conservative classifier is largest in absolute terms."""
    rng = np.random.default_rng(seed)
    half = 1600.0 if seed % 2 == 0 else 3200.0
    scene = mixed_scene(seed=seed, n=int(rng.integers(4000, 12000)), max_depth=int(rng.integers(3, 16)),
                        with_ui=False, with_ready=False, box_half=half)
    scene.camera_pos = np.array([rng.uniform(-half, half) * 0.8, rng.uniform(-5.0, 5.0), rng.uniform(-half, half) * 0.8],
                                np.float32)
    views = mixed_views(yaw=float(rng.uniform(-3.0, 3.0)), with_ui=False)
    total = _run_and_compare(scene, views, f"large-extent seed {seed}")
    assert total > 0


def plane_hugging_scene(seed: int, views, camera_pos, per_plane: int = 60, extent: float = 1500.0) -> SceneDesc:
    """Boxes whose extreme corner lies within a few ulp .. 1e-3 of a frustum plane, on either side: flat entities with
    identity rotation and unit scale (the world box is position +- 0.5 exactly), chains whose leaf ends up there through
    rotated / scaled parents, and rotated boxes. The exact 8-corner test decides every one of them."""
    rng = np.random.default_rng(seed)
    cam = np.asarray(camera_pos, np.float64)
    pos, rot, scl, parent = [], [], [], []
    deltas = np.array([0.0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4, 1e-3, -1e-3, 3e-2, -3e-2])
    for v in range(views.size):
        for i in range(int(views[v]["planeCount"])):
            pl = views[v]["planes"][i].astype(np.float64)
            ln = np.linalg.norm(pl[:3])
            if ln == 0:
                continue
            n, d = pl[:3] / ln, pl[3] / ln
            for j in range(per_plane):
                q = rng.uniform(-extent, extent, 3) * np.array([1.0, 0.02, 1.0])
                q = q - (n @ q + d) * n                      # on the plane (camera-relative space)
                r = 0.5 * np.abs(n).sum()                    # support of the unit box along n
                side = 1.0 if j % 2 else -1.0
                c = q + n * (side * r + deltas[j % deltas.size] * max(1.0, np.abs(q).max() * 1e-3))
                kind = j % 3
                if kind == 0:                                # flat, axis aligned
                    pos.append(c + cam); rot.append([0, 0, 0, 1]); scl.append([1, 1, 1]); parent.append(-1)
                elif kind == 1:                              # rotated + scaled box at the same centre
                    qv = rng.normal(size=4); qv /= np.linalg.norm(qv)
                    pos.append(c + cam); rot.append(qv * rng.uniform(0.9, 1.1)); scl.append(rng.uniform(0.5, 1.5, 3)); parent.append(-1)
                else:                                        # parent far away, child offset back to the plane
                    qv = rng.normal(size=4); qv /= np.linalg.norm(qv)
                    base = len(pos)
                    ppos = c + cam + rng.uniform(-40, 40, 3)
                    pos.append(ppos); rot.append(qv); scl.append(rng.uniform(0.7, 1.3, 3)); parent.append(-1)
                    # child local position = inverse(parent) * (c + cam): good to ~1e-6, i.e. still hugging the plane
                    x, y, z, w = qv
                    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
                    local = (R.T @ (c + cam - ppos)) / np.asarray(scl[-1])
                    pos.append(local); rot.append([0, 0, 0, 1]); scl.append([1, 1, 1]); parent.append(base)
    n_e = len(pos)
    ent = np.arange(n_e, dtype=np.uint32)
    half = ent[::2], ent[1::2]
    pools = [PoolDesc(RT_OPAQUE, half[0], scenes.unit_aabb(half[0].size)),
             PoolDesc(RT_TRANSLUCENT, half[1], scenes.unit_aabb(half[1].size), stride=64)]
    return SceneDesc(np.asarray(pos, np.float32), np.asarray(rot, np.float32), np.asarray(scl, np.float32),
                     np.asarray(parent, np.int32), np.full(n_e, 3, np.uint8), pools, None,
                     np.asarray(camera_pos, np.float32), name=f"hug{seed}")


@pytest.mark.parametrize("seed", list(range(80, 88)))
def test_fuzz_plane_hugging_boxes(sceneprep_lib, oracle_built, seed):
    rng = np.random.default_rng(seed)
    cam = np.array([rng.uniform(-1000, 1000), rng.uniform(-3, 3), rng.uniform(-1000, 1000)], np.float32)
    if seed % 2:
        views, _ = V.camera_and_cascades(float(rng.uniform(-3, 3)), -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    else:
        views, _ = V.perspective_views([(float(rng.uniform(-3, 3)), 0.05), (float(rng.uniform(-3, 3)), -0.3)], 1.3, 16 / 9, 0.01)
    scene = plane_hugging_scene(seed, views, cam, extent=3000.0 if seed % 4 < 2 else 300.0)
    total = _run_and_compare(scene, views, f"plane-hugging seed {seed}")
    assert total > 0
