"""Deterministic synthetic scenes (inputs only) for the configurations in BASELINE.json / SURVEY.md §8d.

Everything is produced by a counter-based integer hash (splitmix64) and exact float32 arithmetic (no libm), so the same
seed gives the same bytes on every machine. A scene is described by flat arrays (`SceneDesc`); `build_aos` lays them
out exactly like the reference's ECS pools would hold them (LinearPool<TransformComponent>, LinearPool<MeshComponent>:
slot i = i-th created component, entity ids 1-based in creation order), which is what the C ABI consumes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .layout import RT_OPAQUE, RT_TRANSLUCENT, TRANSFORM_DTYPE, mesh_dtype

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def hash_u01(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """Uniform float32 in [0, 1) with 24 random bits, a pure function of (seed, stream, index)."""
    with np.errstate(over="ignore"):
        key = index.astype(np.uint64) * np.uint64(0xD1342543DE82EF95) + np.uint64((seed * 1000003 + stream) & 0xFFFFFFFF)
    bits = _splitmix(_splitmix(key)) >> np.uint64(40)
    return (bits.astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


@dataclass
class PoolDesc:
    render_type: int
    entity_index: np.ndarray            # [M] u32: index of the owning entity
    aabb: np.ndarray                    # [M, 6] f32: min xyz, max xyz
    enabled: np.ndarray | None = None   # [M] u8
    ready: np.ndarray | None = None     # [M] u8 ready instance counts (None = frustum-only predicate)
    stride: int = 48
    draw_ready: bool = True
    draw_ready_shadow: bool | None = None  # isDrawReady(shadowPass >= 0); None = same as draw_ready


@dataclass
class SceneDesc:
    position: np.ndarray                # [E, 3] f32
    rotation: np.ndarray                # [E, 4] f32 (xyzw, deliberately not normalised)
    scale: np.ndarray                   # [E, 3] f32
    parent: np.ndarray                  # [E] i32 entity index or -1
    tflags: np.ndarray                  # [E] u8: bit0 has TransformComponent, bit1 modelWithAncestors
    pools: list = field(default_factory=list)
    inactive: np.ndarray | None = None  # entity indices that get setActive(false)
    camera_pos: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))
    name: str = ""

    @property
    def entity_count(self) -> int:
        return int(self.position.shape[0])


def random_trs(seed: int, n: int, box, local: np.ndarray | None = None, local_extent: float = 3.0,
               scale_range=(0.5, 1.5), start: int = 0):
    """TRS for n entities. Roots are uniform in `box` (min xyz, max xyz); entities with local[i] true get a small
    offset from their parent instead. Quaternions are unit quaternions scaled by (0.9 .. 1.1) so normalize4 has work."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    u = [hash_u01(seed, s, idx) for s in range(12)]
    box = np.asarray(box, dtype=np.float32)
    lo, hi = box[:3], box[3:]
    pos = np.stack([lo[k] + u[k] * (hi[k] - lo[k]) for k in range(3)], axis=1).astype(np.float32)
    if local is not None:
        off = np.stack([(u[k] - np.float32(0.5)) * np.float32(2.0 * local_extent) for k in range(3)], axis=1)
        pos = np.where(local[:, None], off, pos).astype(np.float32)
    q = np.stack([u[3 + k] * np.float32(2.0) - np.float32(1.0) for k in range(4)], axis=1).astype(np.float32)
    q[:, 3] += np.float32(0.25)  # keep away from the zero quaternion
    # scale every quaternion into [0.9, 1.1] x unit length without calling libm: divide by an exact power of two
    # bracket of its squared length is overkill; a plain float32 sqrt is IEEE-exact and deterministic.
    norm = np.sqrt((q * q).sum(axis=1, dtype=np.float32)).astype(np.float32)
    norm = np.maximum(norm, np.float32(1e-3))
    q = (q / norm[:, None] * (np.float32(0.9) + u[7] * np.float32(0.2))[:, None]).astype(np.float32)
    s0, s1 = np.float32(scale_range[0]), np.float32(scale_range[1])
    scl = np.stack([s0 + u[8 + k] * (s1 - s0) for k in range(3)], axis=1).astype(np.float32)
    return pos, q, scl


def chain_parents(n: int, depth: int) -> np.ndarray:
    """Chains of depth+1 nodes: every (depth+1)-th entity is a root, each next entity is a child of the previous."""
    idx = np.arange(n, dtype=np.int64)
    parent = idx - 1
    parent[idx % (depth + 1) == 0] = -1
    return parent.astype(np.int32)


def unit_aabb(m: int) -> np.ndarray:
    a = np.empty((m, 6), dtype=np.float32)  # Aabb::one, aabb.hpp
    a[:, :3] = -0.5
    a[:, 3:] = 0.5
    return a


def density_scatter(seed: int, n: int, box, cells: int = 64):
    """Stand-in for fastnoise2 placement (never called by the reference): rejection-free scatter whose density
    follows a value-noise field. Each entity picks a lattice cell with probability proportional to the hashed cell
    weight (by inverse CDF over the sorted cumulative weights) and a uniform offset inside the cell."""
    box = np.asarray(box, dtype=np.float32)
    cell_idx = np.arange(cells * cells, dtype=np.uint64)
    w = hash_u01(seed, 101, cell_idx).astype(np.float64) ** 2 + 0.02
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    idx = np.arange(n, dtype=np.uint64)
    pick = np.searchsorted(cdf, hash_u01(seed, 102, idx).astype(np.float64), side="right").clip(0, cells * cells - 1)
    cx, cz = (pick % cells).astype(np.float32), (pick // cells).astype(np.float32)
    ux, uy, uz = (hash_u01(seed, 103 + k, idx) for k in range(3))
    sx = (box[3] - box[0]) / np.float32(cells)
    sz = (box[5] - box[2]) / np.float32(cells)
    pos = np.stack([box[0] + (cx + ux) * sx, box[1] + uy * (box[4] - box[1]), box[2] + (cz + uz) * sz], axis=1)
    return pos.astype(np.float32)


def make_scene(n: int, depth: int, seed: int, box, translucent_fraction: float = 0.0, scatter: bool = False,
               name: str = "", camera_pos=(0.0, 0.0, 0.0)) -> SceneDesc:
    """N entities, each with a TransformComponent and one mesh component (unit AABB)."""
    parent = chain_parents(n, depth) if depth > 0 else np.full(n, -1, np.int32)
    local = parent >= 0
    pos, rot, scl = random_trs(seed, n, box, local if depth > 0 else None)
    if scatter:
        roots = ~local
        pos[roots] = density_scatter(seed, n, box)[roots]
    tflags = np.full(n, 3, np.uint8)
    ent = np.arange(n, dtype=np.uint32)
    pools = []
    if translucent_fraction > 0.0:
        is_trans = hash_u01(seed, 50, ent.astype(np.uint64)) < np.float32(translucent_fraction)
        o, t = ent[~is_trans], ent[is_trans]
        pools.append(PoolDesc(RT_OPAQUE, o, unit_aabb(o.size)))
        pools.append(PoolDesc(RT_TRANSLUCENT, t, unit_aabb(t.size)))
    else:
        pools.append(PoolDesc(RT_OPAQUE, ent, unit_aabb(n)))
    return SceneDesc(pos, rot, scl, parent, tflags, pools, None, np.asarray(camera_pos, np.float32), name)


def animate_trs(scene: SceneDesc, frame: int, seed: int):
    """Config C3: entities with index % 10 == frame % 10 get new TRS. Returns (entity indices, pos, rot, scale)."""
    n = scene.entity_count
    sel = np.arange(frame % 10, n, 10, dtype=np.uint32)
    local = scene.parent[sel] >= 0
    box = np.concatenate([scene.position[scene.parent < 0].min(axis=0), scene.position[scene.parent < 0].max(axis=0)])
    idx0 = (frame + 1) * n
    u_pos, u_rot, u_scl = random_trs(seed + 7919, sel.size, box, local, start=idx0)
    return sel, u_pos, u_rot, u_scl


def build_aos(scene: SceneDesc):
    """ECS-shaped memory for the scene: (transform pool AoS, [mesh pool AoS ...]).

    Mirrors what the reference's ECS holds after createEntity / add<TransformComponent> / setParent / add<Mesh> in entity
    order: entity id = index + 1, transform slot = rank among entities that have a transform, mesh slot = order of add.
    ancestorsActive is propagated like TransformComponent::setActive does (source/system/transform.cpp:75-127)."""
    e = scene.entity_count
    has_t = (scene.tflags & 1) != 0
    slot_of = np.cumsum(has_t) - 1
    nt = int(has_t.sum())
    t = np.zeros(nt, dtype=TRANSFORM_DTYPE)
    ents = np.nonzero(has_t)[0]
    t["entity"] = ents + 1
    par = scene.parent[ents]
    t["parent"] = np.where(par >= 0, par + 1, 0).astype(np.uint32)
    t["position"] = scene.position[ents]
    t["rotation"] = scene.rotation[ents]
    t["scale"] = scene.scale[ents]
    t["modelWithAncestors"] = ((scene.tflags[ents] >> 1) & 1).astype(np.uint8)
    self_active = np.ones(e, dtype=bool)
    if scene.inactive is not None and len(scene.inactive):
        self_active[np.asarray(scene.inactive, dtype=np.int64)] = False
    # ancestorsActive = all strict ancestors self-active; pointer jumping over the parent array
    anc = np.ones(e, dtype=bool)
    p = scene.parent.astype(np.int64).copy()
    live = p >= 0
    while live.any():
        anc[live] &= self_active[p[live]]
        p[live] = scene.parent[p[live]]
        live = p >= 0
    t["selfActive"] = self_active[ents]
    t["ancestorsActive"] = anc[ents]
    # child bookkeeping the hot path never reads, filled for fidelity: childCount (lane W of posChildCount)
    child_count = np.bincount(par[par >= 0], minlength=e).astype(np.uint32)
    t["childCount"] = child_count[ents]
    cap = np.where(child_count > 0, 1 << np.ceil(np.log2(np.maximum(child_count, 1))).astype(np.uint32), 0)
    t["childCapacity"] = cap[ents].astype(np.uint32)
    pools = []
    for pd in scene.pools:
        m = np.zeros(pd.entity_index.size, dtype=mesh_dtype(pd.stride))
        m["entity"] = pd.entity_index.astype(np.uint32) + 1
        m["isEnabled"] = 1 if pd.enabled is None else pd.enabled
        m["aabbMin"][:, :3] = pd.aabb[:, :3]
        m["aabbMax"][:, :3] = pd.aabb[:, 3:]
        pools.append(m)
    return t, pools


# ---- the BASELINE.json configurations -------------------------------------------------------------------------------
CONFIGS = {
    # name: (N, depth, translucent fraction, scatter, world box half extent xz, y extent)
    "C1": dict(n=10_000, depth=0, trans=0.0, scatter=False, half=100.0, height=10.0),
    "C2": dict(n=1_000_000, depth=4, trans=0.0, scatter=False, half=400.0, height=20.0),
    "C3": dict(n=4_000_000, depth=8, trans=0.25, scatter=False, half=800.0, height=20.0),
    "C4": dict(n=16_000_000, depth=8, trans=0.0, scatter=True, half=1600.0, height=20.0),
    "C5": dict(n=64_000_000, depth=8, trans=0.0, scatter=True, half=3200.0, height=20.0),
}


def config_scene(name: str, n: int | None = None, seed: int = 1234) -> SceneDesc:
    cfg = CONFIGS[name]
    n = cfg["n"] if n is None else n
    # keep the density (entities per unit area) of the named configuration when a reduced N is requested
    half = cfg["half"] * float(np.sqrt(n / cfg["n"]))
    box = (-half, -cfg["height"], -half, half, cfg["height"], half)
    return make_scene(n, cfg["depth"], seed, box, cfg["trans"], cfg["scatter"], name=name)
