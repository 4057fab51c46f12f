"""ctypes access to the checkers under oracle/ (TEST INFRASTRUCTURE — never imported by garden_b200/).

  RefEngine    oracle/_ref/libgarden_ref_{parity,stock}.so — the reference's own translation units + harness
  Oracle       oracle/libsceneprep_oracle.so               — the plain-C restatement
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np

from garden_b200.layout import RECORD_DTYPE, VIEW_DTYPE

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_PARITY = ORACLE_DIR / "_ref" / "libgarden_ref_parity.so"
REF_STOCK = ORACLE_DIR / "_ref" / "libgarden_ref_stock.so"
ORACLE_LIB = ORACLE_DIR / "libsceneprep_oracle.so"

_vp, _u32, _i32 = C.c_void_p, C.c_uint32, C.c_int
_pp, _pu32 = C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)


def build_oracle():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "oracle"], check=True)
    return ORACLE_LIB


def ref_available(kind: str = "parity") -> bool:
    return (REF_PARITY if kind == "parity" else REF_STOCK).exists()


def _records(ptr: C.c_void_p, count: int) -> np.ndarray:
    if count == 0 or not ptr.value:
        return np.zeros(0, dtype=RECORD_DTYPE)
    buf = (C.c_uint8 * (count * 64)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=RECORD_DTYPE, count=count).copy()


class RefEngine:
    """The reference itself, driven through its real ECS API. One instance per process at a time (ECS singletons)."""

    def __init__(self, kind: str = "parity", threads: int = -1, use_oit: bool = False):
        path = REF_PARITY if kind == "parity" else REF_STOCK
        # mpmt's setMainThread() aborts when called twice in one process image (file-static flag in
        # libraries/logy/libraries/mpmt/source/thread.c:178-194), so every engine instance loads a private copy of the
        # library: a dlopen of a distinct file gets fresh statics.
        with tempfile.NamedTemporaryFile(prefix="garden_ref_", suffix=".so", delete=False) as tmp:
            shutil.copyfile(path, tmp.name)
        try:
            self.lib = C.CDLL(tmp.name)
        finally:
            os.unlink(tmp.name)
        L = self.lib
        L.ref_init.argtypes = [_i32, _i32]
        L.ref_add_pool.argtypes = [_i32]
        L.ref_set_pool_draw_ready.argtypes = [_i32, _i32]
        L.ref_set_pool_draw_ready2.argtypes = [_i32, _i32, _i32]
        L.ref_create_entities.argtypes = [_u32, _vp, _vp, _vp, _vp, _vp]
        L.ref_update_trs.argtypes = [_u32, _vp, _vp, _vp, _vp]
        L.ref_set_active.argtypes = [_u32, _vp, _i32]
        L.ref_add_meshes.argtypes = [_i32, _u32, _vp, _vp, _vp, _vp]
        L.ref_destroy_entities.argtypes = [_u32, _vp]
        L.ref_transform_pool.argtypes = [_pp, _pu32, _pu32]
        L.ref_mesh_pool.argtypes = [_i32, _pp, _pu32, _pu32, _pu32]
        L.ref_pool_ready_counts.argtypes = [_i32, _pu32]
        L.ref_pool_ready_counts.restype = C.c_void_p
        L.ref_set_camera.argtypes = [_vp]
        L.ref_frustum_planes.argtypes = [_vp, _vp]
        L.ref_prepare.argtypes = [_vp, _i32, _vp, _i32, _vp, _i32]
        L.ref_unsorted_buffer_count.restype = _u32
        L.ref_sorted_buffer_count.restype = _u32
        L.ref_get_unsorted.argtypes = [_u32, _pp, _pu32, _pu32]
        L.ref_get_sorted_counts.argtypes = [_u32, _pu32, _pu32]
        L.ref_get_sorted.argtypes = [_i32, _pp, _pu32]
        L.ref_calc_model.argtypes = [_u32, _vp, _vp]
        L.ref_time_frames.argtypes = [_u32, _vp, _vp, _vp, _vp, _u32, _vp, _vp]
        L.ref_instance_mvp.argtypes = [_vp, _vp, _u32, _vp]
        L.ref_animate.argtypes = [_u32, _vp, _vp, _vp, _vp, _vp]
        if L.ref_init(threads, 1 if use_oit else 0) != 0:
            raise RuntimeError("reference engine is already initialised in this process")
        self.pool_strides = []

    def close(self):
        if self.lib is not None:
            self.lib.ref_shutdown()
            self.lib = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def thread_count(self) -> int:
        return self.lib.ref_thread_count()

    def load_scene(self, scene):
        """Builds `scene` (garden_b200.scenes.SceneDesc) through createEntity / add<> / setParent / setActive."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pos, rot, scl = f32(scene.position), f32(scene.rotation), f32(scene.scale)
        parent = np.ascontiguousarray(scene.parent, dtype=np.int32)
        flags = np.ascontiguousarray(scene.tflags, dtype=np.uint8)
        self.lib.ref_create_entities(scene.entity_count, pos.ctypes.data, rot.ctypes.data, scl.ctypes.data,
                                     parent.ctypes.data, flags.ctypes.data)
        for pd in scene.pools:
            k = self.lib.ref_add_pool(pd.render_type)
            assert k >= 0
            stride = 48 + 16 * (k % 3)
            self.pool_strides.append(stride)
            ent = np.ascontiguousarray(pd.entity_index, dtype=np.uint32)
            aabb = f32(pd.aabb)
            en = None if pd.enabled is None else np.ascontiguousarray(pd.enabled, dtype=np.uint8)
            rd = None if pd.ready is None else np.ascontiguousarray(pd.ready, dtype=np.uint8)
            self.lib.ref_add_meshes(k, ent.size, ent.ctypes.data, aabb.ctypes.data,
                                    None if en is None else en.ctypes.data, None if rd is None else rd.ctypes.data)
            shadow_ready = pd.draw_ready if pd.draw_ready_shadow is None else pd.draw_ready_shadow
            if not pd.draw_ready or not shadow_ready:
                self.lib.ref_set_pool_draw_ready2(k, 1 if pd.draw_ready else 0, 1 if shadow_ready else 0)
        if scene.inactive is not None and len(scene.inactive):
            idx = np.ascontiguousarray(scene.inactive, dtype=np.uint32)
            self.lib.ref_set_active(idx.size, idx.ctypes.data, 0)
        cam = f32(scene.camera_pos)
        self.lib.ref_set_camera(cam.ctypes.data)

    def set_active(self, indices, active: bool):
        """TransformComponent::setActive for the given entity INDICES (0-based creation order), in order."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        self.lib.ref_set_active(idx.size, idx.ctypes.data, 1 if active else 0)

    def animate(self, indices, flags, frame_a, frame_b, t):
        """TransformSystem::animateAsync for entity INDICES (0-based); frames [n, 10] = position, scale, rotation."""
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        fa = np.ascontiguousarray(frame_a, dtype=np.float32); fb = np.ascontiguousarray(frame_b, dtype=np.float32)
        tt = np.ascontiguousarray(t, dtype=np.float32)
        self.lib.ref_animate(idx.size, idx.ctypes.data, fl.ctypes.data, fa.ctypes.data, fb.ctypes.data, tt.ctypes.data)

    def instance_mvp(self, view_proj, records: np.ndarray) -> np.ndarray:
        """(float4x4)(viewProj * f32x4x4(bakedModel, (0,0,0,1))) per record, by the reference's own math library."""
        vp = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        rec = np.ascontiguousarray(records)
        out = np.zeros((rec.size, 16), dtype=np.float32)
        if rec.size:
            self.lib.ref_instance_mvp(vp.ctypes.data, rec.ctypes.data, rec.size, out.ctypes.data)
        return out

    def destroy_entities(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        self.lib.ref_destroy_entities(idx.size, idx.ctypes.data)

    def update_trs(self, indices, pos, rot, scl):
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        pos, rot, scl = f32(pos), f32(rot), f32(scl)
        self.lib.ref_update_trs(idx.size, idx.ctypes.data, pos.ctypes.data, rot.ctypes.data, scl.ctypes.data)

    def set_camera(self, cam):
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        self.lib.ref_set_camera(cam.ctypes.data)

    def transform_pool(self):
        """(address, stride, occupancy) of the live LinearPool<TransformComponent> memory."""
        ptr, stride, occ = C.c_void_p(), _u32(), _u32()
        self.lib.ref_transform_pool(C.byref(ptr), C.byref(stride), C.byref(occ))
        return ptr.value, stride.value, occ.value

    def mesh_pool(self, pool: int):
        """(address, stride, occupancy, count) of a mesh system's live component pool."""
        ptr, stride, occ, cnt = C.c_void_p(), _u32(), _u32(), _u32()
        self.lib.ref_mesh_pool(pool, C.byref(ptr), C.byref(stride), C.byref(occ), C.byref(cnt))
        return ptr.value, stride.value, occ.value, cnt.value

    def pool_ready_counts(self, pool: int):
        size = _u32()
        ptr = self.lib.ref_pool_ready_counts(pool, C.byref(size))
        if size.value == 0:
            return None
        return np.frombuffer((C.c_uint8 * size.value).from_address(ptr), dtype=np.uint8).copy()

    def pool_bytes(self, pool: int) -> np.ndarray:
        ptr, stride, occ, _ = self.mesh_pool(pool)
        if occ == 0:
            return np.zeros(0, np.uint8)
        return np.frombuffer((C.c_uint8 * (stride * occ)).from_address(ptr), dtype=np.uint8)

    def transform_bytes(self) -> np.ndarray:
        ptr, stride, occ = self.transform_pool()
        if occ == 0:
            return np.zeros(0, np.uint8)
        return np.frombuffer((C.c_uint8 * (stride * occ)).from_address(ptr), dtype=np.uint8)

    def frustum_planes(self, view_proj) -> np.ndarray:
        m = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        out = np.zeros((6, 4), dtype=np.float32)
        self.lib.ref_frustum_planes(m.ctypes.data, out.ctypes.data)
        return out

    def prepare(self, view: np.ndarray):
        """One reference prepareMeshes call for one element of a VIEW_DTYPE array."""
        v = np.ascontiguousarray(view, dtype=VIEW_DTYPE).reshape(1)[0]
        planes = np.ascontiguousarray(v["planes"])
        ui = np.ascontiguousarray(v["uiPlanes"])
        off = np.ascontiguousarray(v["cameraOffset"])
        has_ui = int(v["uiPlaneCount"]) > 0
        self.lib.ref_prepare(planes.ctypes.data, int(v["planeCount"]), ui.ctypes.data if has_ui else None,
                             int(v["uiPlaneCount"]), off.ctypes.data, int(v["shadowPass"]))

    def unsorted_buffer_count(self) -> int:
        return self.lib.ref_unsorted_buffer_count()

    def sorted_buffer_count(self) -> int:
        return self.lib.ref_sorted_buffer_count()

    def get_unsorted(self, buffer: int):
        ptr, draw, inst = C.c_void_p(), _u32(), _u32()
        self.lib.ref_get_unsorted(buffer, C.byref(ptr), C.byref(draw), C.byref(inst))
        return _records(ptr, draw.value), draw.value, inst.value

    def get_sorted_counts(self, buffer: int):
        draw, inst = _u32(), _u32()
        self.lib.ref_get_sorted_counts(buffer, C.byref(draw), C.byref(inst))
        return draw.value, inst.value

    def get_sorted(self, which: int):
        ptr, draw = C.c_void_p(), _u32()
        self.lib.ref_get_sorted(which, C.byref(ptr), C.byref(draw))
        return _records(ptr, draw.value), draw.value

    def calc_model(self, entity_index: int, cam) -> np.ndarray:
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        out = np.zeros(16, dtype=np.float32)
        self.lib.ref_calc_model(entity_index, cam.ctypes.data, out.ctypes.data)
        return out

    def time_frames(self, views: np.ndarray, frames: int):
        views = np.ascontiguousarray(views, dtype=VIEW_DTYPE)
        planes = np.ascontiguousarray(views["planes"], dtype=np.float32)
        counts = np.ascontiguousarray(views["planeCount"], dtype=np.uint8)
        offs = np.ascontiguousarray(views["cameraOffset"], dtype=np.float32)
        passes = np.ascontiguousarray(views["shadowPass"], dtype=np.int8)
        ms = np.zeros(frames, dtype=np.float64)
        vis = np.zeros(frames, dtype=np.uint64)
        self.lib.ref_time_frames(views.size, planes.ctypes.data, counts.ctypes.data, offs.ctypes.data,
                                 passes.ctypes.data, frames, ms.ctypes.data, vis.ctypes.data)
        return ms, vis


class Oracle:
    """The plain-C restatement (oracle/sceneprep_oracle.c)."""

    def __init__(self):
        if not ORACLE_LIB.exists():
            build_oracle()
        self.lib = C.CDLL(str(ORACLE_LIB))
        L = self.lib
        L.oracle_create.restype = _vp
        L.oracle_destroy.argtypes = [_vp]
        L.oracle_set_transforms.argtypes = [_vp, _vp, _u32, _u32]
        L.oracle_set_pool.argtypes = [_vp, _u32, _u32, _u32, _vp, _u32, _u32, _u32, _vp, _u32]
        L.oracle_set_pool_count.argtypes = [_vp, _u32]
        L.oracle_set_pool_draw_ready.argtypes = [_vp, _u32, _u32, _u32]
        L.oracle_set_camera.argtypes = [_vp, _vp]
        L.oracle_prepare.argtypes = [_vp, _vp, _i32]
        L.oracle_unsorted_buffer_count.argtypes = [_vp]
        L.oracle_unsorted_buffer_count.restype = _u32
        L.oracle_sorted_buffer_count.argtypes = [_vp]
        L.oracle_sorted_buffer_count.restype = _u32
        L.oracle_get_unsorted.argtypes = [_vp, _u32, _pp, _pu32, _pu32]
        L.oracle_get_sorted_counts.argtypes = [_vp, _u32, _pu32, _pu32]
        L.oracle_get_sorted.argtypes = [_vp, _i32, _pp, _pu32]
        L.oracle_get_visible.argtypes = [_vp, _u32]
        L.oracle_get_visible.restype = _vp
        L.oracle_calc_model.argtypes = [_vp, _u32, _vp, _vp]
        L.oracle_frustum_planes.argtypes = [_vp, _vp]
        L.oracle_mat_mul.argtypes = [_vp, _vp, _vp]
        L.oracle_local_model.argtypes = [_vp, _vp, _vp, _vp]
        L.oracle_instance_mvp.argtypes = [_vp, _vp, _u32, _vp]
        L.oracle_set_active.argtypes = [_vp, _u32, _u32, _vp, _u32, _i32]
        L.oracle_set_active.restype = _i32
        L.oracle_animate.argtypes = [_vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]
        L.oracle_animate.restype = _i32
        self.h = L.oracle_create()
        self._keep = []
        self.occupancy = {}

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_transforms(self, aos, stride: int, occupancy: int):
        addr = aos.ctypes.data if isinstance(aos, np.ndarray) else aos
        self._keep.append(aos)
        self.lib.oracle_set_transforms(self.h, addr, stride, occupancy)

    def set_pool(self, index: int, render_type: int, aos, stride: int, occupancy: int, count: int | None = None,
                 draw_ready: bool = True, ready_counts=None):
        addr = aos.ctypes.data if isinstance(aos, np.ndarray) else aos
        self._keep.append(aos)
        rc_ptr, rc_size = None, 0
        if ready_counts is not None:
            ready_counts = np.ascontiguousarray(ready_counts, dtype=np.uint8)
            self._keep.append(ready_counts)
            rc_ptr, rc_size = ready_counts.ctypes.data, ready_counts.size
        self.lib.oracle_set_pool(self.h, index, render_type, 1 if draw_ready else 0, addr, stride, occupancy,
                                 occupancy if count is None else count, rc_ptr, rc_size)
        self.occupancy[index] = occupancy

    def set_pool_count(self, n: int):
        self.lib.oracle_set_pool_count(self.h, n)

    def set_pool_draw_ready(self, pool: int, ready_main: bool, ready_shadow: bool):
        self.lib.oracle_set_pool_draw_ready(self.h, pool, 1 if ready_main else 0, 1 if ready_shadow else 0)

    def set_camera(self, cam):
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        self.lib.oracle_set_camera(self.h, cam.ctypes.data)

    def prepare(self, view: np.ndarray, write_visible: bool = False) -> int:
        v = np.ascontiguousarray(view, dtype=VIEW_DTYPE).reshape(1)
        return self.lib.oracle_prepare(self.h, v.ctypes.data, 1 if write_visible else 0)

    def unsorted_buffer_count(self) -> int:
        return self.lib.oracle_unsorted_buffer_count(self.h)

    def sorted_buffer_count(self) -> int:
        return self.lib.oracle_sorted_buffer_count(self.h)

    def get_unsorted(self, buffer: int):
        ptr, draw, inst = C.c_void_p(), _u32(), _u32()
        self.lib.oracle_get_unsorted(self.h, buffer, C.byref(ptr), C.byref(draw), C.byref(inst))
        return _records(ptr, draw.value), draw.value, inst.value

    def get_sorted_counts(self, buffer: int):
        draw, inst = _u32(), _u32()
        self.lib.oracle_get_sorted_counts(self.h, buffer, C.byref(draw), C.byref(inst))
        return draw.value, inst.value

    def get_sorted(self, which: int):
        ptr, draw = C.c_void_p(), _u32()
        self.lib.oracle_get_sorted(self.h, which, C.byref(ptr), C.byref(draw))
        return _records(ptr, draw.value), draw.value

    def get_visible(self, pool: int) -> np.ndarray:
        n = self.occupancy[pool]
        ptr = self.lib.oracle_get_visible(self.h, pool)
        if n == 0:
            return np.zeros(0, np.uint8)
        return np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=np.uint8).copy()

    def calc_model(self, slot: int, cam) -> np.ndarray:
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        out = np.zeros(16, dtype=np.float32)
        rc = self.lib.oracle_calc_model(self.h, slot, cam.ctypes.data, out.ctypes.data)
        assert rc == 0
        return out

    def frustum_planes(self, view_proj) -> np.ndarray:
        m = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        out = np.zeros((6, 4), dtype=np.float32)
        self.lib.oracle_frustum_planes(m.ctypes.data, out.ctypes.data)
        return out

    def instance_mvp(self, view_proj, records: np.ndarray) -> np.ndarray:
        vp = np.ascontiguousarray(view_proj, dtype=np.float32).reshape(16)
        rec = np.ascontiguousarray(records)
        out = np.zeros((rec.size, 16), dtype=np.float32)
        if rec.size:
            self.lib.oracle_instance_mvp(vp.ctypes.data, rec.ctypes.data, rec.size, out.ctypes.data)
        return out

    def animate(self, transforms: np.ndarray, stride: int, occupancy: int, entity_ids, flags, frame_a, frame_b, t) -> int:
        """TransformSystem::animateAsync on AoS transform bytes IN PLACE (entity ids are 1-based)."""
        ids = np.ascontiguousarray(entity_ids, dtype=np.uint32)
        fl = np.ascontiguousarray(flags, dtype=np.uint8)
        fa = np.ascontiguousarray(frame_a, dtype=np.float32); fb = np.ascontiguousarray(frame_b, dtype=np.float32)
        tt = np.ascontiguousarray(t, dtype=np.float32)
        return self.lib.oracle_animate(transforms.ctypes.data, stride, occupancy, ids.size, ids.ctypes.data, fl.ctypes.data,
                                       fa.ctypes.data, fb.ctypes.data, tt.ctypes.data)

    def set_active(self, transforms: np.ndarray, stride: int, occupancy: int, entity_ids, active: bool) -> int:
        """TransformComponent::setActive on AoS transform bytes IN PLACE (entity ids are 1-based)."""
        ids = np.ascontiguousarray(entity_ids, dtype=np.uint32)
        return self.lib.oracle_set_active(transforms.ctypes.data, stride, occupancy, ids.ctypes.data, ids.size, 1 if active else 0)
