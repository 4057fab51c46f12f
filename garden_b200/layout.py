"""Byte layouts of the reference's ECS components and draw-list records, as numpy structured dtypes.

These mirror (Release, non-editor build, verified with offsetof against the real headers by oracle/ref_harness.cpp):
  TransformComponent   include/garden/system/transform.hpp:31-60          80 bytes
  MeshRenderComponent  include/garden/system/render/mesh.hpp:45-55        48 bytes (derived components are larger)
  UnsortedMesh/SortedMesh  include/garden/system/render/mesh.hpp:191-205  64 bytes
"""
from __future__ import annotations

import numpy as np

# MeshRenderType, mesh.hpp:30-40
RT_COLOR, RT_OPAQUE, RT_TRANSLUCENT, RT_OIT, RT_REFRACTED, RT_TRANS_DEPTH, RT_UI = range(7)

TRANSFORM_DTYPE = np.dtype({
    "names": ["entity", "parent", "uid", "position", "childCount", "scale", "childCapacity", "rotation", "childs",
              "selfActive", "ancestorsActive", "modelWithAncestors"],
    "formats": ["<u4", "<u4", "<u8", ("<f4", 3), "<u4", ("<f4", 3), "<u4", ("<f4", 4), "<u8", "u1", "u1", "u1"],
    "offsets": [0, 4, 8, 16, 28, 32, 44, 48, 64, 72, 73, 74],
    "itemsize": 80,
})


def mesh_dtype(stride: int = 48) -> np.dtype:
    """MeshRenderComponent (or a larger derived component with `stride` bytes)."""
    assert stride >= 48 and stride % 16 == 0
    return np.dtype({
        "names": ["entity", "isEnabled", "isVisible", "aabbMin", "aabbMax"],
        "formats": ["<u4", "u1", "u1", ("<f4", 4), ("<f4", 4)],
        "offsets": [0, 14, 15, 16, 32],
        "itemsize": stride,
    })


RECORD_DTYPE = np.dtype({
    "names": ["componentOffset", "bakedModel", "distanceSq", "bufferIndex"],
    "formats": ["<u8", ("<f4", 12), "<f4", "<u4"],
    "offsets": [0, 8, 56, 60],
    "itemsize": 64,
})

# gsp_view, include/garden_sceneprep.h
VIEW_DTYPE = np.dtype({
    "names": ["planes", "planeCount", "uiPlanes", "uiPlaneCount", "cameraOffset", "shadowPass"],
    "formats": [("<f4", (6, 4)), "<u4", ("<f4", (6, 4)), "<u4", ("<f4", 4), "<i4"],
    "offsets": [0, 96, 100, 196, 200, 216],
    "itemsize": 220,
})


def make_views(planes, camera_offsets, shadow_passes, plane_counts=None, ui_planes=None) -> np.ndarray:
    """Packs per-view inputs into an array of gsp_view / OracleView structs (identical layouts)."""
    planes = np.asarray(planes, dtype=np.float32).reshape(-1, 6, 4)
    n = planes.shape[0]
    views = np.zeros(n, dtype=VIEW_DTYPE)
    views["planes"] = planes
    views["planeCount"] = 6 if plane_counts is None else np.asarray(plane_counts, dtype=np.uint32)
    views["cameraOffset"] = np.asarray(camera_offsets, dtype=np.float32).reshape(n, 4)
    views["shadowPass"] = np.asarray(shadow_passes, dtype=np.int32)
    if ui_planes is not None:
        ui = np.asarray(ui_planes, dtype=np.float32).reshape(-1, 6, 4)
        views["uiPlanes"] = ui
        views["uiPlaneCount"] = 6
    return views
