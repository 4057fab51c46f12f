"""Host-side view producers: the inputs prepareMeshes receives (viewProj -> frustum planes, cameraOffset).

These stay on the host in the reference and here; they are restated so that bench.py and the tests can
build the same views without the reference tree. Matrices are column-major 4x4 stored as [c0 c1 c2 c3] (16 floats),
the memory order of math::f32x4x4.

  calcPerspProjInfRevZ / calcPerspProjRevZ / calcOrthoProjRevZ   libraries/math/include/math/matrix/projection.hpp:39-99
  lookAt                                                          libraries/math/include/math/matrix/transform.hpp:291-299
  Frustum(viewProj)                                               libraries/math/include/math/frustum.hpp:51-61
  camera-relative view (translation zeroed), viewProj = proj * view   source/system/graphics.cpp:192-243
  calcLightViewProj (cascade viewProj + cameraOffset)             source/system/render/csm.cpp:260-308
"""
from __future__ import annotations

import math

import numpy as np

from .layout import make_views


def _mat(cols) -> np.ndarray:
    """4x4 from 4 columns, stored so that m[i] is column i."""
    return np.array(cols, dtype=np.float64)


def mat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Column-major product a * b (columns of the result are a applied to the columns of b)."""
    return np.stack([sum(a[k] * b[i][k] for k in range(4)) for i in range(4)])


def mat_vec(a: np.ndarray, v) -> np.ndarray:
    return sum(a[k] * v[k] for k in range(4))


def persp_inf_rev_z(fov: float, aspect: float, near: float) -> np.ndarray:
    t = math.tan(fov * 0.5)
    # float4x4(c0r0, c1r0, c2r0, c3r0, ...) row-wise arguments, matrix/float.hpp:504-512
    rows = [[1.0 / (aspect * t), 0, 0, 0], [0, -1.0 / t, 0, 0], [0, 0, 0, near], [0, 0, 1, 0]]
    return np.array(rows, dtype=np.float64).T.copy()


def persp_rev_z(fov: float, aspect: float, near: float, far: float) -> np.ndarray:
    t = math.tan(fov * 0.5)
    rows = [[1.0 / (aspect * t), 0, 0, 0], [0, -1.0 / t, 0, 0],
            [0, 0, near / (near - far), -(near * far) / (near - far)], [0, 0, 1, 0]]
    return np.array(rows, dtype=np.float64).T.copy()


def ortho_rev_z(width, height, depth) -> np.ndarray:
    rows = [[2.0 / (width[1] - width[0]), 0, 0, -(width[1] + width[0]) / (width[1] - width[0])],
            [0, -2.0 / (height[1] - height[0]), 0, (height[1] + height[0]) / (height[1] - height[0])],
            [0, 0, 1.0 / (depth[0] - depth[1]), -depth[1] / (depth[0] - depth[1])],
            [0, 0, 0, 1]]
    return np.array(rows, dtype=np.float64).T.copy()


def _normalize3(v):
    v = np.asarray(v, dtype=np.float64)
    return v / math.sqrt(float(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))


def look_at(frm, to, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    frm = np.asarray(frm, dtype=np.float64)[:3]
    f = _normalize3(np.asarray(to, dtype=np.float64)[:3] - frm)
    s = _normalize3(np.cross(up, f))
    u = np.cross(f, s)
    rows_as_cols = _mat([[*s, -np.dot(s, frm)], [*u, -np.dot(u, frm)], [*f, -np.dot(f, frm)], [0, 0, 0, 1]])
    return rows_as_cols.T.copy()  # transpose4x4


def rotation_view(yaw: float, pitch: float = 0.0) -> np.ndarray:
    """Camera-relative view matrix (rotation only: graphics.cpp:201 zeroes the translation)."""
    direction = (math.sin(yaw) * math.cos(pitch), math.sin(pitch), math.cos(yaw) * math.cos(pitch))
    return look_at((0, 0, 0), direction)


def frustum_planes(view_proj: np.ndarray) -> np.ndarray:
    """Frustum(viewProj): 6 unnormalised planes [normal xyz, distance], in float32 like the reference computes them."""
    m = np.asarray(view_proj, dtype=np.float32)
    t = m.T.copy()  # t[i] = column i of transpose = row i of viewProj
    planes = np.stack([t[3] + t[0], t[3] - t[0], t[3] - t[1], t[3] + t[1], t[2], t[3] - t[2]]).astype(np.float32)
    return planes


def light_view_proj(view: np.ndarray, light_dir, fov: float, aspect: float, near: float, far: float,
                    z_coeff: float = 10.0, shadow_map_size: int = 2048):
    """calcLightViewProj: cascade viewProj and cameraOffset for the camera sub-frustum [near, far]."""
    light_dir = np.asarray(light_dir, dtype=np.float64)
    proj = persp_rev_z(fov, aspect, near, far)
    inv = np.linalg.inv(mat_mul(proj, view).T).T  # inverse of the column-major matrix
    corners = []
    for z in range(2):
        for y in range(2):
            for x in range(2):
                c = mat_vec(inv, (x * 2.0 - 1.0, y * 2.0 - 1.0, float(z), 1.0))
                corners.append(c / c[3])
    center = sum(corners) * (1.0 / 8.0)
    lv = look_at(center[:3] - light_dir, center[:3])
    trf = np.stack([mat_vec(lv, c) for c in corners])
    mn, mx = trf.min(axis=0), trf.max(axis=0)
    mn[2] = mn[2] * z_coeff if mn[2] < 0 else mn[2] / z_coeff
    mx[2] = mx[2] / z_coeff if mx[2] < 0 else mx[2] * z_coeff
    units = (mx[0] - mn[0]) / shadow_map_size
    lcp = mat_vec(lv, center)
    lcp[0] = math.floor(lcp[0] / units) * units
    lcp[2] = math.floor(lcp[2] / units) * units
    snapped = mat_vec(np.linalg.inv(lv.T).T, lcp)
    slv = look_at(snapped[:3] - light_dir, snapped[:3])
    camera_offset = -(light_dir * mn[2] + center[:3])
    lp = ortho_rev_z((mn[0], mx[0]), (mn[1], mx[1]), (mn[2], mx[2]))
    return mat_mul(lp, slv), np.array([*camera_offset, 0.0])


def camera_and_cascades(yaw: float, pitch: float, fov: float, aspect: float, near: float, shadow_distance: float,
                        splits, light_dir=(0.35, -0.85, 0.4)):
    """The frame's views in the reference's order: shadow passes first (renderShadows, mesh.cpp:795-847),
    the main camera view last (mesh.cpp:902). Returns (views array, view_projs)."""
    view = rotation_view(yaw, pitch)
    light = _normalize3(light_dir)
    planes, offsets, passes, vps = [], [], [], []
    cascade_count = len(splits)
    for i in range(cascade_count):  # prepareShadowRender, csm.cpp:311-329
        n = near if i == 0 else shadow_distance * splits[i - 1]
        f = shadow_distance * splits[i]
        vp, off = light_view_proj(view, light, fov, aspect, n, f)
        planes.append(frustum_planes(vp)); offsets.append(off); passes.append(i); vps.append(vp)
    vp = mat_mul(persp_inf_rev_z(fov, aspect, near), view)
    planes.append(frustum_planes(vp)); offsets.append(np.zeros(4)); passes.append(-1); vps.append(vp)
    return make_views(np.stack(planes), np.stack(offsets), passes), vps


def perspective_views(directions, fov: float, aspect: float, near: float):
    """Independent main views (split-screen / cube-probe faces): one perspective frustum per (yaw, pitch)."""
    planes, vps = [], []
    for yaw, pitch in directions:
        vp = mat_mul(persp_inf_rev_z(fov, aspect, near), rotation_view(yaw, pitch))
        planes.append(frustum_planes(vp)); vps.append(vp)
    n = len(planes)
    return make_views(np.stack(planes), np.zeros((n, 4)), [-1] * n), vps
