"""Shared test helpers: run a scene through the oracle / the reference / the CUDA path and compare draw lists."""
from __future__ import annotations

import numpy as np

from garden_b200.layout import RECORD_DTYPE, RT_OIT, RT_TRANSLUCENT, RT_UI
from garden_b200.scenes import SceneDesc, build_aos

from reflib import Oracle


def aos_inputs(scene: SceneDesc, strides=None):
    """numpy-built ECS-shaped memory for `scene` (no reference needed). strides: per pool component size."""
    if strides is not None:
        for pd, s in zip(scene.pools, strides):
            pd.stride = s
    t, pools = build_aos(scene)
    return t, pools


def canonical(records: np.ndarray, descending: bool, sort: bool = True) -> np.ndarray:
    """Canonical order for lists whose tie order the reference leaves unspecified (SURVEY.md finding 4):
    key order first, then (bufferIndex, componentOffset) ascending."""
    if records.size == 0 or not sort:
        return records
    key = records["distanceSq"].astype(np.float64)
    if descending:
        key = -key
    idx = np.lexsort((records["componentOffset"], records["bufferIndex"], key))
    return records[idx]


def assert_records_equal(a: np.ndarray, b: np.ndarray, what: str, ignore_buffer_index: bool = False):
    assert a.size == b.size, f"{what}: draw count {a.size} != {b.size}"
    if a.size == 0:
        return
    ka, kb = a["distanceSq"].view(np.uint32), b["distanceSq"].view(np.uint32)
    assert np.array_equal(ka, kb), f"{what}: key bits differ at {np.nonzero(ka != kb)[0][:8]}"
    assert np.array_equal(a["componentOffset"], b["componentOffset"]), \
        f"{what}: draw order differs at {np.nonzero(a['componentOffset'] != b['componentOffset'])[0][:8]}"
    ma, mb = a["bakedModel"].view(np.uint32), b["bakedModel"].view(np.uint32)
    assert np.array_equal(ma, mb), f"{what}: bakedModel bits differ in {np.count_nonzero((ma != mb).any(axis=1))} records"
    if not ignore_buffer_index:
        assert np.array_equal(a["bufferIndex"], b["bufferIndex"]), f"{what}: bufferIndex differs"


class OracleRun:
    """Oracle results of a whole frame (all views), in the shape the GPU results are compared against."""

    def __init__(self, tf, pools, render_types, views, camera_pos, ready=None, draw_ready=None, counts=None,
                 write_visible=False):
        self.o = Oracle()
        t_addr, t_stride, t_occ = tf
        self.o.set_transforms(t_addr, t_stride, t_occ)
        for k, (addr, stride, occ) in enumerate(pools):
            self.o.set_pool(k, render_types[k], addr, stride, occ, None if counts is None else counts[k],
                            True if draw_ready is None else draw_ready[k], None if ready is None else ready[k])
        self.o.set_pool_count(len(pools))
        self.o.set_camera(camera_pos)
        self.views = []
        for v in range(views.size):
            rc = self.o.prepare(views[v], write_visible)
            assert rc == 0, f"oracle_prepare failed with {rc}"
            nb = self.o.unsorted_buffer_count()
            ns = self.o.sorted_buffer_count()
            res = {
                "unsorted": [self.o.get_unsorted(b) for b in range(nb)],
                "sorted_counts": [self.o.get_sorted_counts(b) for b in range(ns)],
                "trans": self.o.get_sorted(0),
                "ui": self.o.get_sorted(1),
            }
            if int(views[v]["shadowPass"]) < 0:
                res["visible"] = [self.o.get_visible(k) for k in range(len(pools))]
            self.views.append(res)


def compare_gpu_to_oracle(sp, orun: OracleRun, render_types, views, tag=""):
    """sp: garden_b200.binding.ScenePrep after run(); checks every list of every view bit for bit."""
    unsorted_types = [rt for rt in render_types if rt not in (RT_TRANSLUCENT, RT_UI)]
    for v in range(views.size):
        ov = orun.views[v]
        assert sp.unsorted_buffer_count(v) == len(ov["unsorted"]), f"{tag} view {v}: unsorted buffer count"
        assert sp.sorted_buffer_count(v) == len(ov["sorted_counts"]), f"{tag} view {v}: sorted buffer count"
        for b, (orec, odraw, oinst) in enumerate(ov["unsorted"]):
            grec, gdraw, ginst = sp.get_unsorted(v, b)
            assert (gdraw, ginst) == (odraw, oinst), f"{tag} view {v} buffer {b}: counts {(gdraw, ginst)} != {(odraw, oinst)}"
            assert_records_equal(grec, orec, f"{tag} view {v} unsorted buffer {b}", ignore_buffer_index=False)
        for b, (odraw, oinst) in enumerate(ov["sorted_counts"]):
            assert sp.get_sorted_counts(v, b) == (odraw, oinst), f"{tag} view {v} sorted buffer {b} counts"
        for which, name in ((0, "trans"), (1, "ui")):
            orec, odraw = ov[name]
            grec, gdraw = sp.get_sorted(v, which)
            assert gdraw == odraw, f"{tag} view {v} {name}: draw count {gdraw} != {odraw}"
            assert_records_equal(grec, orec, f"{tag} view {v} {name} list")
