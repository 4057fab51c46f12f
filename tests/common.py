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
                 write_visible=False, draw_ready_shadow=None):
        self.o = Oracle()
        t_addr, t_stride, t_occ = tf
        self.o.set_transforms(t_addr, t_stride, t_occ)
        for k, (addr, stride, occ) in enumerate(pools):
            self.o.set_pool(k, render_types[k], addr, stride, occ, None if counts is None else counts[k],
                            True if draw_ready is None else draw_ready[k], None if ready is None else ready[k])
        if draw_ready_shadow is not None:
            for k in range(len(pools)):
                self.o.set_pool_draw_ready(k, True if draw_ready is None else draw_ready[k], draw_ready_shadow[k])
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


def canonical_oit(records: np.ndarray) -> np.ndarray:
    """OIT buffers are filled but never sorted (mesh.cpp:273-277) and the thread pool appends in arbitrary order:
    compare them as sets, ordered by componentOffset."""
    if records.size == 0:
        return records
    return records[np.argsort(records["componentOffset"], kind="stable")]


def ref_frame(ref, views) -> list:
    """Runs every view through the reference engine (tests/reflib.RefEngine) and returns OracleRun-shaped results with
    the lists in canonical order."""
    out = []
    n_pools = len(ref.pool_strides)
    for v in range(views.size):
        ref.prepare(views[v])
        res = {"unsorted": [], "sorted_counts": [], "trans": None, "ui": None}
        for b in range(ref.unsorted_buffer_count()):
            res["unsorted"].append(ref.get_unsorted(b))
        for b in range(ref.sorted_buffer_count()):
            res["sorted_counts"].append(ref.get_sorted_counts(b))
        res["trans"] = ref.get_sorted(0)
        res["ui"] = ref.get_sorted(1) if int(views[v]["shadowPass"]) < 0 else (np.zeros(0, RECORD_DTYPE), 0)
        if int(views[v]["shadowPass"]) < 0:
            res["visible"] = [ref.pool_bytes(k).reshape(-1, ref.pool_strides[k])[:, 15].copy() if ref.mesh_pool(k)[2] else
                              np.zeros(0, np.uint8) for k in range(n_pools)]
        out.append(res)
    return out


def assert_frames_equal(got: list, want: list, render_types, tag: str = "", canonicalise_got: bool = True):
    """got / want: per-view dicts (ref_frame / OracleRun.views shape). Lists are canonicalised before comparing;
    with canonicalise_got=False the `got` lists must ALREADY be in canonical order (the CUDA path's contract:
    draw order with ties broken by (mesh system, slot))."""
    ident = lambda r, *a: r
    canon_g, canon_oit_g = (canonical, canonical_oit) if canonicalise_got else (ident, ident)
    unsorted_types = [rt for rt in render_types if rt not in (RT_TRANSLUCENT, RT_UI)]
    assert len(got) == len(want)
    for v, (g, w) in enumerate(zip(got, want)):
        assert len(g["unsorted"]) == len(w["unsorted"]), f"{tag} view {v}: unsorted buffer count"
        for b, ((grec, gdraw, ginst), (wrec, wdraw, winst)) in enumerate(zip(g["unsorted"], w["unsorted"])):
            assert (gdraw, ginst) == (wdraw, winst), f"{tag} view {v} buffer {b}: counts {(gdraw, ginst)} != {(wdraw, winst)}"
            if unsorted_types[b] == RT_OIT:
                grec, wrec = canon_oit_g(grec), canonical_oit(wrec)
            else:
                grec, wrec = canon_g(grec, False), canonical(wrec, False)
            assert_records_equal(grec, wrec, f"{tag} view {v} unsorted buffer {b}")
        assert g["sorted_counts"] == w["sorted_counts"], f"{tag} view {v}: sorted buffer counts"
        for name in ("trans", "ui"):
            (grec, gdraw), (wrec, wdraw) = g[name], w[name]
            assert gdraw == wdraw, f"{tag} view {v} {name}: draw count {gdraw} != {wdraw}"
            assert_records_equal(canon_g(grec[:gdraw], True), canonical(wrec[:wdraw], True), f"{tag} view {v} {name} list")
        if "visible" in w and "visible" in g:
            for k, (gv, wv) in enumerate(zip(g["visible"], w["visible"])):
                sel = wv != 0xFF  # 0xFF: the reference would not have written this slot's isVisible
                assert np.array_equal(gv[sel] != 0, wv[sel] != 0), f"{tag} view {v} pool {k}: isVisible differs"


def gpu_frame(sp, views, pools_aos=None) -> list:
    """Results of a completed ScenePrep.run() in the OracleRun.views shape, lists exactly as the library returns them.
    pools_aos: list of (array, stride) to run gsp_writeback_visible into (then 'visible' is filled for main views)."""
    out = []
    for v in range(views.size):
        res = {"unsorted": [sp.get_unsorted(v, b) for b in range(sp.unsorted_buffer_count(v))],
               "sorted_counts": [sp.get_sorted_counts(v, b) for b in range(sp.sorted_buffer_count(v))],
               "trans": sp.get_sorted(v, 0), "ui": sp.get_sorted(v, 1)}
        out.append(res)
    if pools_aos is not None:
        mains = [v for v in range(views.size) if int(views[v]["shadowPass"]) < 0]
        if mains:
            vis = []
            for k, (aos, stride) in enumerate(pools_aos):
                raw = aos.view(np.uint8).reshape(-1, stride) if aos.size else np.zeros((0, stride), np.uint8)
                raw[:, 15] = 0xFF
                sp.writeback_visible(k, raw, stride)
                vis.append(raw[:, 15].copy())
            out[mains[-1]]["visible"] = vis  # the last main view wins, like repeated prepareMeshes calls
    return out


class GoldenCase:
    """One tests/golden/*.npz file: the reference's ECS bytes (inputs) and the reference's own results."""

    def __init__(self, path):
        z = np.load(path)
        self.name = str(path).split("/")[-1][:-4]
        self.transforms = np.ascontiguousarray(z["transforms"])
        self.meta = z["pool_meta"]
        self.pools = [np.ascontiguousarray(z[f"pool{k}"]) for k in range(len(self.meta))]
        self.ready = [np.ascontiguousarray(z[f"ready{k}"]) if self.meta[k][5] else None for k in range(len(self.meta))]
        self.render_types = [int(m[0]) for m in self.meta]
        self.views = z["views"]
        self.camera_pos = z["camera_pos"]
        n_unsorted = sum(1 for rt in self.render_types if rt not in (RT_TRANSLUCENT, RT_UI))
        self.frames = []
        for v in range(self.views.size):
            res = {"unsorted": [], "sorted_counts": [tuple(int(x) for x in row) for row in z[f"v{v}_sorted_counts"]]}
            for b in range(n_unsorted):
                c = z[f"v{v}_unsorted{b}_counts"]
                res["unsorted"].append((z[f"v{v}_unsorted{b}"], int(c[0]), int(c[1])))
            res["trans"] = (z[f"v{v}_trans"], int(z[f"v{v}_trans"].size))
            res["ui"] = (z[f"v{v}_ui"], int(z[f"v{v}_ui"].size))
            if f"v{v}_visible0" in z.files:
                res["visible"] = [z[f"v{v}_visible{k}"] for k in range(len(self.meta))]
            self.frames.append(res)

    def draw_ready_shadow(self):
        """isDrawReady(shadowPass >= 0) per pool (7th meta column; older files: same as the main-pass readiness)."""
        return [bool(m[6]) if len(m) > 6 else bool(m[4]) for m in self.meta]

    def oracle_run(self):
        t = self.transforms
        return OracleRun((t, t.shape[1], t.shape[0]), [(p, int(m[1]), int(m[2])) for p, m in zip(self.pools, self.meta)],
                         self.render_types, self.views, self.camera_pos, ready=self.ready,
                         draw_ready=[bool(m[4]) for m in self.meta], counts=[int(m[3]) for m in self.meta],
                         draw_ready_shadow=self.draw_ready_shadow())

    def stage(self, sp):
        t = self.transforms
        sp.set_transforms(t, t.shape[1], t.shape[0])
        sp.set_pool_count(len(self.pools))
        for k, (p, m) in enumerate(zip(self.pools, self.meta)):
            # readiness per view, the way the shim asks isDrawReady(view.shadowPass)
            main_ready, shadow_ready = bool(m[4]), self.draw_ready_shadow()[k]
            sp.set_mesh_pool(k, int(m[0]), p, int(m[1]), int(m[2]), int(m[3]), main_ready or shadow_ready, self.ready[k])
            mask = 0
            for v in range(self.views.size):
                if (main_ready if int(self.views[v]["shadowPass"]) < 0 else shadow_ready):
                    mask |= 1 << v
            sp.set_pool_view_mask(k, mask)
        sp.set_views(self.views, self.camera_pos)


def golden_cases():
    from pathlib import Path
    return sorted((Path(__file__).resolve().parent / "golden").glob("*.npz"))
