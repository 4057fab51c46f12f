#!/usr/bin/env python
"""Runs a few frames of a workload through the C ABI (for ncu captures; prints nothing timing-related)."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from garden_b200 import scenes  # noqa: E402
from garden_b200.binding import ScenePrep  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C4")
ap.add_argument("--entities", type=int, default=4_000_000)
ap.add_argument("--frames", type=int, default=3)
args = ap.parse_args()
cfg = bench.WORKLOADS[args.workload][0]
scene = scenes.config_scene(cfg, n=args.entities)
scene.camera_pos = bench.camera_pos()
views = bench.frame_views(args.workload)
t, pools = scenes.build_aos(scene)
sp = ScenePrep(0)
sp.set_transforms(t, t.dtype.itemsize, t.size)
sp.set_pool_count(len(pools))
for k, m in enumerate(pools):
    sp.set_mesh_pool(k, scene.pools[k].render_type, m, m.dtype.itemsize, m.size)
sp.set_views(views, scene.camera_pos)
sp.set_profiling(True)
phase = np.zeros(6)
for _ in range(args.frames):
    sp.run()
    phase = sp.phase_times()
print("visible", sp.last_visible_total(), "launches", sp.last_launch_count())
print("phase ms (last frame): link %.4f | kCull %.4f | scan+scatter %.4f | prepass %.4f | sort %.4f | emit %.4f | sum %.4f" % (*phase, phase.sum()))
