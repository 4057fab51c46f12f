"""GPU test of the gather + k-way merge, with the ranks emulated on one device: R shards of one scene are prepared one
after the other, their runs concatenated the way the NCCL all-gather would lay them out, and every rank's merged
key-range slice is checked against (a) the numpy statement of the merge and (b) a single-GPU sort of the whole scene."""
import numpy as np
import pytest

from garden_b200 import scenes, views as V
from garden_b200.dist import merge_reference, plan_gather

pytestmark = pytest.mark.gpu


def _prepare(sp, scene, views):
    t, pools = scenes.build_aos(scene)
    sp.set_transforms(t, t.dtype.itemsize, t.size)
    sp.set_pool_count(len(pools))
    for k, m in enumerate(pools):
        sp.set_mesh_pool(k, scene.pools[k].render_type, m, m.dtype.itemsize, m.size)
    sp.set_views(views, scene.camera_pos)
    sp.run()


def _runs_to_host(sp, torch):
    counts = sp.list_counts().astype(np.int64)
    total = int(counts.sum())
    k = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
    p = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
    sp.export_runs(k.data_ptr(), p.data_ptr(), max(total, 1))
    sp.sync()
    torch.cuda.synchronize()
    return counts, k[:total].cpu().numpy().view(np.uint32), p[:total].cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("ranks", [2, 3])
def test_emulated_ranks_merge_equals_single_sort(sceneprep_lib, ranks):
    import torch
    from garden_b200.binding import ScenePrep
    chains, depth = 3000, 4
    n = chains * (depth + 1)
    box = (-150, -10, -150, 150, 10, 150)
    whole = scenes.make_scene(n, depth, 5, box)
    whole.camera_pos = np.array([2.0, 1.0, -3.0], np.float32)
    views, _ = V.camera_and_cascades(0.3, -0.1, 1.2, 16 / 9, 0.01, 100.0, (0.05, 0.1, 0.25, 1.0))
    sp = ScenePrep(0)
    _prepare(sp, whole, views)
    full_counts, full_k, full_p = _runs_to_host(sp, torch)
    lists = full_counts.size

    # shards: contiguous entity ranges cut at chain boundaries
    per = (chains // ranks) * (depth + 1)
    starts = [r * per for r in range(ranks)] + [n]
    shard_runs = []
    for r in range(ranks):
        a, b = starts[r], starts[r + 1]
        sc = scenes.SceneDesc(whole.position[a:b], whole.rotation[a:b], whole.scale[a:b],
                              np.where(whole.parent[a:b] >= 0, whole.parent[a:b] - a, -1).astype(np.int32), whole.tflags[a:b],
                              [scenes.PoolDesc(whole.pools[0].render_type, np.arange(b - a, dtype=np.uint32), whole.pools[0].aabb[a:b])],
                              None, whole.camera_pos)
        _prepare(sp, sc, views)
        shard_runs.append(_runs_to_host(sp, torch))

    all_counts = np.stack([c for c, _, _ in shard_runs])
    assert np.array_equal(all_counts.sum(axis=0), full_counts)
    offsets, stride, out_offsets, totals = plan_gather(all_counts)
    gk = torch.zeros(stride * ranks, dtype=torch.int32, device="cuda")
    gp = torch.zeros(stride * ranks, dtype=torch.int32, device="cuda")
    for r, (c, k, p) in enumerate(shard_runs):
        gk[r * stride: r * stride + k.size] = torch.from_numpy(k.view(np.int32)).cuda()
        gp[r * stride: r * stride + p.size] = torch.from_numpy(p.view(np.int32)).cuda()
    meta = torch.from_numpy(np.concatenate([offsets.reshape(-1), all_counts.astype(np.uint32).reshape(-1), out_offsets]).astype(np.int32)).cuda()
    nrl = ranks * lists
    total = int(totals.sum())
    merged_k = np.zeros(total, np.uint32); merged_p = np.zeros(total, np.uint32); merged_r = np.zeros(total, np.uint8)
    covered = np.zeros(total, bool)
    for me in range(ranks):
        out_k = torch.zeros(total, dtype=torch.int32, device="cuda"); out_p = torch.zeros_like(out_k)
        out_r = torch.zeros(total, dtype=torch.uint8, device="cuda")
        bounds = torch.zeros(lists * ranks * 2, dtype=torch.int32, device="cuda")
        sinfo = torch.zeros(lists * 2, dtype=torch.int32, device="cuda")
        rc = sp.lib.gsp_merge_gathered(0, ranks, me, lists, stride, gk.data_ptr(), gp.data_ptr(), meta[:nrl].data_ptr(),
                                       meta[nrl:2 * nrl].data_ptr(), int(all_counts.max()), bounds.data_ptr(), sinfo.data_ptr(),
                                       out_k.data_ptr(), out_p.data_ptr(), out_r.data_ptr(), meta[2 * nrl:].data_ptr())
        assert rc == 0
        torch.cuda.synchronize()
        info = sinfo.cpu().numpy().reshape(lists, 2)
        ok, op, orr = out_k.cpu().numpy().view(np.uint32), out_p.cpu().numpy().view(np.uint32), out_r.cpu().numpy()
        for l in range(lists):
            start, length = int(info[l, 0]), int(info[l, 1])
            runs_k = [shard_runs[r][1][offsets[r, l]: offsets[r, l] + all_counts[r, l]] for r in range(ranks)]
            runs_p = [shard_runs[r][2][offsets[r, l]: offsets[r, l] + all_counts[r, l]] for r in range(ranks)]
            ek, ep, er, estart = merge_reference(runs_k, runs_p, my_rank=me)
            assert (start, length) == (estart, ek.size), f"rank {me} list {l}: slice bounds"
            o = int(out_offsets[l])
            assert np.array_equal(ok[o:o + length], ek) and np.array_equal(op[o:o + length], ep) and np.array_equal(orr[o:o + length], er)
            g = slice(o + start, o + start + length)
            assert not covered[g].any()
            covered[g] = True
            merged_k[g], merged_p[g], merged_r[g] = ek, ep, er
    assert covered.all()
    # the merged order equals the single-GPU sort of the whole scene (payload = pool << 28 | slot, slot shifted by the shard start)
    global_p = (merged_p & 0xF0000000) | ((merged_p & 0x0FFFFFFF) + np.array(starts, np.uint32)[merged_r])
    assert np.array_equal(merged_k, full_k)
    assert np.array_equal(global_p, full_p)
