"""CPU tests of the multi-GPU host logic: gather layout, merge semantics, and a world_size-2 gloo exchange."""
import os
import socket

import numpy as np
import pytest

from garden_b200.dist import (EX_HEADER_FIXED, EX_HEADER_WORDS, exchange_counts, merge_reference, pack_block, plan_from_blocks,
                              plan_gather)


def _runs(rng, ranks, n_max, key_bits=10):
    keys, pays = [], []
    for r in range(ranks):
        n = int(rng.integers(0, n_max))
        k = np.sort(rng.integers(0, 1 << key_bits, n).astype(np.uint32), kind="stable")
        p = np.zeros(n, np.uint32)
        # ties inside a run are ordered by payload (stable sort of a slot-ordered compaction)
        for key in np.unique(k):
            sel = np.nonzero(k == key)[0]
            p[sel] = np.sort(rng.choice(1 << 20, sel.size, replace=False)).astype(np.uint32)
        keys.append(k); pays.append(p)
    return keys, pays


def test_plan_gather_layout():
    counts = np.array([[3, 0, 5], [1, 7, 2], [0, 0, 0]])
    offsets, stride, out_offsets, totals = plan_gather(counts)
    assert stride == 10
    assert offsets.tolist() == [[0, 3, 3], [0, 1, 8], [0, 0, 0]]
    assert totals.tolist() == [4, 7, 7]
    assert out_offsets.tolist() == [0, 4, 11]


@pytest.mark.parametrize("ranks", [1, 2, 3, 8])
def test_key_range_slices_tile_the_full_merge(ranks):
    rng = np.random.default_rng(ranks)
    for trial in range(20):
        keys, pays = _runs(rng, ranks, 400, key_bits=4 if trial % 2 else 12)  # many ties / few ties
        full_k, full_p, full_r = merge_reference(keys, pays)
        assert np.all(np.diff(full_k.astype(np.int64)) >= 0)
        # ties: rank ascending, then payload ascending
        same = np.nonzero(np.diff(full_k.astype(np.int64)) == 0)[0]
        assert np.all((full_r[same] < full_r[same + 1]) | ((full_r[same] == full_r[same + 1]) & (full_p[same] < full_p[same + 1])))
        pos = 0
        for r in range(ranks):
            k, p, s, start = merge_reference(keys, pays, my_rank=r)
            assert start == pos
            assert np.array_equal(k, full_k[pos:pos + k.size]) and np.array_equal(p, full_p[pos:pos + k.size])
            assert np.array_equal(s, full_r[pos:pos + k.size])
            pos += k.size
        assert pos == full_k.size


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = np.array([10 * (rank + 1), rank, 7], dtype=np.int64)
    allc = exchange_counts(counts)
    offsets, stride, out_offsets, totals = plan_gather(allc)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([allc.reshape(-1), offsets.reshape(-1), [stride], totals]))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_counts_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b)  # every rank derives the same layout
    assert a[:6].tolist() == [10, 0, 7, 20, 1, 7]
    assert a[12] == 28  # rank stride = longest packed block
    assert a[13:].tolist() == [30, 1, 14]


def _lists_for_rank(rank, lists=3):
    rng = np.random.default_rng(100 + rank)
    keys, pays = _runs(rng, lists, 300, key_bits=10)  # one sorted run per LIST of this rank
    return keys, pays


def _gloo_block_worker(rank, world, port, out_dir, capacity):
    """The packed exchange over gloo: every rank packs its lists into one fixed-capacity block, ONE all-gather of equal
    blocks, every rank plans the merge from the gathered headers and merges its key range of every list."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys, pays = _lists_for_rank(rank)
    block = pack_block(keys, pays, capacity)
    mine = torch.from_numpy(block.view(np.int32))
    gathered = torch.empty(world * mine.numel(), dtype=torch.int32)
    dist.all_gather_into_tensor(gathered, mine)
    g = gathered.numpy().view(np.uint32)
    lists = len(keys)
    offsets, counts, out_offsets, flags = plan_from_blocks(g, world, lists, capacity, out_capacity=capacity * world)
    words = EX_HEADER_WORDS + 2 * capacity
    out = {"flags": flags, "counts": counts, "out_offsets": out_offsets}
    if flags[0] == 0:
        for l in range(lists):
            rk = [g[r * words + EX_HEADER_WORDS + offsets[r, l]: r * words + EX_HEADER_WORDS + offsets[r, l] + counts[r, l]] for r in range(world)]
            rp = [g[r * words + EX_HEADER_WORDS + capacity + offsets[r, l]: r * words + EX_HEADER_WORDS + capacity + offsets[r, l] + counts[r, l]]
                  for r in range(world)]
            k, p, s, start = merge_reference(rk, rp, my_rank=rank)
            out[f"k{l}"], out[f"p{l}"], out[f"s{l}"], out[f"start{l}"] = k, p, s, np.array([start])
    np.savez(os.path.join(out_dir, f"b{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_packed_block_exchange_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world, lists = 2, 3
    per_rank = [_lists_for_rank(r, lists) for r in range(world)]
    need = max(sum(len(k) for k in keys) for keys, _ in per_rank)
    mp.spawn(_gloo_block_worker, args=(world, _free_port(), str(tmp_path), need + 7), nprocs=world, join=True)
    res = [np.load(tmp_path / f"b{r}.npz") for r in range(world)]
    for r in range(world):
        assert res[r]["flags"][0] == 0 and res[r]["flags"][2] == need
        assert np.array_equal(res[r]["counts"], [[len(k) for k in keys] for keys, _ in per_rank])
    for l in range(lists):
        full_k, full_p, full_r = merge_reference([per_rank[r][0][l] for r in range(world)], [per_rank[r][1][l] for r in range(world)])
        pos = 0
        for r in range(world):  # the ranks' slices tile the full merge of the list, in rank order
            assert int(res[r][f"start{l}"][0]) == pos
            n = res[r][f"k{l}"].size
            assert np.array_equal(res[r][f"k{l}"], full_k[pos:pos + n]) and np.array_equal(res[r][f"p{l}"], full_p[pos:pos + n])
            assert np.array_equal(res[r][f"s{l}"], full_r[pos:pos + n])
            pos += n
        assert pos == full_k.size
    # overflow protocol: a capacity one rank cannot meet is flagged by every rank, nothing is planned, the need is reported
    mp.spawn(_gloo_block_worker, args=(world, _free_port(), str(tmp_path), need - 1), nprocs=world, join=True)
    for r in range(world):
        z = np.load(tmp_path / f"b{r}.npz")
        assert z["flags"][0] & 1 and z["flags"][2] == need and not z["counts"].any()


def test_plan_from_blocks_flags():
    cap = 16
    a = pack_block([np.arange(3, dtype=np.uint32), np.arange(2, dtype=np.uint32)], [np.arange(3, dtype=np.uint32), np.arange(2, dtype=np.uint32)], cap)
    b = pack_block([np.arange(5, dtype=np.uint32), np.zeros(0, np.uint32)], [np.arange(5, dtype=np.uint32), np.zeros(0, np.uint32)], cap)
    g = np.concatenate([a, b])
    offsets, counts, out_offsets, flags = plan_from_blocks(g, 2, 2, cap, out_capacity=32)
    assert counts.tolist() == [[3, 2], [5, 0]] and offsets.tolist() == [[0, 3], [0, 5]] and out_offsets.tolist() == [0, 8]
    assert flags[:3].tolist() == [0, 10, 5]
    assert plan_from_blocks(g, 2, 2, cap, out_capacity=9)[3][0] == 4 and not plan_from_blocks(g, 2, 2, cap, out_capacity=9)[1].any()
    bad = g.copy(); bad[0] = 1
    assert plan_from_blocks(bad, 2, 2, cap, out_capacity=32)[3][0] == 2
    assert plan_from_blocks(g, 2, 3, cap, out_capacity=32)[3][0] == 2  # list count mismatch


# ---- the all-to-all protocol: common splitters from samples, one sub-block per destination ---------------------------------
def test_common_splitters_balance_and_agree():
    from garden_b200.dist import EX_SAMPLES, common_splitters, sample_run, split_bounds
    rng = np.random.default_rng(5)
    for ranks in (1, 2, 3, 8):
        for trial in range(6):
            runs = [np.sort(rng.integers(0, 1 << (8 if trial % 2 else 30), int(rng.integers(0, 20000))).astype(np.uint32)) for _ in range(ranks)]
            if trial == 5:
                runs[0] = np.zeros(0, np.uint32)  # an empty run carries no weight
            samples = np.stack([sample_run(k) for k in runs])
            assert samples.shape == (ranks, EX_SAMPLES + 1) and [int(s[-1]) for s in samples] == [len(k) for k in runs]
            sp = common_splitters(samples)
            assert sp.size == ranks - 1 and np.all(np.diff(sp.astype(np.int64)) >= 0)
            bounds = [split_bounds(k, sp) for k in runs]
            for b, k in zip(bounds, runs):
                assert b[0] == 0 and b[-1] == len(k) and np.all(np.diff(b) >= 0)
            # every key lands in exactly one destination, destinations are key ranges, and (few ties) they are balanced
            total = sum(len(k) for k in runs)
            sizes = [sum(int(b[d + 1] - b[d]) for b in bounds) for d in range(ranks)]
            assert sum(sizes) == total
            for d in range(ranks - 1):
                left = [k[b[d]:b[d + 1]] for k, b in zip(runs, bounds)]
                right = [k[b[d + 1]:b[d + 2]] for k, b in zip(runs, bounds)]
                lmax = max((int(x.max()) for x in left if x.size), default=-1)
                rmin = min((int(x.min()) for x in right if x.size), default=1 << 40)
                assert lmax < rmin
            if trial % 2 == 0 and total > 4000:
                assert max(sizes) <= total / ranks * 1.25 + 64, (ranks, sizes)


def _gloo_alltoall_worker(rank, world, port, out_dir):
    """The all-to-all exchange over gloo with the numpy statements: samples all-gathered, common splitters, every rank cuts
    its runs and hands destination d its pieces, every rank merges what it received; the slices tile the full merge."""
    import torch
    import torch.distributed as dist
    from garden_b200.dist import common_splitters, sample_run, split_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    keys, pays = _lists_for_rank(rank)
    lists = len(keys)
    mine = torch.from_numpy(np.stack([sample_run(k) for k in keys]).astype(np.int64))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    samples = np.stack([g.numpy() for g in gathered]).astype(np.uint32)  # [ranks, lists, S + 1]
    out = {}
    pieces = [[None] * lists for _ in range(world)]
    for l in range(lists):
        sp = common_splitters(samples[:, l, :])
        b = split_bounds(keys[l], sp)
        for d in range(world):
            # what a sub-block carries: keys, payloads and "elements of my run below your key range" (kPackByDestination's header)
            pieces[d][l] = (keys[l][b[d]:b[d + 1]], pays[l][b[d]:b[d + 1]], int(b[d]))
    received = [None] * world
    # (gloo has no all_to_all: the object channel stands in for ncclSend / ncclRecv)
    everyone = [None] * world
    dist.all_gather_object(everyone, pieces)
    for src in range(world):
        received[src] = everyone[src][rank]
    start = np.zeros(lists, np.int64)
    for l in range(lists):
        k, p, s = merge_reference([received[src][l][0] for src in range(world)], [received[src][l][1] for src in range(world)])
        out[f"k{l}"], out[f"p{l}"], out[f"s{l}"] = k, p, s
        start[l] = sum(received[src][l][2] for src in range(world))  # kSliceStarts: no further collective
    out["start"] = start
    np.savez(os.path.join(out_dir, f"a{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


def test_alltoall_exchange_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world, lists = 2, 3
    per_rank = [_lists_for_rank(r, lists) for r in range(world)]
    mp.spawn(_gloo_alltoall_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"a{r}.npz") for r in range(world)]
    for l in range(lists):
        full_k, full_p, full_r = merge_reference([per_rank[r][0][l] for r in range(world)], [per_rank[r][1][l] for r in range(world)])
        pos = 0
        for r in range(world):
            assert int(res[r]["start"][l]) == pos
            n = res[r][f"k{l}"].size
            assert np.array_equal(res[r][f"k{l}"], full_k[pos:pos + n]) and np.array_equal(res[r][f"p{l}"], full_p[pos:pos + n])
            assert np.array_equal(res[r][f"s{l}"], full_r[pos:pos + n])
            pos += n
        assert pos == full_k.size


@pytest.mark.parametrize("ranks,key_bits", [(1, 8), (2, 32), (3, 4), (5, 6), (8, 10), (8, 2)])
def test_merge_tree_statement_equals_the_full_merge(ranks, key_bits):
    """The pairwise merge-path tree (what kTreePartition + kMergeTree compute, stated in numpy with the same tile cuts and
    the same tie rule) equals the (key, rank, payload) sort of all runs, for runs with many ties and of very different
    lengths; every diagonal cut is consistent with the merged order."""
    from garden_b200.dist import merge_path_cut, merge_reference, merge_tree_reference
    rng = np.random.default_rng(17 * ranks + key_bits)
    lengths = rng.integers(0, 9000, size=ranks)
    lengths[rng.integers(0, ranks)] = 0
    runs_k, runs_p = [], []
    for n in lengths:
        k = np.sort(rng.integers(0, 1 << key_bits, size=int(n), dtype=np.uint64).astype(np.uint32))
        p = rng.integers(0, 1 << 28, size=int(n), dtype=np.uint64).astype(np.uint32)
        p = p[np.lexsort((p, k))]
        runs_k.append(k); runs_p.append(p)
    ek, ep, er = merge_reference(runs_k, runs_p)
    for tile in (2048, 7):
        tk, tp, tr = merge_tree_reference(runs_k, runs_p, tile=tile)
        assert np.array_equal(tk, ek) and np.array_equal(tp, ep) and np.array_equal(tr, er)
    if ranks >= 2:
        a, b = runs_k[0], runs_k[1]
        mk, _, mr = merge_reference([a, b], [runs_p[0], runs_p[1]])
        for diag in rng.integers(0, len(a) + len(b) + 1, size=20):
            assert merge_path_cut(a, b, int(diag)) == int((mr[:diag] == 0).sum())
