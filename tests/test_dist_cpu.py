"""CPU tests of the multi-GPU host logic: gather layout, merge semantics, and a world_size-2 gloo exchange."""
import os
import socket

import numpy as np
import pytest

from garden_b200.dist import exchange_counts, merge_reference, plan_gather


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
