"""Multi-GPU plumbing for the sharded scene preparation (one process per GPU, torch.distributed).

Sharding: every rank owns a contiguous entity range (all views), culls and sorts it with the single-GPU path, then
  1. all-gather of the per-list lengths              (tiny; tells every rank the layout of the gathered runs)
  2. all-gather of the sorted (key, payload) runs    (NCCL over NVLink; 8 bytes per visible entity and view)
  3. k-way merge on the device: each rank merges ITS key range of every list (gsp_merge_gathered)
The merged draw order equals a single sort over all entities with ties broken by global entity order, because ranks hold
contiguous ranges in rank order and the merge breaks ties by (rank, payload).

`plan_gather` and `merge_reference` are pure numpy (they define the layout / the expected result) and are what the CPU
tests check, including a world_size-2 gloo run of `exchange_counts`.
"""
from __future__ import annotations

import numpy as np


def plan_gather(counts: np.ndarray):
    """counts: [ranks, lists] list lengths. Returns (offsets [ranks, lists], rank_stride, out_offsets [lists], totals [lists]).

    Every rank packs its lists back to back; rank r's block starts at r * rank_stride in the gathered buffer;
    the merged list l is at out_offsets[l] (sized for the whole list, a rank only fills its slice)."""
    counts = np.asarray(counts, dtype=np.int64)
    offsets = np.zeros_like(counts)
    offsets[:, 1:] = np.cumsum(counts, axis=1)[:, :-1]
    rank_stride = int(max(1, counts.sum(axis=1).max()))
    totals = counts.sum(axis=0)
    out_offsets = np.zeros(counts.shape[1], dtype=np.int64)
    out_offsets[1:] = np.cumsum(totals)[:-1]
    return offsets.astype(np.uint32), rank_stride, out_offsets.astype(np.uint32), totals.astype(np.uint32)


def merge_reference(runs_keys, runs_payloads, my_rank: int | None = None):
    """numpy statement of the merge for ONE list: runs_* are per-rank arrays (sorted by key, ties by payload).
    Returns (keys, payloads, ranks) of the full merged list, or of rank `my_rank`'s key-range slice plus its start."""
    ranks = len(runs_keys)
    keys = np.concatenate(runs_keys) if ranks else np.zeros(0, np.uint32)
    pays = np.concatenate(runs_payloads) if ranks else np.zeros(0, np.uint32)
    src = np.concatenate([np.full(len(k), r, np.uint8) for r, k in enumerate(runs_keys)]) if ranks else np.zeros(0, np.uint8)
    order = np.lexsort((pays, src, keys))  # key, then rank, then payload
    keys, pays, src = keys[order], pays[order], src[order]
    if my_rank is None:
        return keys, pays, src
    run0 = runs_keys[int(np.argmax([len(k) for k in runs_keys]))]  # longest run, lowest rank on ties
    n0 = len(run0)
    lo_key = run0[my_rank * n0 // ranks] if (my_rank > 0 and n0) else None
    hi_key = run0[(my_rank + 1) * n0 // ranks] if (my_rank + 1 < ranks and n0) else None
    lo = int(np.searchsorted(keys, lo_key, side="left")) if lo_key is not None else 0
    hi = int(np.searchsorted(keys, hi_key, side="left")) if hi_key is not None else len(keys)
    hi = max(hi, lo)
    return keys[lo:hi], pays[lo:hi], src[lo:hi], lo


def exchange_counts(counts: np.ndarray, group=None) -> np.ndarray:
    """All-gather of the per-list lengths. Works with any backend (tensors live where the backend needs them)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.as_tensor(np.asarray(counts, dtype=np.int64), device=device)
    out = torch.empty(world * mine.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy().reshape(world, mine.numel())


class RunMerger:
    """Gather + merge of one ScenePrep's sorted runs across the ranks of the default process group (NCCL)."""

    launches_per_frame = 2  # kMergeBounds + kMergeSlice

    def __init__(self, sp):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.sp = sp
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.send_keys = self.send_pays = None
        self.gk = self.gp = None
        self.out_keys = self.out_pays = self.out_ranks = None
        self.bounds = self.slice_info = None
        self.last = None

    def _ensure(self, name: str, n: int, dtype):
        t = getattr(self, name)
        if t is None or t.numel() < n:
            t = self.torch.empty(max(int(n * 1.25), 16), dtype=dtype, device=self.dev)
            setattr(self, name, t)
        return t

    def gather_and_merge(self):
        """Must be called after sp.run_async(); leaves this rank's merged slices on the device.
        Returns dict(counts, totals, slice_info (device tensor [lists, 2]))."""
        torch, dist, sp = self.torch, self.dist, self.sp
        sp.sync()  # list lengths are needed on the host to size the exchange
        counts = sp.list_counts().astype(np.int64)
        lists = counts.size
        all_counts = exchange_counts(counts)
        offsets, stride, out_offsets, totals = plan_gather(all_counts)
        send_k = self._ensure("send_keys", stride, torch.int32)
        send_p = self._ensure("send_pays", stride, torch.int32)
        sp.export_runs(send_k.data_ptr(), send_p.data_ptr(), stride)
        gk = self._ensure("gk", stride * self.world, torch.int32)
        gp = self._ensure("gp", stride * self.world, torch.int32)
        dist.all_gather_into_tensor(gk[: stride * self.world], send_k[:stride])
        dist.all_gather_into_tensor(gp[: stride * self.world], send_p[:stride])
        total = int(totals.sum())
        out_k = self._ensure("out_keys", total, torch.int32)
        out_p = self._ensure("out_pays", total, torch.int32)
        out_r = self._ensure("out_ranks", total, torch.uint8)
        bounds = self._ensure("bounds", lists * self.world * 2, torch.int32)
        sinfo = self._ensure("slice_info", lists * 2, torch.int32)
        meta = torch.as_tensor(np.concatenate([offsets.reshape(-1), all_counts.astype(np.uint32).reshape(-1),
                                               out_offsets]).astype(np.int64), device=self.dev).to(torch.int32)
        n_rl = self.world * lists
        d_off, d_cnt, d_out_off = meta[:n_rl], meta[n_rl:2 * n_rl], meta[2 * n_rl:]
        stream = torch.cuda.current_stream().cuda_stream
        rc = sp.lib.gsp_merge_gathered(stream, self.world, self.rank, lists, stride, gk.data_ptr(), gp.data_ptr(),
                                       d_off.data_ptr(), d_cnt.data_ptr(), int(all_counts.max()), bounds.data_ptr(),
                                       sinfo.data_ptr(), out_k.data_ptr(), out_p.data_ptr(), out_r.data_ptr(),
                                       d_out_off.data_ptr())
        if rc != 0:
            raise RuntimeError(f"gsp_merge_gathered failed with {rc}")
        self.last = {"counts": all_counts, "totals": totals, "out_offsets": out_offsets, "slice_info": sinfo[: lists * 2],
                     "meta": meta, "bytes_gathered": int(stride * self.world * 8)}
        return self.last

    def slices_to_host(self):
        """Downloads this rank's merged slices: list of (start, keys, payloads, ranks) per list (for tests)."""
        info = self.last["slice_info"].cpu().numpy().astype(np.uint32).reshape(-1, 2)
        out = []
        for l, (start, length) in enumerate(info):
            o = int(self.last["out_offsets"][l])
            sl = slice(o, o + int(length))
            out.append((int(start), self.out_keys[sl].cpu().numpy().view(np.uint32),
                        self.out_pays[sl].cpu().numpy().view(np.uint32), self.out_ranks[sl].cpu().numpy()))
        return out
