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


# ---- the packed exchange block (garden_b200/csrc/merge.cu: kExportPacked / kMergePlan), stated in numpy -----------------
EX_HEADER_WORDS, EX_HEADER_FIXED, EX_MAGIC = 256, 8, 0x47535031


def pack_block(list_keys, list_payloads, capacity: int) -> np.ndarray:
    """What kExportPacked writes for one rank: header {magic, lists, total, capacity, overflow, 0, 0, 0, count[lists]} |
    keys[capacity] | payloads[capacity]. An overflowing rank (total > capacity) carries zero counts and no elements."""
    lists = len(list_keys)
    assert lists <= EX_HEADER_WORDS - EX_HEADER_FIXED
    block = np.zeros(EX_HEADER_WORDS + 2 * capacity, dtype=np.uint32)
    counts = np.array([len(k) for k in list_keys], dtype=np.int64)
    total = int(counts.sum())
    overflow = total > capacity
    block[0:5] = [EX_MAGIC, lists, total, capacity, 1 if overflow else 0]
    if not overflow:
        block[EX_HEADER_FIXED:EX_HEADER_FIXED + lists] = counts
        if total:
            block[EX_HEADER_WORDS:EX_HEADER_WORDS + total] = np.concatenate(list_keys)
            block[EX_HEADER_WORDS + capacity:EX_HEADER_WORDS + capacity + total] = np.concatenate(list_payloads)
    return block


def plan_from_blocks(gathered: np.ndarray, ranks: int, lists: int, capacity: int, out_capacity: int):
    """What kMergePlan derives from the gathered blocks: (offsets [ranks, lists], counts [ranks, lists], out_offsets [lists],
    flags [8]) with flags = {error bits (1 overflow, 2 bad header, 4 merged lists exceed out_capacity), merged total,
    largest per-rank total, 0...}. With a non-zero error word every count is zero (nothing may be merged)."""
    words = EX_HEADER_WORDS + 2 * capacity
    hdr = np.asarray(gathered, dtype=np.uint32).reshape(ranks, words)[:, :EX_HEADER_WORDS]
    error = 0
    for r in range(ranks):
        if hdr[r, 0] != EX_MAGIC or hdr[r, 1] != lists:
            error |= 2
        if hdr[r, 4]:
            error |= 1
    counts = np.zeros((ranks, lists), dtype=np.int64) if error else hdr[:, EX_HEADER_FIXED:EX_HEADER_FIXED + lists].astype(np.int64)
    offsets = np.zeros_like(counts)
    offsets[:, 1:] = np.cumsum(counts, axis=1)[:, :-1]
    totals = counts.sum(axis=0)
    out_offsets = np.zeros(lists, dtype=np.int64)
    out_offsets[1:] = np.cumsum(totals)[:-1]
    merged = int(totals.sum())
    if merged > out_capacity:
        error |= 4
        counts = np.zeros_like(counts)
    flags = np.zeros(8, dtype=np.int64)
    flags[0], flags[1], flags[2] = error, merged, int(hdr[:, 2].max())
    return offsets, counts, out_offsets, flags


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


class PipelinedRunMerger:
    """Gather + merge without any host synchronisation in the steady state, overlapped with the next frame.

    Per frame, on the compute stream (the ScenePrep's stream, which must be torch's current stream):
        gsp_run_async -> gsp_export_runs_packed (lengths read on the device) -> event
    and on a second (exchange) stream:
        wait(event) -> ONE NCCL all-gather of the fixed-capacity blocks -> gsp_merge_gathered_packed -> flags to pinned memory
    so frame k's exchange runs while frame k+1 is being culled and sorted. Blocks, gathered buffers and merged slices are
    double-buffered; a buffer set is reused only after the exchange that read it has finished (event wait, device side).

    The block capacity is speculative (sized from a measured frame, with head room). The merge reports overflow through
    the plan flags; `poll()` looks at the flags of finished frames and returns the frames that must be repeated after
    `grow()` — a draw list is only valid once its frame's flags have been seen clean (`finish()` checks all of them)."""

    HEAD_ROOM = 1.25
    launches_per_frame = 4  # kExportPacked + kMergePlan + kMergeBounds + kMergeSlice

    def __init__(self, sp, capacity: int | None = None, overlap: bool = True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.sp = torch, dist, sp
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.compute = torch.cuda.current_stream()
        self.exchange = torch.cuda.Stream() if overlap else self.compute
        self.lists = sp.list_count()
        self.capacity = 0
        self.sets = []
        self.frame_index = 0
        self.pending = []  # (frame index, buffer set) whose flags have not been inspected yet
        self.flags_seen = {}
        self.timing = None  # set to a list to collect per-frame events (export | all-gather | merge)
        if capacity is None:
            capacity = self.measure_capacity()
        self._allocate(int(capacity))

    def measure_capacity(self) -> int:
        """One synchronous frame: the largest per-rank total over all ranks, with head room (every rank gets the same value)."""
        torch, dist = self.torch, self.dist
        self.sp.run()
        self.lists = self.sp.list_count()
        mine = torch.tensor([int(self.sp.list_counts().astype(np.int64).sum())], dtype=torch.int64, device=self.dev)
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
        return int(int(mine.item()) * self.HEAD_ROOM) + 4096

    def _allocate(self, capacity: int):
        torch, lib = self.torch, self.sp.lib
        self.capacity = capacity
        words = lib.gsp_exchange_block_words(capacity)
        plan_words = lib.gsp_merge_plan_words(self.world, max(self.lists, 1))
        out_cap = capacity * self.world  # every rank's slice fits even if one rank's key range swallowed everything
        self.sets = []
        for _ in range(2):
            self.sets.append({
                "send": torch.empty(words, dtype=torch.int32, device=self.dev),
                "gathered": torch.empty(words * self.world, dtype=torch.int32, device=self.dev),
                "plan": torch.zeros(plan_words, dtype=torch.int32, device=self.dev),
                "slice_info": torch.zeros(max(self.lists, 1) * 2, dtype=torch.int32, device=self.dev),
                "out_keys": torch.empty(out_cap, dtype=torch.int32, device=self.dev),
                "out_pays": torch.empty(out_cap, dtype=torch.int32, device=self.dev),
                "out_ranks": torch.empty(out_cap, dtype=torch.uint8, device=self.dev),
                "flags": torch.zeros(8, dtype=torch.int32).pin_memory(),
                "exported": torch.cuda.Event(), "done": torch.cuda.Event(), "used": False, "out_cap": out_cap,
            })

    def grow(self, needed: int):
        """Re-sizes the blocks for `needed` elements per rank (collective: every rank must call it with the same value)."""
        self.finish(check=False)
        self._allocate(int(needed * self.HEAD_ROOM) + 4096)

    def frame(self):
        """Enqueues one frame: scene preparation, export, all-gather, merge. Returns the buffer set that will hold the result."""
        torch, dist, sp = self.torch, self.dist, self.sp
        s = self.sets[self.frame_index & 1]
        if s["used"]:
            self.compute.wait_event(s["done"])  # the exchange that last read this set's block has finished
        timing = self.timing
        sp.run_async()
        if timing is not None:
            t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            t[0].record(self.compute)
        sp.export_runs_packed(s["send"].data_ptr(), self.capacity)
        s["exported"].record(self.compute)
        with torch.cuda.stream(self.exchange):
            self.exchange.wait_event(s["exported"])
            if timing is not None:
                t[1].record(self.exchange)
            dist.all_gather_into_tensor(s["gathered"], s["send"])
            if timing is not None:
                t[2].record(self.exchange)
            rc = sp.lib.gsp_merge_gathered_packed(self.exchange.cuda_stream, self.world, self.rank, self.lists, self.capacity,
                                                  s["gathered"].data_ptr(), s["plan"].data_ptr(), s["slice_info"].data_ptr(),
                                                  s["out_keys"].data_ptr(), s["out_pays"].data_ptr(),
                                                  s["out_ranks"].data_ptr(), s["out_cap"])
            if rc != 0:
                raise RuntimeError(f"gsp_merge_gathered_packed failed with {rc}")
            if timing is not None:
                t[3].record(self.exchange)
                timing.append(t)
            s["flags"].copy_(s["plan"][-8:], non_blocking=True)
            s["done"].record(self.exchange)
        s["used"] = True
        s["frame"] = self.frame_index
        self.pending.append((self.frame_index, s))
        self.frame_index += 1
        return s

    def poll(self, wait: bool = False):
        """Inspects the flags of frames whose exchange has finished. Returns [(frame, error bits, needed capacity)] of the
        frames that failed (empty when everything seen so far is clean)."""
        failed, still = [], []
        for idx, s in self.pending:
            if s.get("frame") != idx:  # its buffers were reused: the flags were already overwritten by a later frame
                continue
            if wait:
                s["done"].synchronize()
            if not s["done"].query():
                still.append((idx, s))
                continue
            f = s["flags"].numpy().astype(np.int64) & 0xFFFFFFFF
            self.flags_seen[idx] = f.copy()
            if f[0]:
                failed.append((idx, int(f[0]), int(f[2])))
        self.pending = still
        return failed

    def timing_summary(self):
        """Mean milliseconds of the export, all-gather and merge steps over the frames collected in `self.timing`."""
        if not self.timing:
            return None
        self.torch.cuda.synchronize()
        n = len(self.timing)
        return {"export_ms": sum(t[0].elapsed_time(t[1]) for t in self.timing) / n,
                "allgather_ms": sum(t[1].elapsed_time(t[2]) for t in self.timing) / n,
                "merge_ms": sum(t[2].elapsed_time(t[3]) for t in self.timing) / n, "frames": n}

    def finish(self, check: bool = True):
        """Waits for every enqueued exchange (device side for the compute stream, host side for the flags)."""
        for s in self.sets:
            if s["used"]:
                self.compute.wait_event(s["done"])
        failed = self.poll(wait=True)
        if check and failed:
            raise RuntimeError(f"exchange overflow in frames {failed}: call grow() and repeat them")
        return failed

    def last_result(self):
        """Buffer set of the most recent frame plus its host-side description (for tests): synchronises."""
        s = self.sets[(self.frame_index - 1) & 1]
        s["done"].synchronize()
        lists, world = self.lists, self.world
        plan = s["plan"].cpu().numpy().view(np.uint32)
        n_rl = world * lists
        counts = plan[n_rl:2 * n_rl].reshape(world, lists).astype(np.int64)
        out_offsets = plan[2 * n_rl:2 * n_rl + lists].astype(np.int64)
        info = s["slice_info"].cpu().numpy().view(np.uint32).reshape(-1, 2)
        slices = []
        for l in range(lists):
            o, start, length = int(out_offsets[l]), int(info[l, 0]), int(info[l, 1])
            sl = slice(o, o + length)
            slices.append((start, s["out_keys"][sl].cpu().numpy().view(np.uint32),
                           s["out_pays"][sl].cpu().numpy().view(np.uint32), s["out_ranks"][sl].cpu().numpy()))
        return {"counts": counts, "out_offsets": out_offsets, "flags": plan[-8:].copy(), "slices": slices,
                "bytes_gathered": int(s["gathered"].numel() * 4)}
