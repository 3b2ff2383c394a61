"""Sharded string index: one doc-range shard per rank, one process per GPU (BASELINE config 4).

The reference is a single process and has nothing like this; the contract comes from its result order.  A document
lives in exactly one shard and ``string_index::query`` reports rows in ascending doc index (src/index.cpp:316-322),
so when rank g indexes the documents ``[g*nd/W, (g+1)*nd/W)`` the full row of a pattern is the concatenation of the
shard rows in rank order — per-document counts never need summing.  What the ranks exchange per batch is small:

* the packed pattern batch, broadcast from the rank that received the request,
* per-pattern (row length, occurrences) of every shard (4-byte integers when every shard has < 2^31 suffixes), one
  ``all_gather`` -> global CSR offsets by a scan in rank order, and the occurrence totals the reference's ``count`` needs as column sums,
* optionally the rows themselves, gathered to one rank (``gather_rows``), for callers that need the flat answer.

Plumbing is ``torch.distributed`` (NCCL over NVLink on GPUs; the CPU tests run the same code over gloo with a fake
local engine).  All device compute stays in the C-ABI library; this file moves no corpus bytes between ranks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import StringIndex, pack


def shard_range(nd: int, rank: int, world: int) -> tuple[int, int]:
    """Documents [lo, hi) of shard `rank`: contiguous, ascending with rank, sizes differ by at most one."""
    return rank * nd // world, (rank + 1) * nd // world


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory (torch wraps it without a copy)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedResult:
    """Result of one sharded batch on this rank.

    row_off, pairs   this shard's CSR rows (torch tensors on the compute device; pairs is [total, 2])
    shard_rows       int64 [world, npat]   row length of every pattern on every shard
    global_row_off   int64 [npat + 1]      CSR offsets of the concatenated (rank-ordered) answer
    rank_base        int64 [npat]          where this shard's part of row q starts inside global row q
    occurrences      int64 [npat]          total occurrences of every pattern over all shards
    """

    def __init__(self, row_off, pairs, shard_rows, occurrences, rank, keep=None, exchanged=None):
        self.row_off, self.pairs = row_off, pairs
        self._shard_rows, self._occ = shard_rows, occurrences
        self._exchanged = exchanged  # [world, 2, npat] as it came out of the all_gather (widened on first use)
        self.rank = rank
        self._gro = None
        self._keep = keep  # owner of the device buffers row_off / pairs alias

    @property
    def shard_rows(self):
        if self._shard_rows is None:
            if self._exchanged is not None:
                self._shard_rows = self._exchanged[:, 0, :].to(torch.int64)
            else:  # one shard: nothing was exchanged
                self._shard_rows = (self.row_off[1:] - self.row_off[:-1]).view(1, -1)
        return self._shard_rows

    @property
    def occurrences(self):
        """Total occurrences of every pattern over all shards = column sums of the exchanged counts."""
        if self._occ is None:
            if self._exchanged is not None:
                self._occ = self._exchanged[:, 1, :].sum(dim=0, dtype=torch.int64)
            else:
                self._occ = self._local_occ()
        return self._occ

    @property
    def global_row_off(self):
        """Scan of the per-shard row lengths in rank order (computed on first use)."""
        if self._gro is None:
            tot = self.shard_rows.sum(dim=0)
            self._gro = torch.zeros(tot.numel() + 1, dtype=torch.int64, device=tot.device)
            torch.cumsum(tot, 0, out=self._gro[1:])
        return self._gro

    @property
    def rank_base(self):
        return self.shard_rows[: self.rank].sum(dim=0)

    @property
    def npat(self) -> int:
        return self.row_off.numel() - 1


class ShardedStringIndex:
    """One shard of a doc-range-sharded ``string_index``.  Every rank of `group` constructs one, adds ITS documents
    (in global doc order) and calls ``build()`` / ``locate_batch()`` collectively."""

    def __init__(self, group=None, device: torch.device | None = None, index_factory=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = device
        make = index_factory or (lambda: StringIndex(device=device.index if device.type == "cuda" else -1))
        self.local = make()
        # collectives and p2p calls take GLOBAL ranks; `src` / `dst` arguments of this class are ranks inside `group`
        self._global = (lambda r: dist.get_global_rank(group, r)) if (dist.is_initialized() and group is not None) else (lambda r: r)
        self.nd_local = 0
        self.nd_global = None
        self.doc_base = None
        self.narrow = False
        # CDB_SHARD_TRACE=1: (broadcast, local locate, all_gather) milliseconds of every batch, CUDA events (profiling aid)
        import os
        self._trace = [] if os.environ.get("CDB_SHARD_TRACE") else None
        self._side = None  # stream of the per-batch exchange

    # -- corpus --------------------------------------------------------------------------------------------------
    def add_many(self, ids, text, doc_off):
        """This rank's documents (string_index::add, src/index.cpp:174-177, for every document of the shard)."""
        self.local.add_many(ids, text, doc_off)
        self.nd_local += len(ids)

    def build_device(self, d_text_ptr, d_doc_off_ptr, d_ids_ptr, nd, stream=0, keep=()):
        self.local.build_device(d_text_ptr, d_doc_off_ptr, d_ids_ptr, nd, stream, keep=keep)
        self.nd_local = nd
        self._exchange_sizes()

    def build(self):
        """Every shard builds its own suffix array; no communication (a suffix never leaves its document)."""
        self.local.build()
        self._exchange_sizes()

    def _exchange_sizes(self):
        info = getattr(self.local, "info", None)
        n_local = int(info()["n"]) if info is not None else 1 << 62  # suffixes of this shard (unknown: assume large)
        sizes = torch.zeros(2 * self.world, dtype=torch.int64, device=self.device)
        mine = torch.tensor([self.nd_local, n_local], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_gather_into_tensor(sizes, mine, group=self.group)
        else:
            sizes.copy_(mine)
        sizes = sizes.cpu().view(self.world, 2)
        self.nd_global = int(sizes[:, 0].sum())
        self.doc_base = int(sizes[: self.rank, 0].sum())
        # per-pattern row lengths and occurrence counts of a shard are bounded by its suffix count: when every shard
        # has fewer than 2^31 suffixes the per-batch exchange carries 4-byte instead of 8-byte integers
        self.narrow = bool(int(sizes[:, 1].max()) < (1 << 31))

    # -- query ---------------------------------------------------------------------------------------------------
    def broadcast_patterns(self, patterns=None, pat_off=None, src: int = 0):
        """The rank that holds the request broadcasts the packed batch -> (pat uint8 tensor, pat_off int64 tensor)
        on the compute device of every rank."""
        hdr = torch.zeros(2, dtype=torch.int64, device=self.device)
        if self.rank == src:
            if pat_off is None:
                pat, pat_off = pack(list(patterns))
            else:
                pat, pat_off = np.ascontiguousarray(patterns, np.uint8), np.ascontiguousarray(pat_off, np.int64)
            hdr[0], hdr[1] = len(pat), len(pat_off) - 1
        if self.world > 1:
            dist.broadcast(hdr, self._global(src), group=self.group)
        nbytes, npat = int(hdr[0]), int(hdr[1])
        d_pat = torch.zeros(nbytes + 8, dtype=torch.uint8, device=self.device)
        d_off = torch.zeros(npat + 1, dtype=torch.int64, device=self.device)
        if self.rank == src:
            d_pat[:nbytes].copy_(torch.from_numpy(pat))
            d_off.copy_(torch.from_numpy(pat_off))
        if self.world > 1:
            dist.broadcast(d_pat, self._global(src), group=self.group)
            self._broadcast_offsets(d_off, d_pat.numel(), src)
        return d_pat, d_off

    def _broadcast_offsets(self, d_off, nbytes: int, src: int):
        """Pattern offsets travel as 4-byte integers whenever the packed batch is shorter than 2 GB."""
        if nbytes >= (1 << 31):
            dist.broadcast(d_off, self._global(src), group=self.group)
            return
        off32 = d_off.to(torch.int32) if self.rank == src else torch.empty(d_off.numel(), dtype=torch.int32, device=self.device)
        dist.broadcast(off32, self._global(src), group=self.group)
        if self.rank != src:
            d_off.copy_(off32)

    def locate_local(self, d_pat, d_off, rows_ready=None):
        """This shard's rows for the batch -> (row_off, pairs[total,2], stats32[2*npat], keep).  rows_ready(stats32 tensor)
        is called as soon as the per-pattern counts are enqueued, before the pairs are filled (the exchange starts there)."""
        npat = d_off.numel() - 1
        if self.device.type == "cuda":
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if rows_ready is not None and npat:
                dev = self.device
                res = self.local.locate_batch_device_ex(
                    d_pat.data_ptr(), d_off.data_ptr(), npat, stream,
                    lambda ptr, n, st: rows_ready(torch.as_tensor(_DevArray(ptr, 2 * n, "<i4"), device=dev), st))
            else:
                res = self.local.locate_batch_device(d_pat.data_ptr(), d_off.data_ptr(), npat, stream)
            row_off = torch.as_tensor(_DevArray(res.row_off, npat + 1), device=self.device)
            tp = res.total_pairs
            pairs = (torch.as_tensor(_DevArray(res.pairs, 2 * tp), device=self.device).view(tp, 2) if tp
                     else torch.zeros((0, 2), dtype=torch.int64, device=self.device))
            # (row length, occurrences) of every pattern as 32-bit integers, written by the library: the exchange payload
            stats = torch.as_tensor(_DevArray(res.stats32, 2 * npat, "<i4"), device=self.device) if npat and res.stats32 else None
            return row_off, pairs, stats, _ResultOwner(self.local, res)
        # host engine (CPU tests inject one; the product StringIndex raises without a CUDA device)
        nbytes = int(d_off[-1]) if npat else 0
        ro, pr = self.local.locate_batch(d_pat[:nbytes].numpy(), d_off.numpy())
        pr = np.asarray(pr, np.int64).reshape(-1, 2)
        ro_t = torch.from_numpy(np.asarray(ro, np.int64))
        occ = torch.zeros(npat, dtype=torch.int64)
        if len(pr):
            occ.index_add_(0, torch.repeat_interleave(torch.arange(npat), ro_t[1:] - ro_t[:-1]), torch.from_numpy(pr[:, 1].copy()))
        rows = ro_t[1:] - ro_t[:-1]
        return ro_t, torch.from_numpy(pr.copy()), torch.stack([rows, occ]).reshape(-1), None

    def locate_batch(self, patterns=None, pat_off=None, src: int = 0, device_patterns=None) -> ShardedResult:
        """Collective.  `patterns` (list of bytes, or packed uint8 + offsets) is read on rank `src` only;
        `device_patterns=(d_pat, d_off)` skips the packing when the batch is already on the device of rank `src`."""
        trace = self._trace
        if trace is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        if device_patterns is not None:
            d_pat, d_off = device_patterns
            if self.world > 1:
                d_pat, d_off = self._broadcast_device_patterns(d_pat, d_off, src)
        else:
            d_pat, d_off = self.broadcast_patterns(patterns, pat_off, src)
        npat = d_off.numel() - 1
        if trace is not None:
            ev[1].record()
        # Narrow exchange on GPUs: the all_gather of the 32-bit counts starts on a side stream from inside the locate, as
        # soon as the counts are enqueued, and runs under the kernel that fills the pairs.
        early = {}
        hook = None
        if self.world > 1 and self.narrow and self.device.type == "cuda" and npat:
            def hook(stats_t, ready_stream=None):
                # the stream the counts are ready on: the launching stream, or the library's side stream
                main = torch.cuda.ExternalStream(ready_stream, device=self.device) if ready_stream else torch.cuda.current_stream(self.device)
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.device)
                ready = torch.cuda.Event()
                ready.record(main)
                self._side.wait_event(ready)
                with torch.cuda.stream(self._side):
                    out_t = torch.empty(self.world * 2 * npat, dtype=torch.int32, device=self.device)
                    dist.all_gather_into_tensor(out_t, stats_t, group=self.group)
                early["allst"] = out_t
        row_off, pairs, stats, keep = self.locate_local(d_pat, d_off, hook)
        if trace is not None:
            ev[2].record()
        if self.world == 1:  # nothing to exchange: per-pattern counts are derived from this shard's result on first use
            res = ShardedResult(row_off, pairs, None, None, 0, keep)
            dev, local_stats = self.device, stats
            if keep is not None:
                res._local_occ = lambda: (torch.as_tensor(_DevArray(keep.res.right, npat), device=dev)
                                          - torch.as_tensor(_DevArray(keep.res.left, npat), device=dev)) if npat else d_off[:0]
            else:
                res._local_occ = lambda: local_stats[npat:].to(torch.int64)
            return res
        if stats is None:  # empty batch
            stats = torch.zeros(2 * npat, dtype=torch.int32, device=self.device)
        # ONE collective per batch: every shard's (row lengths, occurrences) per pattern, as the 32-bit integers the
        # library wrote when every shard has < 2^31 suffixes (no element-wise work between the locate and the collective).
        # The occurrence totals the reference's `count` needs are the column sums, so no separate all_reduce; shard_rows,
        # occurrences and the global CSR offsets are derived from the gathered block on first use.
        wide = not self.narrow
        if wide and stats.dtype != torch.int64:  # a shard with >= 2^31 suffixes: counts may not fit 32 bits
            rows = row_off[1:] - row_off[:-1]
            occ = (torch.as_tensor(_DevArray(keep.res.right, npat), device=self.device)
                   - torch.as_tensor(_DevArray(keep.res.left, npat), device=self.device)) if keep is not None else stats[npat:].to(torch.int64)
            stats = torch.stack([rows, occ]).reshape(-1)
        if "allst" in early:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            allst = early["allst"]
            allst.record_stream(torch.cuda.current_stream(self.device))
        else:
            allst = torch.empty(self.world * 2 * npat, dtype=stats.dtype, device=self.device)
            dist.all_gather_into_tensor(allst, stats.contiguous(), group=self.group)
        allst = allst.view(self.world, 2, npat)
        if trace is not None:
            ev[3].record()
            ev[3].synchronize()
            trace.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
        return ShardedResult(row_off, pairs, None, None, self.rank, keep, exchanged=allst)

    def _broadcast_device_patterns(self, d_pat, d_off, src: int):
        """One broadcast per batch: [pattern offsets as 4-byte integers | pattern bytes] in one buffer (batches of 2 GB
        or more keep the two-broadcast path with 8-byte offsets)."""
        nbytes, npat = d_pat.numel(), d_off.numel() - 1
        if nbytes >= (1 << 31):
            dist.broadcast(d_pat, self._global(src), group=self.group)
            dist.broadcast(d_off, self._global(src), group=self.group)
            return d_pat, d_off
        head = 4 * (npat + 1)
        buf = torch.empty(head + nbytes, dtype=torch.uint8, device=self.device)
        if self.rank == src:
            buf[:head].view(torch.int32).copy_(d_off)
            buf[head:].copy_(d_pat)
        dist.broadcast(buf, self._global(src), group=self.group)
        if self.rank != src:
            d_off.copy_(buf[:head].view(torch.int32))
            d_pat = buf[head:]
        return d_pat, d_off

    def gather_rows(self, res: ShardedResult, dst: int = 0):
        """Collective.  Assembles the flat answer on rank `dst`: (global_row_off, pairs) with row q = the shard rows of
        q concatenated in rank order = string_index::query(q) on the whole corpus.  Other ranks get None."""
        npat = res.npat
        totals = res.shard_rows.sum(dim=1).cpu().tolist()
        if self.rank == dst:
            parts = []
            for r in range(self.world):
                if r == dst:
                    parts.append(res.pairs)
                else:
                    buf = torch.empty((totals[r], 2), dtype=torch.int64, device=self.device)
                    if totals[r]:
                        dist.recv(buf, src=self._global(r), group=self.group)
                    parts.append(buf)
            out = torch.empty((int(res.global_row_off[-1]), 2), dtype=torch.int64, device=self.device)
            # destination of entry j of shard r's row q: global_row_off[q] + sum_{r' < r} rows[r'][q] + j
            base = torch.zeros(npat, dtype=torch.int64, device=self.device)
            for r in range(self.world):
                rows = res.shard_rows[r]
                if totals[r]:
                    starts = res.global_row_off[:-1] + base                   # per pattern
                    local_off = torch.cumsum(rows, 0) - rows                  # this shard's own CSR offsets
                    q_of = torch.repeat_interleave(torch.arange(npat, device=self.device), rows)
                    pos = starts[q_of] + (torch.arange(totals[r], device=self.device) - local_off[q_of])
                    out[pos] = parts[r]
                base += rows
            return res.global_row_off, out
        if totals[self.rank]:
            dist.send(res.pairs.contiguous(), dst=self._global(dst), group=self.group)
        return None

    def close(self):
        self.local.close()


class _ResultOwner:
    """Frees the library-owned device result when the tensors that alias it go away."""

    def __init__(self, index: StringIndex, res):
        self.index, self.res = index, res

    def __del__(self):
        try:
            self.index.device_result_free(self.res)
        except Exception:
            pass
