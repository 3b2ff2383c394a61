#!/usr/bin/env python
"""bench.py — measures the string-index hot path on B200 (BASELINE.json metric: substring queries/sec over a
10 GB corpus; SA build GB/s vs HBM peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference ...                     # the reference's own CPU query() on the host cores

A "step" is one batched locate of the workload's pattern batch over the resident index.  Workloads
(SURVEY.md §8d): cfg3 = 10^8 docs x 100 B a-z (10 GB) + 10^6 five-byte patterns (the configuration the metric is
quoted on; fits one 180 GB B200), cfg2 = 10^7 docs (1 GB) + 10^5 patterns, cfg1 = 1000 x 1 KB + 100 patterns.
With N > 1 the cfg3 corpus is sharded by contiguous doc range over the ranks (BASELINE config 4), every rank
gets the whole pattern batch (NCCL broadcast), runs its shard and the per-pattern row lengths / occurrence
counts are merged with NCCL — total work is fixed, so scaling is "strong".

Prints ONE JSON line (rank 0).  `value` = queries/s with patterns and results resident in HBM; `e2e` = the same
through the host-buffer C-ABI call (cdb_locate_batch: H2D of the patterns, D2H of the full CSR result inside the
timed region).  `roofline` describes the dominant kernel, `cpu_baseline` the reference's query() timed on this
box's host cores in the same run.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": dict(nd=1_000, doclen=1024, npat=100, m=3, seed=1),
    "cfg2": dict(nd=10_000_000, doclen=100, npat=100_000, m=5, seed=2),
    "cfg3": dict(nd=100_000_000, doclen=100, npat=1_000_000, m=5, seed=3),
}
NBLOCKS = 8  # the corpus is defined as 8 equal doc-range blocks so that it is the same for every N
METRIC = "substring_queries_per_sec"
UNIT = "queries/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class DevArray:
    """__cuda_array_interface__ view of library-owned device memory (so torch can wrap it without a copy)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def make_patterns(w):
    from tests import corpora
    return corpora.uniform_patterns(w["npat"], w["m"], seed=1000 + w["seed"], hi=ord("a") + w.get("sigma", 26) - 1)


def make_shard(w, rank: int, world: int, device):
    """Shard `rank` of `world`: blocks [rank*8/world, (rank+1)*8/world) of the corpus, generated on the device."""
    import torch
    nd, L = w["nd"], w["doclen"]
    per = nd // NBLOCKS
    b0, b1 = rank * NBLOCKS // world, (rank + 1) * NBLOCKS // world
    if nd < NBLOCKS * 8:
        assert world == 1
        b0, b1, per = 0, 1, nd
    snd = (b1 - b0) * per
    text = torch.zeros(snd * L + 64, dtype=torch.uint8, device=device)  # 64 B of padding (cdb_build_device contract)
    for b in range(b0, b1):
        g = torch.Generator(device=device)
        g.manual_seed(w["seed"] * 1000 + b)
        lo = (b - b0) * per * L
        text[lo:lo + per * L] = torch.randint(97, 97 + w.get("sigma", 26), (per * L,), dtype=torch.uint8, device=device,
                                                   generator=g)
    doc_off = torch.arange(snd + 1, dtype=torch.int64, device=device) * L
    gdoc = torch.arange(b0 * per, b0 * per + snd, dtype=torch.int64, device=device)
    ids = (gdoc * 2654435761) % (1 << 40) + 10 ** 12  # unique, not in doc order
    return text, doc_off, ids, snd


def algorithmic_bytes_per_pattern(n, w, occ, d, m=0, directory_symbols=0):
    """64*S + w*occ + 24*d with S = 2*ceil(log2 n) probes.  When the prefix directory resolves the
    keyword (m <= its symbols) no probe is made: the search then reads the keyword and its offset, two directory
    entries, and writes the interval."""
    S = 2 * int(np.ceil(np.log2(max(n, 2))))
    search = 64.0 * S
    if directory_symbols and 0 < m <= directory_symbols:
        search, S = float(m + 8 + 2 * 8 + 2 * 8), 0
    return {"search": search, "gather": float(w) * occ + 24.0 * d, "S": S}


def pick_workload(name, world):
    if name != "auto":
        return name
    env = os.environ.get("CDB_BENCH_WORKLOAD")
    if env:
        return env
    try:
        import torch
        free, _tot = torch.cuda.mem_get_info()
        if world == 1 and free < 150 * (1 << 30):
            return "cfg2"
    except Exception:
        pass
    return "cfg3"



def brute_rows(text, ids, L, snd, kws):
    """Independent check, written in torch: (id, count) rows of every keyword in ascending doc index of THIS shard, from
    a plain scan of its text (no suffix array, no library code)."""
    import torch
    n = snd * L
    rows = []
    for kw in kws:
        m = len(kw)
        hit = text[: n - m + 1] == kw[0]
        for j in range(1, m):
            hit &= text[j: n - m + 1 + j] == kw[j]
        pos = torch.nonzero(hit).flatten()
        del hit
        pos = pos[(pos % L) <= L - m]  # an occurrence never crosses a document end
        docs, counts = torch.unique(pos // L, return_counts=True)
        rows.append(torch.stack([ids[docs], counts], dim=1).cpu().numpy())
    return rows


def sharded_row_parity(sh, text, ids, w, snd, pat, poff, rank, world, nsample=48):
    """Row-level parity of the sharded path (N > 1), inside the bench run: a sample of the batch's keywords is located
    collectively, the flat answer is assembled on rank 0 (gather_rows: shard rows concatenated in rank order) and
    compared with the concatenation of every shard's brute-force rows."""
    import torch.distributed as dist
    step = max(1, w["npat"] // nsample)
    kws = [bytes(pat[poff[q]:poff[q + 1]]) for q in range(0, w["npat"], step)][:nsample]
    res = sh.locate_batch(kws if rank == 0 else None, src=0)
    flat = sh.gather_rows(res, dst=0)
    mine = brute_rows(text, ids, w["doclen"], snd, kws)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank != 0:
        return None
    gro, gp = flat
    gro, gp = gro.cpu().numpy(), gp.cpu().numpy()
    bad = 0
    for q in range(len(kws)):
        want = np.concatenate([gathered[r][q] for r in range(world)]) if world else np.zeros((0, 2), np.int64)
        bad += not np.array_equal(gp[gro[q]:gro[q + 1]], want)
    return {"status": "ok" if bad == 0 else f"MISMATCH in {bad} rows", "patterns": len(kws), "pairs": int(gro[-1]),
            "how": "gather_rows on rank 0 vs per-shard brute-force scans (torch) concatenated in rank order"}


def spans_leg(ix, last, text, ids, d_pat, d_poff, npat, w, dev, hbm_peak, pat, poff, max_docs=32, steps=3):
    """cfg3's highlight part: for every keyword of the batch, merged highlight spans inside its first
    <= 32 hit documents — one cdb_locate_spans_batch_device call per step (10^6 requests, ~3.2e7 texts), results resident
    in HBM.  The doc index of an id comes from a side map, as an integration keeps one."""
    import torch
    import oracle
    L = w["doclen"]
    row_off = last.row_off
    rowlen = row_off[1:] - row_off[:-1]
    take = torch.clamp(rowlen, max=max_docs)
    ntext = int(take.sum())
    text_req = torch.repeat_interleave(torch.arange(npat, device=dev), take)
    starts = torch.cumsum(take, 0) - take
    within = torch.arange(ntext, device=dev) - starts[text_req]
    hit_ids = last.pairs[row_off[:-1][text_req] + within, 0].contiguous()
    sorted_ids, perm = torch.sort(ids)  # id -> doc index (ids are unique)
    text_doc = perm[torch.searchsorted(sorted_ids, hit_ids)].contiguous()
    del sorted_ids, perm, within, starts, hit_ids
    req_kw_off = torch.arange(npat + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def call():
        return ix.spans_batch_device(d_pat.data_ptr(), d_poff.data_ptr(), npat, req_kw_off.data_ptr(), npat, text_req.data_ptr(),
                                     text_doc.data_ptr(), ntext, stream)

    sp = call()
    total_spans = sp.total_spans
    # parity on a sample of the texts: spans == the reference highlighter's (oracle port of database.cpp:58-77)
    soff = torch.as_tensor(DevArray(sp.span_off, ntext + 1), device=dev)
    spans = torch.as_tensor(DevArray(sp.spans, 2 * max(total_spans, 1)), device=dev).view(-1, 2)
    ok = True
    for t in range(0, ntext, max(1, ntext // 200)):
        d, q = int(text_doc[t]), int(text_req[t])
        doc = text[d * L:(d + 1) * L].cpu().numpy().tobytes()
        got = spans[int(soff[t]):int(soff[t + 1])].cpu().numpy()
        ok = ok and np.array_equal(got, oracle.port.spans([bytes(pat[poff[q]:poff[q + 1]])], doc))
    ix.device_spans_free(sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ix.device_spans_free(call())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # algorithmic bytes: every text's document read once + its (request, doc) pair + keyword, 16 B per span written
    alg = ntext * (L + 16 + w["m"] + 8) + 16 * total_spans
    return {"texts_per_step": ntext, "spans_per_step": int(total_spans), "ms_per_step": ms, "texts_per_sec": ntext / (ms / 1e3),
            "spans_per_sec": total_spans / (ms / 1e3), "algorithmic_bytes_per_step": alg,
            "achieved_GBps": alg / (ms / 1e3) / 1e9, "frac_of_hbm_peak": alg / (ms / 1e3) / 1e9 / hbm_peak,
            "parity_with_oracle": "ok" if ok else "MISMATCH",
            "what": f"merged highlight spans of every keyword in its first <= {max_docs} hit documents, one batched call"}



def cfg5_shard_leg(dev, nbytes, seed=55, rank=0, world=1):
    """One shard of BASELINE configs[4] on one GPU: `nbytes` of valid UTF-8 in documents of up to 64 KB (64-bit elements,
    note-N1 layout: the search runs the reference's own recurrences), two numeric columns, and 10^5 requests
    `substring (4..12 bytes, half sampled / half random) AND year in a 10 % range` through cdb_filter with span [0,32).
    Checked on a sample against the doc-ordered locate rows (whose parity with the compiled reference at this shape is
    tests/test_gpu_cfg5.py) intersected with the numeric predicate in numpy."""
    import torch
    import coffeedb_b200 as cdb
    from tests import corpora
    if world > 1:
        import torch.distributed as dist
    t5, o5, i5, nd5, n5 = corpora.utf8_corpus_on_device(nbytes, seed=seed, device_index=dev.index or 0)
    i5 += rank * (1 << 40)  # ids are unique over the shards
    ix5 = cdb.StringIndex(device=dev.index or 0)
    t0 = time.perf_counter()
    ix5.build_device(t5.data_ptr(), o5.data_ptr(), i5.data_ptr(), nd5, torch.cuda.current_stream().cuda_stream, keep=(t5, o5, i5))
    wall = time.perf_counter() - t0
    b5, inf5, v5 = ix5.build_stats(), ix5.info(), ix5.verify_sa()
    rng = np.random.default_rng(501)  # the same requests on every rank (a broadcast batch); columns differ per shard
    crng = np.random.default_rng(900 + rank)
    h_ids = i5.cpu().numpy()
    year = crng.integers(1900, 2100, size=nd5).astype(np.int64)
    score = crng.random(nd5)
    ycol = cdb.NumericIndex(0, h_ids, year, device=dev.index or 0)
    scol = cdb.NumericIndex(1, h_ids, score, device=dev.index or 0)
    nreq = 100_000
    m = rng.integers(4, 13, size=nreq)
    starts = torch.from_numpy(np.random.default_rng(777).integers(0, n5 - 16, size=nreq)).to(dev)  # own generator: the shared one must advance identically on every rank
    win_t = t5[(starts.unsqueeze(1) + torch.arange(12, device=dev).unsqueeze(0)).reshape(-1)].contiguous()
    if world > 1:
        dist.broadcast(win_t, 0)  # the sampled keywords come from rank 0's shard and go to every rank
    win = win_t.cpu().numpy().reshape(nreq, 12)
    rnd = rng.integers(0x20, 0x7F, size=(nreq, 12), dtype=np.uint8)
    use_rnd = rng.random(nreq) < 0.5
    win[use_rnd] = rnd[use_rnd]
    koff = np.zeros(nreq + 1, np.int64)
    koff[1:] = np.cumsum(m)
    kw = np.concatenate([win[r, : m[r]] for r in range(nreq)]).astype(np.uint8)
    y0 = rng.integers(1900, 2080, size=nreq).astype(np.int64)
    ranges = np.stack([y0, np.zeros(nreq, np.int64), y0 + 20, np.zeros(nreq, np.int64)], axis=1)  # "[y0, y0+20)"
    terms = np.zeros(2 * nreq, cdb.TERM_DTYPE)
    terms["key"][0::2], terms["range"][0::2] = 0, -1
    terms["kw_begin"][0::2], terms["kw_end"][0::2] = koff[:-1], koff[1:]
    terms["key"][1::2], terms["range"][1::2] = 1, np.arange(nreq)
    rto = np.arange(nreq + 1, dtype=np.int64) * 2
    span = np.tile(np.array(FILTER_SPAN, np.int64), (nreq, 1))
    args = [pinned_copy(kw), pinned_copy(ranges), pinned_copy(terms), pinned_copy(rto)]

    def call(sp):
        return cdb.filter_raw([ix5, ycol, scol], args[0], args[1], args[2], args[3], None, sp)

    def step():
        """one batch: every shard answers the requests on its documents; the per-request match counts (what `count`
        reports, src/interface.cpp:289-300) are summed over the shards with NCCL"""
        res = call(span)
        m_host = np.ctypeslib.as_array(res.matched, shape=(nreq,))
        ret = res.total_pairs
        if world > 1:
            m_dev = torch.from_numpy(m_host).to(dev, non_blocking=False)
            dist.all_reduce(m_dev)
            tot = int(m_dev.sum())
        else:
            tot = int(m_host.sum())
        cdb.filter_result_free(res)
        return ret, tot

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 3
    for _ in range(steps):
        returned, matched_total = step()
    torch.cuda.synchronize()
    dt_t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    dt = float(dt_t.item())
    # parity on a sample: full answers (no span) against locate rows ∩ numeric predicate
    res = call(None)
    ro = np.ctypeslib.as_array(res.row_off, shape=(nreq + 1,))
    pr = np.ctypeslib.as_array(res.pairs, shape=(max(res.total_pairs, 1), 2))
    year_of = dict(zip(h_ids.tolist(), year.tolist()))
    ok, nonempty = True, 0
    sample = list(range(0, nreq, nreq // 300))
    lro, lpr = ix5.locate_batch([bytes(kw[koff[r]:koff[r + 1]]) for r in sample])
    for j, r in enumerate(sample):
        row = lpr[lro[j]:lro[j + 1]]
        want = sorted((int(i), int(c)) for i, c in row if y0[r] <= year_of[int(i)] < y0[r] + 20)
        got = [(int(a), int(b)) for a, b in pr[ro[r]:ro[r + 1]]]
        nonempty += bool(want)
        ok = ok and sorted(got) == want and [c for _i, c in got] == sorted((c for _i, c in got), reverse=True)
    cdb.filter_result_free(res)
    out = {"requests_per_sec": nreq / dt, "ms_per_step": dt * 1e3, "requests_per_step": nreq, "objects_matched_per_step": matched_total,
           "pairs_returned_per_step": int(returned), "parity_sample": "ok" if ok else "MISMATCH", "sample_requests": len(sample),
           "sample_requests_with_hits": nonempty, "docs": nd5, "corpus_bytes": n5, "sa_width": inf5["width"],
           "build_ms": b5["total_ms"], "build_wall_s": wall, "build_corpus_GB_per_s": n5 / 1e9 / (b5["total_ms"] / 1e3),
           "rounds": b5["rounds"], "chunks": b5["chunks"], "verified": bool(v5["ok"]), "signed_rule_pairs": v5["signed_rule_pairs"],
           "prefix_directory_symbols": ix5.prefix_directory()["symbols"],
           "what": "ONE of the 8 shards of BASELINE configs[4] (40 GB / 8): UTF-8 documents of up to 64 KB, 64-bit elements, "
                   "note-N1 layout; requests = substring (4..12 bytes, half sampled from the corpus, half random) AND "
                   "year in a 10 % range, span [0,32), through cdb_filter with pinned host buffers"}
    ycol.close()
    scol.close()
    ix5.close()
    return out


def query_pool_leg():
    """The reference server's calling pattern in C++ (tools/query_pool_bench.cpp): T worker threads, each calling
    string_index::query(keyword) — ONE keyword per call (src/database.cpp:387-393) — on a 1 GB index of the cfg2 shape
    (10^7 docs x 100 B, five-byte keywords, ~84 pairs per answer).  Calls are coalesced into device batches by cdb_query;
    lone calls take the small-batch path."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "tools", "_build", "query_pool_bench")
    if not os.path.exists(exe):
        return {"error": "tools/_build/query_pool_bench has not been built (__graft_entry__.build())"}
    r = subprocess.run([exe, "10000000", "100", "5", "2000"], capture_output=True, text=True, timeout=300)
    out = {"threads": {}, "what": "C++ worker threads calling string_index::query, one 5-byte keyword per call, 10^7 docs x 100 B"}
    for line in r.stdout.splitlines():
        m = re.match(r"\s*(\d+) threads:\s+(\d+) queries/s\s+\((\d+) queries in (\d+) device batches", line)
        if m:
            out["threads"][m.group(1)] = {"queries_per_sec": int(m.group(2)), "device_batches": int(m.group(4))}
    if not out["threads"]:
        out["error"] = (r.stdout + r.stderr)[-300:]
    return out


def secondary_workloads(sh, ix, text, snd, w, dev, hbm_peak):
    """Extra, driver-visible numbers beside the headline (N = 1): W8s on the same index — 8-byte keywords sampled from
    the corpus, longer than the prefix directory, so the search refines by binary search — and the cfg2
    workload (1 GB corpus, 10^5 keywords) on its own index."""
    import torch
    import coffeedb_b200 as cdb
    out = {}

    def timed(sh_, d_pat, d_poff, npat, steps=3):
        for _ in range(2):
            sh_.locate_batch(device_patterns=(d_pat, d_poff), src=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ph = {k: 0.0 for k in ("search_ms", "gather_ms", "translate_ms", "listing_ms", "total_ms")}
        e0.record()
        for _ in range(steps):
            res = sh_.locate_batch(device_patterns=(d_pat, d_poff), src=0)
            st = cdb.last_locate_stats()
            for k in ph:
                ph[k] += st[k] / steps
            pairs, occ = int(res.pairs.shape[0]), int(st["occurrences"])
            del res
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"queries_per_sec": npat / (ms / 1e3), "ms_per_step": ms, "phases_ms": ph, "pairs_per_step": pairs,
                "occurrences_per_step": occ}

    # W8s
    npat, L = w["npat"], w["doclen"]
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    docs_s = torch.randint(0, snd, (npat,), device=dev, generator=g)
    offs_s = torch.randint(0, L - 8 + 1, (npat,), device=dev, generator=g)
    idx = (docs_s * L + offs_s).unsqueeze(1) + torch.arange(8, device=dev).unsqueeze(0)
    d_pat = torch.zeros(npat * 8 + 8, dtype=torch.uint8, device=dev)
    d_pat[: npat * 8] = text[idx.reshape(-1)]
    d_poff = torch.arange(npat + 1, dtype=torch.int64, device=dev) * 8
    del idx, docs_s, offs_s
    r = timed(sh, d_pat, d_poff, npat)
    inf = ix.info()
    S = 2 * int(np.ceil(np.log2(max(inf["n"], 2))))
    r["algorithmic_bytes_per_step"] = (64.0 * S) * npat + inf["width"] * r["occurrences_per_step"] + 24.0 * r["pairs_per_step"]
    r["frac_of_hbm_peak"] = r["algorithmic_bytes_per_step"] / (r["phases_ms"]["total_ms"] / 1e3) / 1e9 / hbm_peak
    r["what"] = "10^6 eight-byte keywords sampled from the corpus (every one hits), same index; 64*S + w*occ + 24*d bytes"
    out["w8s"] = r
    del d_pat, d_poff
    # cfg2
    try:
        from coffeedb_b200.sharded import ShardedStringIndex
        w2 = dict(WORKLOADS["cfg2"])
        t2, o2, i2, snd2 = make_shard(w2, 0, 1, dev)
        sh2 = ShardedStringIndex(device=dev)
        sh2.build_device(t2.data_ptr(), o2.data_ptr(), i2.data_ptr(), snd2, torch.cuda.current_stream().cuda_stream, keep=(t2, o2, i2))
        b2 = sh2.local.build_stats()
        p2, po2 = make_patterns(w2)
        dp = torch.zeros(len(p2) + 8, dtype=torch.uint8, device=dev)
        dp[: len(p2)] = torch.from_numpy(p2).to(dev)
        dpo = torch.from_numpy(po2).to(dev)
        r2 = timed(sh2, dp, dpo, w2["npat"], steps=5)
        r2["build_ms"] = b2["total_ms"]
        r2["build_corpus_GB_per_s"] = w2["nd"] * w2["doclen"] / 1e9 / (b2["total_ms"] / 1e3)
        r2["what"] = "BASELINE configs[1]: 10^7 docs x 100 B, 10^5 five-byte keywords per step"
        out["cfg2"] = r2
        sh2.close()
        del t2, o2, i2, dp, dpo
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 - secondary numbers must not take the bench line down
        out["cfg2"] = {"error": repr(e)[:200]}
    # cfg5 shape on one GPU: 1 GiB of valid UTF-8 in documents of up to 64 KB — 64-bit elements, the note-N1 layout, no
    # prefix directory: the search runs the reference's own recurrences (src/index.cpp:262-287), one thread per keyword
    try:
        from coffeedb_b200.sharded import ShardedStringIndex
        from tests import corpora
        t5, o5, i5, nd5, n5 = corpora.utf8_corpus_on_device(1 << 30, seed=55, device_index=dev.index or 0)
        sh5 = ShardedStringIndex(device=dev)
        sh5.build_device(t5.data_ptr(), o5.data_ptr(), i5.data_ptr(), nd5, torch.cuda.current_stream().cuda_stream, keep=(t5, o5, i5))
        b5 = sh5.local.build_stats()
        v5 = sh5.local.verify_sa()
        np5 = 100_000
        g = torch.Generator(device=dev)
        g.manual_seed(77)
        starts = torch.randint(0, n5 - 8, (np5,), device=dev, generator=g)
        dp = torch.zeros(np5 * 8 + 8, dtype=torch.uint8, device=dev)
        dp[: np5 * 8] = t5[(starts.unsqueeze(1) + torch.arange(8, device=dev).unsqueeze(0)).reshape(-1)]
        dpo = torch.arange(np5 + 1, dtype=torch.int64, device=dev) * 8
        r5 = timed(sh5, dp, dpo, np5, steps=5)
        inf5 = sh5.local.info()
        r5.update({"build_ms": b5["total_ms"], "build_corpus_GB_per_s": n5 / 1e9 / (b5["total_ms"] / 1e3), "rounds": b5["rounds"],
                   "docs": nd5, "corpus_bytes": n5, "sa_width": inf5["width"], "verified": bool(v5["ok"]), "verify": v5,
                   "signed_rule_pairs": v5["signed_rule_pairs"], "prefix_directory_symbols": sh5.local.prefix_directory()["symbols"],
                   "what": "BASELINE configs[4] shape on one GPU (1 GiB): UTF-8 documents of up to 64 KB, note-N1 layout; "
                           "10^5 eight-byte keywords sampled from the corpus (they may straddle documents: then no hit)"})
        out["cfg5_shape"] = r5
        sh5.close()
        del t5, o5, i5, dp, dpo
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out["cfg5_shape"] = {"error": repr(e)[:200]}
    return out


# ------------------------------------------------------------------------------------------------ reference
def reference_index_from_gpu(w, verbose=False):
    """Full-size index for the reference's query(): corpus generated as in our arm, suffix array built on the GPU
    (the reference's own build of 10^10 suffixes takes tens of minutes) and injected into an unmodified
    string_index; query() itself is 100 % the reference's code on the host cores."""
    import torch
    import coffeedb_b200 as cdb
    import oracle
    dev = torch.device("cuda", 0)
    text, doc_off, ids, snd = make_shard(w, 0, 1, dev)
    ix = cdb.StringIndex(device=0)
    ix.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), snd, torch.cuda.current_stream().cuda_stream,
                    keep=(text, doc_off, ids))
    inf = ix.info()
    h_text = text[: inf["n"]].cpu().numpy()
    h_off = doc_off.cpu().numpy()
    h_ids = ids.cpu().numpy()
    sa = np.empty(max(inf["n"], 1), np.uint32 if inf["width"] == 4 else np.uint64)
    cdb._check(cdb.lib().cdb_export_sa(ix._h, sa.ctypes.data, sa.nbytes))
    sa = sa[: inf["n"]]
    ix.close()
    del text, doc_off, ids
    torch.cuda.empty_cache()
    r = oracle.Ref()
    r.add_many_borrowed(h_ids, h_text, h_off)
    r.adopt_sa(sa, inf["bits"])
    return r, inf


def reference_self_built(w, nd_cap=1_000_000):
    """No GPU: the reference builds its own index on a bounded prefix of the corpus."""
    import oracle
    from tests import corpora
    nd = min(w["nd"], nd_cap)
    text, off, ids = corpora.uniform(nd, w["doclen"], seed=w["seed"])
    r = oracle.Ref()
    r.add_many_borrowed(ids, text, off)
    t0 = time.perf_counter()
    r.build()
    return r, {"n": int(off[-1]), "nd": nd, "build_s": time.perf_counter() - t0}


def run_cfg5(args):
    """BASELINE configs[4]: 40 GB of UTF-8 documents (<= 64 KB each) sharded over the ranks (5 GB per GPU at N = 8), mixed
    numeric-range + substring constraints through cdb_filter on every shard, per-request match counts merged with NCCL.
    `python -m torch.distributed.run ... bench.py --cfg5 --gpus N`; prints one JSON line on rank 0."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    r = cfg5_shard_leg(dev, args.cfg5_bytes, seed=55 + rank, rank=rank, world=world)
    summary = torch.tensor([r["corpus_bytes"], r["docs"], r["build_ms"], 1.0 if r["parity_sample"] == "ok" and r["verified"] else 0.0],
                           dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(summary) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, summary)
    else:
        allr = [summary]
    if rank == 0:
        tot_bytes = sum(float(a[0]) for a in allr)
        out = {"metric": "cfg5_requests_per_sec", "value": r["requests_per_sec"], "unit": "requests/s", "n_gpus": world,
               "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "dtype": "int64", "data": "synthetic",
               "config": {"workload": "cfg5", "corpus_bytes_total": tot_bytes, "shards": world,
                          "docs_total": int(sum(float(a[1]) for a in allr)), "requests_per_step": r["requests_per_step"],
                          "parallelism": f"doc-range shards x{world}: the same request batch on every shard (cdb_filter: substring AND "
                                         "year range), per-request match counts summed with one NCCL all_reduce"},
               "objects_matched_per_step_all_shards": r["objects_matched_per_step"],
               "build_ms_slowest_shard": max(float(a[2]) for a in allr),
               "build_corpus_GB_per_s": tot_bytes / 1e9 / (max(float(a[2]) for a in allr) / 1e3),
               "all_shards_verified_and_parity_ok": all(float(a[3]) == 1.0 for a in allr),
               "rank0": r}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    wname = pick_workload(args.workload, 1)
    w = WORKLOADS[wname]
    pat, poff = make_patterns(w)
    threads = oracle.hardware_threads()
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    if have_gpu:
        ref, inf = reference_index_from_gpu(w)
        how = (f"reference query() (src/index.cpp:237-326, compiled unmodified) on the full {wname} index; the packed "
               f"suffix array was built on the GPU and injected because the reference build of n={inf['n']:.3g} suffixes "
               "takes tens of minutes")
    else:
        ref, inf = reference_self_built(w)
        how = f"reference build()+query() on the first {inf['nd']} documents of {wname} (no GPU to build the full index)"
    # every step runs the SAME batch as our arm's step (all npat keywords of the workload); on a slow host the step is
    # capped at ~20 s of host time and says so
    probe = min(w["npat"], 20_000)
    s, _tp, _to = ref.query_batch_timed(pat[: poff[probe]], poff[: probe + 1], threads)
    per_step = int(min(w["npat"], max(probe, 20.0 / max(s / probe, 1e-9))))
    times = []
    tp = to = 0
    for it in range(args.warmup + args.steps):
        s, tp, to = ref.query_batch_timed(pat[: poff[per_step]], poff[: per_step + 1], threads)
        if it >= args.warmup:
            times.append(s)
    total = float(np.sum(times))
    val = per_step * len(times) / total
    # beside the headline (which stays query() alone, the less work of the two): the same keywords as one-keyword requests
    # through the rest of the reference's request path — the sorts of filter() and span [0,32), what our arm's e2e call does
    fs = None
    try:
        nf = int(min(per_step, 200_000))
        s2, _r, _p, _m = ref.filter_span_batch(pat[: poff[nf]], poff[: nf + 1], FILTER_SPAN[0], FILTER_SPAN[1], threads, keep=False)
        fs = {"value": nf / s2, "unit": UNIT, "sample": f"{nf} requests",
              "what": "query() + std::ranges::sort + std::sort by descending $correlation + span [0,32) per request "
                      "(src/interface.cpp:82,143-146,196-209 restated around the unmodified query(), oracle/ref_harness.cpp)"}
    except Exception as e:  # noqa: BLE001
        fs = {"error": repr(e)[:200]}
    # the reference's OWN build, on a bounded prefix of the same corpus (the full 10^10-suffix build takes tens of
    # minutes on host cores): string_index::build() (src/index.cpp:178-236), all host threads
    rb = None
    try:
        r2, inf2 = reference_self_built(w, nd_cap=1_000_000)
        rb = {"docs": inf2["nd"], "corpus_bytes": inf2["n"], "seconds": inf2["build_s"],
              "corpus_GB_per_s": inf2["n"] / 1e9 / inf2["build_s"], "threads": threads,
              "what": "string_index::build() of the unmodified reference on 10^6 documents of the workload's shape (100 MB)"}
        r2.close()
    except Exception as e:  # noqa: BLE001
        rb = {"error": repr(e)[:200]}
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": wname, "docs": w["nd"], "doc_bytes": w["doclen"], "corpus_bytes": w["nd"] * w["doclen"],
                   "patterns_per_step": per_step, "pattern_bytes": w["m"], "n_suffixes_per_gpu": inf["n"],
                   "sa_width": inf.get("width"),
                   "l2": "host caches; inputs far larger than any cache",
                   "parallelism": f"{threads} host threads pulling 64-keyword blocks (the reference's httplib pool pattern)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{per_step} of {w['npat']} patterns per step; {how}",
                         "pairs_per_step": tp, "occurrences_per_step": to},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "filter_span": fs,
        "reference_build": rb,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import coffeedb_b200 as cdb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — coffeedb_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wname = pick_workload(args.workload, world)
    w = dict(WORKLOADS[wname])
    if args.npat:
        w["npat"] = args.npat
    if args.sigma != 26:  # profiling aid only: smaller alphabet = longer intervals at a smaller corpus
        w["sigma"] = args.sigma
        wname += f"-sigma{args.sigma}"
    hbm_peak, peak_src = peaks()

    # ---- corpus shard + index build (setup; build throughput is reported, not part of the locate step)
    text, doc_off, ids, snd = make_shard(w, rank, world, dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    from coffeedb_b200.sharded import ShardedStringIndex
    sh = ShardedStringIndex(device=dev)  # one doc-range shard per rank (world == 1: the whole corpus)
    ix = sh.local
    t0 = time.perf_counter()
    sh.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), snd, stream, keep=(text, doc_off, ids))
    build_wall = time.perf_counter() - t0
    inf, bst = ix.info(), ix.build_stats()
    n_shard, width = inf["n"], inf["width"]
    # independent device-side check of the built suffix array: every adjacent pair under the reference's comparator
    # (src/index.cpp:92-93) + permutation bit map — not part of any timed region
    t0 = time.perf_counter()
    verify = ix.verify_sa() if args.verify else None
    verify_s = time.perf_counter() - t0
    bst_warm = None
    if args.rebuild:  # second build of the same corpus: device memory already touched once, allocator warm
        sh.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), snd, stream, keep=(text, doc_off, ids))
        bst_warm = ix.build_stats()

    # ---- patterns: pinned host copy (e2e) and device copy (value)
    pat, poff = make_patterns(w)
    npat = w["npat"]
    if args.patterns == "w8s":
        # secondary workload W8s: 8-byte substrings sampled from the corpus (every pattern hits; on this
        # rank's shard when sharded).  Keywords longer than the prefix directory: the search refines by binary search.
        g = torch.Generator(device=dev)
        g.manual_seed(4242)
        L = w["doclen"]
        docs_s = torch.randint(0, snd, (npat,), device=dev, generator=g)
        offs_s = torch.randint(0, L - 8 + 1, (npat,), device=dev, generator=g)
        idx = (docs_s * L + offs_s).unsqueeze(1) + torch.arange(8, device=dev).unsqueeze(0)
        pat = text[idx.reshape(-1)].cpu().numpy()
        poff = np.arange(npat + 1, dtype=np.int64) * 8
        w["m"] = 8
        wname += "-w8s"
    h_pat = torch.from_numpy(pat).pin_memory()
    h_poff = torch.from_numpy(poff).pin_memory()
    d_pat = torch.zeros(len(pat) + 8, dtype=torch.uint8, device=dev)
    d_poff = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
    if rank == 0:
        d_pat[: len(pat)].copy_(h_pat, non_blocking=True)
        d_poff.copy_(h_poff, non_blocking=True)

    def step():
        """One pass of the hot path over one pattern batch, inputs resident in HBM: pattern broadcast from rank 0
        (N > 1), local locate, NCCL merge of per-pattern row lengths and occurrence totals (rows stay sharded)."""
        res = sh.locate_batch(device_patterns=(d_pat, d_poff), src=0)
        return int(res.pairs.shape[0]), None, res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    phase = {k: 0.0 for k in ("search_ms", "gather_ms", "large_ms", "tail_ms", "translate_ms", "listing_ms", "total_ms")}
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = cdb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pairs, _occ, res = step()
        st = cdb.last_locate_stats()
        for k in phase:
            phase[k] += st[k]
        del res  # the CSR result goes back to the stream-ordered pool before the next step allocates its own
    e1.record()
    barrier()
    launches = cdb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if getattr(sh, "_trace", None):
        tr = np.array(sh._trace[-args.steps:])
        print(f"[rank {rank}] broadcast / locate / all_gather ms per batch: {tr.mean(axis=0).round(3).tolist()}", file=sys.stderr)
    st_last = cdb.last_locate_stats()
    occs = st_last["occurrences"]  # this shard's occurrences per step
    _p, _o, last = step()  # one more (untimed) step whose result is kept for the whole-job totals
    global_pairs = int(last.global_row_off[-1])
    global_occ = int(last.occurrences.sum())
    spans_info = None
    if world == 1 and args.spans:
        try:
            spans_info = spans_leg(ix, last, text, ids, d_pat, d_poff, npat, w, dev, hbm_peak, pat, poff)
        except Exception as e:  # noqa: BLE001 - must not take the bench line down
            spans_info = {"error": repr(e)[:300]}
    del last
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = npat * args.steps / (ms_total / 1e3)
    for k in phase:
        phase[k] /= args.steps
    listed_rows, listed_pairs = st_last["nlisted"], st_last["listed_pairs"]
    linfo = ix.listing_info(0)

    # ---- the same batches on the suffix-array path alone (CDB_LISTING_USE=0: search -> gather -> translate), so that the
    # line also carries what keywords NOT of the directory's length cost, and stays comparable with round 1
    sa_path = None
    if listed_rows:
        os.environ["CDB_LISTING_USE"] = "0"
        try:
            for _ in range(args.warmup):
                step()
            ph2 = {k: 0.0 for k in phase}
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.steps):
                p2, _occ2, res2 = step()
                st2 = cdb.last_locate_stats()
                for k in ph2:
                    ph2[k] += st2[k]
                del res2
            f1.record()
            barrier()
            ms2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
            for k in ph2:
                ph2[k] /= args.steps
            sa_ms = float(ms2.item()) / args.steps
            sa_bytes = (algorithmic_bytes_per_pattern(n_shard, width, st2["occurrences"] / npat, p2 / npat, w["m"],
                                                      ix.prefix_directory()["symbols"]))
            sa_path_bytes = (sa_bytes["search"] + sa_bytes["gather"]) * npat
            sa_path = {"value": npat / (sa_ms / 1e3), "unit": UNIT, "ms_per_step": sa_ms, "phases_ms": ph2,
                       "pairs_equal": bool(p2 == pairs),
                       "algorithmic_bytes_per_step": sa_path_bytes,
                       "frac_of_hbm_peak": sa_path_bytes / (ph2["total_ms"] / 1e3) / 1e9 / hbm_peak,
                       "translate_frac": 24.0 * (p2 / npat) * npat / (ph2["translate_ms"] / 1e3) / 1e9 / hbm_peak if ph2["translate_ms"] > 0 else None,
                       "gather_frac": width * (st2["occurrences"] / npat) * npat / (ph2["gather_ms"] / 1e3) / 1e9 / hbm_peak if ph2["gather_ms"] > 0 else None,
                       "what": "the same batches with the document listing bypassed (CDB_LISTING_USE=0): search_kernel -> gather_kernel "
                               "-> translate_kernel, w*occ + 24*d algorithmic bytes per keyword (SURVEY.md 8d)"}
        finally:
            del os.environ["CDB_LISTING_USE"]

    # ---- e2e: host buffers through cdb_locate_batch, H2D + D2H inside the timed region
    def e2e_step():
        res = ix.locate_batch_raw(h_pat.numpy(), h_poff.numpy())  # pinned host buffers
        b = (res.npat + 1) * 8 + res.total_pairs * 16
        first = res.pairs[0] if res.total_pairs else 0  # touch the host result
        ix.result_free(res)
        return b, first

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        d2h_bytes, _ = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        d2h_bytes, _ = e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = npat * e2e_steps / float(e2e_s.item())
    h2d_bytes = int(pat.nbytes + poff.nbytes)

    # ---- e2e through cdb_filter (N = 1): the request the reference's server actually answers — filter() over one keyword,
    # the descending-$correlation sort and a span (src/interface.cpp:46-147, 196-209) — so that only result[0, 32) of
    # every request crosses PCIe instead of the full rows.  Host buffers pinned, H2D of the packed requests and D2H of the
    # slices inside the timed region.
    e2e_full = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                "steps": e2e_steps, "call": "cdb_locate_batch: full (id, count) rows of every keyword to host memory"}
    e2e_main = dict(e2e_full)
    filt = None
    if world == 1 and args.filter:
        try:
            filt = filter_leg(ix, pat, poff, npat, e2e_steps)
            e2e_main = {"value": filt["value"], "unit": UNIT, "h2d_bytes_per_step": filt["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": filt["d2h_bytes_per_step"], "steps": e2e_steps, "call": filt["call"]}
        except Exception as e:  # noqa: BLE001 - falls back to the full-row number, and says so
            filt = {"error": repr(e)[:300]}

    # ---- latency of ONE keyword through the host ABI (the reference's own calling pattern: query() per request)
    one_pat = np.ascontiguousarray(pat[: poff[1]])
    one_off = np.ascontiguousarray(poff[:2])
    for _ in range(20):
        ix.result_free(ix.locate_batch_raw(one_pat, one_off))
    t0 = time.perf_counter()
    nlat = 200
    for _ in range(nlat):
        ix.result_free(ix.locate_batch_raw(one_pat, one_off))
    single_us = (time.perf_counter() - t0) / nlat * 1e6

    # ---- the reference server's calling pattern: a pool of worker threads, ONE keyword per call (database.cpp:387-393);
    # cdb_query coalesces concurrent callers into device batches.  Reported beside the headline, never part of it.
    concurrent = None
    if rank == 0:
        try:
            concurrent = concurrent_single_queries(ix, pat, poff, threads=16, per_thread=250)
        except Exception as e:  # noqa: BLE001 - an auxiliary number must not take the bench line down
            concurrent = {"error": repr(e)[:200]}

    # ---- roofline of the dominant kernel (CUDA-event phase times measured inside the library on `stream`)
    occ_pp = occs / npat
    d_pp = pairs / npat
    alg = algorithmic_bytes_per_pattern(n_shard, width, occ_pp, d_pp, w["m"], ix.prefix_directory()["symbols"])
    # algorithmic bytes per launch, split over the kernels of the path: the search reads 64*S per
    # pattern; gather_kernel (phase A) reads the SA interval (w*occ); translate_kernel (phase B) reads ids[] and
    # writes the (id, count) pairs (24*d).  The compact intermediate rows between A and B (8*d written, 8*d read)
    # are this design's own overhead and are NOT counted as algorithmic bytes.
    # Rows answered from the document listing (keywords of exactly the directory's length): listing_emit_kernel reads
    # 4 + hi_bytes per occurrence and writes the 16-byte pair — no ids[] lookup, no intermediate; those are the bytes it
    # is charged with (fewer than the survey's w*occ + 24*d, which assumes the suffix-array interval is read).
    lw = 4 + (linfo["hi_bytes"] if linfo["present"] else 0)
    frac_listed = listed_pairs / pairs if pairs else 0.0  # share of the step's result that came from the listing
    occ_listed = occ_pp * frac_listed if listed_rows else 0.0  # (all rows or none in the bench's batches)
    d_listed = d_pp * frac_listed if listed_rows else 0.0
    kernels = {
        "search_kernel": (phase["search_ms"], alg["search"] * npat),
        "gather_kernel": (phase["gather_ms"], width * (occ_pp - occ_listed) * npat),
        "translate_kernel": (phase["translate_ms"], 24.0 * (d_pp - d_listed) * npat),
        "listing_emit_kernel": (phase["listing_ms"], (lw * occ_listed + 16.0 * d_listed) * npat),
    }
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[dom]
    achieved = dom_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    path_bytes = alg["search"] * npat + sum(v[1] for k, v in kernels.items() if k != "search_kernel")
    survey_bytes = (alg["search"] + alg["gather"]) * npat
    # DRAM traffic of the dominant kernel per launch, from the committed ncu capture of the same configuration
    traffic, traffic_src = None, None
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic_cfg3.json")))  # the newest capture (by round tag)
    tpath = tfiles[-1] if tfiles else ""
    if world == 1 and wname == "cfg3" and tpath and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if dom in tj.get("kernels", {}):
            traffic, traffic_src = tj["kernels"][dom]["traffic"], tj["source"]
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
        "phases_ms": phase,
        "kernels": {k: {"ms": v[0], "algorithmic_bytes": v[1],
                        "achieved": (v[1] / (v[0] / 1e3) / 1e9 if v[0] > 0 else 0.0),
                        "frac": (v[1] / (v[0] / 1e3) / 1e9 / hbm_peak if v[0] > 0 else 0.0)} for k, v in kernels.items()},
        "path": {"algorithmic_bytes_per_step": path_bytes, "achieved": path_bytes / (phase["total_ms"] / 1e3) / 1e9,
                 "frac": path_bytes / (phase["total_ms"] / 1e3) / 1e9 / hbm_peak,
                 "note": "bytes the path has to move per step: the search's (64*S probes, or 45 B when the prefix directory "
                         "resolves the keyword without probing) + per keyword (4 + hi_bytes)*occ + 16*d when its row is "
                         "streamed from the document listing, w*occ + 24*d (SURVEY.md 8d) when it is gathered from the suffix array",
                 "survey_formula_bytes_per_step": survey_bytes,
                 "survey_formula_frac": survey_bytes / (phase["total_ms"] / 1e3) / 1e9 / hbm_peak,
                 "probes_per_pattern": alg["S"]},
        "listing": {"present": linfo["present"], "bytes_per_suffix": lw if linfo["present"] else 0, "bytes": linfo["bytes"],
                    "build_ms": linfo["build_ms"], "rows_per_step": listed_rows, "pairs_per_step": listed_pairs},
    }

    # ---- CPU baseline: the reference's query() on this box's host cores, same index, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(ix, text, doc_off, ids, inf, pat, poff, w, wname, filt)
        except Exception as e:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e!r}"}

    # ---- N > 1: row-level parity of the sharded answer, inside this run (the GPU test box has one GPU)
    parity_sharded = None
    if world > 1:
        try:
            parity_sharded = sharded_row_parity(sh, text, ids, w, snd, pat, poff, rank, world)
        except Exception as e:  # noqa: BLE001
            parity_sharded = {"status": f"failed: {e!r}"[:300]}
    # ---- secondary workloads (N = 1, the 10 GB configuration): W8s and cfg2 as extra keys
    extras = None
    if world == 1 and args.extras and wname == "cfg3" and args.patterns == "w5":
        try:
            extras = secondary_workloads(sh, ix, text, snd, w, dev, hbm_peak)
        except Exception as e:  # noqa: BLE001
            extras = {"error": repr(e)[:300]}

    build_ms = torch.tensor([bst["total_ms"], bst_warm["total_ms"] if bst_warm else 0.0,
                             0.0 if (verify is None or verify["ok"]) else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(build_ms, op=dist.ReduceOp.MAX)  # every shard builds concurrently: the job takes the slowest
    verified_all = float(build_ms[2].item()) == 0.0
    build_ms, rebuild_ms = float(build_ms[0].item()), float(build_ms[1].item())
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic",
            "config": {"workload": wname + ("" if world == 1 else f"-sharded{world}"), "docs": w["nd"],
                       "doc_bytes": w["doclen"], "corpus_bytes": w["nd"] * w["doclen"], "patterns_per_step": npat,
                       "pattern_bytes": w["m"], "n_suffixes_per_gpu": n_shard, "sa_width": width,
                       "l2": "inputs (suffix array + text) far larger than L2, no flush needed",
                       "answer_path": ("document listing: every keyword is as long as the prefix directory is deep, its row is streamed "
                                       "from the per-bucket id list (sa_path = the same batches through the suffix-array path)"
                                       if listed_rows == npat else
                                       ("suffix array (search, gather, translate)" if not listed_rows else "mixed: listing + suffix array")),
                       "parallelism": "replica" if world == 1 else f"doc-range shards x{world}, NCCL pattern bcast + count merge"},
            "e2e": e2e_main,
            "e2e_full_rows": e2e_full,
            "sa_path": sa_path,
            "filter": filt,
            "gpu_launches": int(launches),
            "single_query_latency_us": single_us,
            "concurrent_single_queries": concurrent,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "build": {"corpus_GB_per_s": w["nd"] * w["doclen"] / 1e9 / (build_ms / 1e3), "ms": build_ms,
                      "note": "whole corpus / slowest shard's build (CUDA events inside cdb_build_device)",
                      "sort_ms": bst["sort_ms"], "rounds": bst["rounds"], "chunks": bst["chunks"], "wall_s": build_wall,
                      "rebuild_ms": rebuild_ms if bst_warm else None,
                      "rebuild_corpus_GB_per_s": (w["nd"] * w["doclen"] / 1e9 / (rebuild_ms / 1e3)) if bst_warm else None,
                      "verified": (bool(verified_all) if verify is not None else None),
                      "verify": (dict(verify, seconds=verify_s, what="cdb_verify_sa on every shard: adjacent suffix pairs under "
                                      "std::string_view < (src/index.cpp:92-93), permutation bit map; rank 0's counters shown")
                                 if verify is not None else None),
                      "compulsory_bytes": n_shard * (1 + width),
                      "frac_of_hbm_peak": n_shard * (1 + width) / 1e9 / (bst["total_ms"] / 1e3) / hbm_peak},
            "pairs_per_step": global_pairs, "occurrences_per_step": global_occ,
            "spans": spans_info,
            "parity_sharded": parity_sharded,
            "extras": extras,
        }
    ix.close()
    if rank == 0:
        # one shard of config 5 at its real size, after the 10 GB index has been released (N = 1 only)
        if world == 1 and args.extras and wname == "cfg3" and args.patterns == "w5":
            try:
                del sh, text, doc_off, ids
                import gc
                gc.collect()
                torch.cuda.empty_cache()
                out.setdefault("extras", {})
                if out["extras"] is None:
                    out["extras"] = {}
                out["extras"]["cfg5_shard"] = cfg5_shard_leg(dev, 5_000_000_000)
                out["extras"]["query_pool_cpp"] = query_pool_leg()
            except Exception as e:  # noqa: BLE001
                if isinstance(out.get("extras"), dict):
                    out["extras"]["cfg5_shard"] = {"error": repr(e)[:300]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def concurrent_single_queries(ix, pat, poff, threads, per_thread):
    """`threads` host threads, each issuing `per_thread` single-keyword query() calls (ctypes drops the GIL inside the
    call; the Python glue around it does not, so this is a lower bound of what a C++ server pool would see)."""
    import threading
    total = threads * per_thread
    kws = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(total)]
    before = ix.query_stats()
    errs = []

    def work(k):
        try:
            for kw in kws[k::threads]:
                ix.query_array(kw)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    if errs:
        raise RuntimeError(errs[0])
    after = ix.query_stats()
    return {"threads": threads, "queries": total, "queries_per_sec": total / dt,
            "device_batches": after["batches"] - before["batches"], "largest_batch": after["largest"]}


def pinned_copy(a: np.ndarray) -> np.ndarray:
    """numpy view of a page-locked copy of `a` (the e2e legs read their inputs from pinned host memory)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t.numpy().view(a.dtype).reshape(a.shape)


FILTER_SPAN = (0, 32)


def filter_requests(pat, poff, npat):
    """npat one-keyword requests with span [0, 32) as packed cdb_filter arrays (pinned)."""
    import coffeedb_b200 as cdb
    terms = np.zeros(npat, cdb.TERM_DTYPE)
    terms["range"] = -1
    terms["kw_begin"], terms["kw_end"] = poff[:npat], poff[1:npat + 1]
    span = np.tile(np.array(FILTER_SPAN, np.int64), (npat, 1))
    return (pinned_copy(np.ascontiguousarray(pat[: poff[npat]])), pinned_copy(terms), pinned_copy(np.arange(npat + 1, dtype=np.int64)),
            pinned_copy(span))


def filter_leg(ix, pat, poff, npat, steps):
    """End-to-end through cdb_filter: npat requests {key: keyword, span: [0,32)} per step from pinned host arrays, answers
    (id, $correlation)[0:32] in the reference's order + matched counts back in host memory."""
    import coffeedb_b200 as cdb
    kw, terms, rto, span = filter_requests(pat, poff, npat)
    none = np.zeros((0, 4), np.int64)

    def one():
        res = cdb.filter_raw([ix], kw, none, terms, rto, None, span)
        nbytes = (npat + 1) * 8 + res.total_pairs * 16 + npat * 8
        first = res.pairs[0] if res.total_pairs else 0  # touch the host result
        tot = res.total_pairs
        cdb.filter_result_free(res)
        return nbytes, tot, first

    for _ in range(3):  # the first call also builds the id-order tables of the index (once per index)
        one()
    launches0 = cdb.launch_count()
    per_step = []
    t0 = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        d2h, tot, _f = one()
        per_step.append((time.perf_counter() - t1) * 1e3)
    dt = time.perf_counter() - t0
    lst = cdb.last_locate_stats()  # the id-ordered locate inside the last call
    return {"value": npat * steps / dt, "ms_per_step": dt / steps * 1e3, "ms_each_step": [round(x, 3) for x in per_step],
            "locate_phases_ms": {k: lst[k] for k in ("search_ms", "gather_ms", "translate_ms", "listing_ms", "total_ms")},
            "h2d_bytes_per_step": int(kw.nbytes + terms.nbytes + rto.nbytes + span.nbytes), "d2h_bytes_per_step": int(d2h),
            "pairs_returned_per_step": int(tot), "launches_per_step": (cdb.launch_count() - launches0) // steps,
            "call": f"cdb_filter: {npat} requests {{key: keyword, span: [{FILTER_SPAN[0]},{FILTER_SPAN[1]})}} -> (id, $correlation) in "
                    "the order of src/interface.cpp:143-146, rows located in id order, ids not in doc order"}


def cpu_baseline(ix, text, doc_off, ids, inf, pat, poff, w, wname, filt=None):
    import coffeedb_b200 as cdb
    import oracle
    h_text = text[: inf["n"]].cpu().numpy()
    h_off = doc_off.cpu().numpy()
    h_ids = ids.cpu().numpy()
    sa = np.empty(max(inf["n"], 1), np.uint32 if inf["width"] == 4 else np.uint64)
    cdb._check(cdb.lib().cdb_export_sa(ix._h, sa.ctypes.data, sa.nbytes))
    sa = sa[: inf["n"]]
    ref = oracle.Ref()
    ref.add_many_borrowed(h_ids, h_text, h_off)
    ref.adopt_sa(sa, inf["bits"])
    threads = oracle.hardware_threads()
    probe = min(w["npat"], 20_000)
    s, _a, _b = ref.query_batch_timed(pat[: poff[probe]], poff[: probe + 1], threads)
    nsample = int(min(w["npat"], max(probe, 12.0 / max(s / probe, 1e-9))))
    s, tp, to = ref.query_batch_timed(pat[: poff[nsample]], poff[: nsample + 1], threads)
    # full-size parity check on the same sample: identical totals of (id,count) pairs and occurrences
    row_off, pairs = ix.locate_batch(pat[: poff[nsample]], poff[: nsample + 1])
    ok = (len(pairs) == tp) and (int(pairs[:, 1].sum()) == to)
    # and identical rows on a handful of patterns
    for q in range(0, nsample, max(1, nsample // 16)):
        got = pairs[row_off[q]:row_off[q + 1]]
        ok = ok and np.array_equal(got, ref.query(bytes(pat[poff[q]:poff[q + 1]])))
    # the same requests as the e2e leg (cdb_filter): query() + the sorts + the span, timed on the same sample, and the
    # answers compared element by element with cdb_filter's (order included: ties follow std::sort)
    fs = None
    if filt is not None and "error" not in filt:
        try:
            nf = min(nsample, 200_000)
            s2, _r, _p, _m = ref.filter_span_batch(pat[: poff[nf]], poff[: nf + 1], FILTER_SPAN[0], FILTER_SPAN[1], threads, keep=False)
            ncheck = min(nf, 20_000)
            _s, r_off, r_pairs, r_matched = ref.filter_span_batch(pat[: poff[ncheck]], poff[: ncheck + 1], FILTER_SPAN[0],
                                                                  FILTER_SPAN[1], threads)
            kw, terms, rto, span = filter_requests(pat, poff, ncheck)
            res = cdb.filter_raw([ix], kw, np.zeros((0, 4), np.int64), terms, rto, None, span)
            g_off = np.ctypeslib.as_array(res.row_off, shape=(ncheck + 1,))
            g_pairs = np.ctypeslib.as_array(res.pairs, shape=(max(res.total_pairs, 1), 2))[: res.total_pairs]
            g_matched = np.ctypeslib.as_array(res.matched, shape=(ncheck,))
            same = (np.array_equal(g_off, r_off) and np.array_equal(g_pairs, r_pairs) and np.array_equal(g_matched, r_matched))
            cdb.filter_result_free(res)
            fs = {"value": nf / s2, "unit": UNIT, "sample": f"{nf} requests", "requests_compared": ncheck,
                  "parity_with_gpu": "ok" if same else "MISMATCH",
                  "what": "query() + std::ranges::sort + std::sort by descending $correlation + span [0,32) per request "
                          "(src/interface.cpp:82,143-146,196-209 restated around the unmodified query(), oracle/ref_harness.cpp)"}
        except Exception as e:  # noqa: BLE001
            fs = {"error": repr(e)[:300]}
    ref.close()
    return {"value": nsample / s, "unit": UNIT, "cores": threads, "kind": "reference", "filter_span": fs,
            "sample": (f"{nsample} of {w['npat']} patterns on the full {wname} index; reference query() "
                       "(src/index.cpp:237-326, compiled unmodified) over the GPU-built packed suffix array, "
                       f"{threads} host threads pulling 64-pattern blocks"),
            "pairs": int(tp), "occurrences": int(to), "parity_with_gpu": "ok" if ok else "MISMATCH"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--npat", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--patterns", default="w5", choices=["w5", "w8s"],
                    help="w5 = uniform 5-byte keywords (the metric's workload); w8s = sampled 8-byte substrings")
    ap.add_argument("--no-rebuild", dest="rebuild", action="store_false",
                    help="skip the second (warm) build of the same corpus (build.rebuild_ms)")
    ap.add_argument("--no-verify", dest="verify", action="store_false", help="skip cdb_verify_sa after the build")
    ap.add_argument("--no-spans", dest="spans", action="store_false", help="skip the highlight-span leg")
    ap.add_argument("--no-filter", dest="filter", action="store_false", help="e2e through cdb_locate_batch (full rows) only")
    ap.add_argument("--no-extras", dest="extras", action="store_false", help="skip the secondary workloads (W8s, cfg2)")
    ap.add_argument("--cfg5", action="store_true", help="BASELINE configs[4] instead of the headline workload (see run_cfg5)")
    ap.add_argument("--cfg5-bytes", type=int, default=5_000_000_000, help="bytes of UTF-8 text per shard for --cfg5")
    ap.add_argument("--sigma", type=int, default=26, help="alphabet size (profiling aid; the named workloads use 26)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.cfg5:
        run_cfg5(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
