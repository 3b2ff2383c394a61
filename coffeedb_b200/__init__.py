"""coffeedb_b200 — B200-native string index (suffix-array build + batched substring locate) behind
CoffeeDB's ``string_index`` surface (src/index.h:54-86 of the reference).

This package is a thin ctypes binding of the C-ABI library ``libcoffeedb_b200.so`` (include/coffeedb_b200.h).
All computation happens in hand-written sm_100a CUDA kernels inside that library; there is no CPU
fallback — importing works anywhere (so the symbol-export test can run without a GPU) but every compute
call raises if the library or a CUDA device is missing.

``StringIndex`` mirrors the reference class: ``add(id, value)``, ``build()``, ``query(keyword)`` returning the
list of ``(id, count)`` pairs in ascending doc index, raising ``RuntimeError`` with the reference's messages
(src/index.cpp:196,199,240).  ``locate_batch`` is the batched form of ``query``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CDB_LIB") or os.path.join(HERE, "libcoffeedb_b200.so")  # CDB_LIB: A/B runs of a variant build
CSRC = os.path.join(HERE, "csrc")

# every symbol include/coffeedb_b200.h declares
EXPORTS = [
    "cdb_last_error", "cdb_version", "cdb_device_count", "cdb_create", "cdb_destroy", "cdb_add", "cdb_add_many", "cdb_staging_stats",
    "cdb_build", "cdb_build_device", "cdb_save", "cdb_build_or_load", "cdb_info", "cdb_prefix_directory", "cdb_listing_info", "cdb_export_sa", "cdb_sa_device_ptr", "cdb_locate_batch",
    "cdb_result_free", "cdb_locate_batch_device", "cdb_locate_batch_device_ex", "cdb_device_result_free", "cdb_locate_spans", "cdb_locate_spans_batch",
    "cdb_locate_spans_batch_device", "cdb_device_spans_free", "cdb_spans_free",
    "cdb_splice", "cdb_verify_sa", "cdb_compare_sa", "cdb_build_stats", "cdb_last_locate_stats", "cdb_last_locate_stats_ex", "cdb_launch_count", "cdb_query", "cdb_query_stats", "cdb_trim",
    "cdb_numeric_create", "cdb_numeric_destroy", "cdb_numeric_query", "cdb_filter", "cdb_filter_result_free",
    "cdb_sharded_create", "cdb_sharded_destroy", "cdb_sharded_add", "cdb_sharded_add_many", "cdb_sharded_build",
    "cdb_sharded_locate_batch", "cdb_sharded_shard", "cdb_sharded_count",
]

CDB_OK = 0


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("compat_signed", C.c_int32), ("workspace_bytes", C.c_int64),
                ("keep_host_copy", C.c_int32), ("reserved", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("npat", C.c_int64), ("total_pairs", C.c_int64), ("total_occurrences", C.c_int64),
                ("row_off", C.POINTER(C.c_int64)), ("pairs", C.POINTER(C.c_int64)), ("_owner", C.c_void_p)]


class DeviceResult(C.Structure):
    _fields_ = [("npat", C.c_int64), ("total_pairs", C.c_int64), ("total_occurrences", C.c_int64),
                ("row_off", C.c_void_p), ("pairs", C.c_void_p), ("left", C.c_void_p), ("right", C.c_void_p),
                ("stats32", C.c_void_p), ("row_flags", C.c_void_p), ("_owner", C.c_void_p)]


class Spans(C.Structure):
    _fields_ = [("ntext", C.c_int64), ("total_spans", C.c_int64), ("span_off", C.POINTER(C.c_int64)),
                ("spans", C.POINTER(C.c_int64)), ("_owner", C.c_void_p)]


class DeviceSpans(C.Structure):
    _fields_ = [("ntext", C.c_int64), ("total_spans", C.c_int64), ("span_off", C.c_void_p), ("spans", C.c_void_p),
                ("_owner", C.c_void_p)]


class FilterKey(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("index", C.c_void_p)]


class FilterBatch(C.Structure):
    _fields_ = [("keys", C.POINTER(FilterKey)), ("nkeys", C.c_int32), ("reserved", C.c_int32), ("kw", C.c_void_p),
                ("kw_len", C.c_int64), ("ranges", C.c_void_p), ("nranges", C.c_int64), ("terms", C.c_void_p),
                ("req_term_off", C.c_void_p), ("nreq", C.c_int64), ("corr_range", C.c_void_p), ("span", C.c_void_p)]


class FilterResult(C.Structure):
    _fields_ = [("nreq", C.c_int64), ("total_pairs", C.c_int64), ("row_off", C.POINTER(C.c_int64)),
                ("pairs", C.POINTER(C.c_int64)), ("matched", C.POINTER(C.c_int64)), ("_owner", C.c_void_p)]


# cdb_rows_ready_fn: (user, stats32 device pointer, npat, stream)
ROWS_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)

# cdb_filter_term (include/coffeedb_b200.h)
TERM_DTYPE = np.dtype([("key", "<i4"), ("range", "<i4"), ("kw_begin", "<i8"), ("kw_end", "<i8")])


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-j4"] + (["-B"] if force else [])
    r = subprocess.run(args, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcoffeedb_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def lib():
    """Loads the C-ABI library.  Fails loudly when it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(coffeedb_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
        L.cdb_last_error.restype = C.c_char_p
        L.cdb_version.restype = C.c_char_p
        L.cdb_device_count.restype = C.c_int
        L.cdb_create.argtypes = [C.POINTER(Options), C.POINTER(vp)]
        L.cdb_destroy.argtypes = [vp]
        L.cdb_destroy.restype = None
        L.cdb_add.argtypes = [vp, C.c_int64, vp, C.c_int64]
        L.cdb_add_many.argtypes = [vp, vp, vp, vp, C.c_int64]
        L.cdb_staging_stats.argtypes = [vp, i64p, i64p]
        L.cdb_build.argtypes = [vp]
        L.cdb_save.argtypes = [vp, C.c_char_p]
        L.cdb_build_or_load.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32)]
        L.cdb_build_device.argtypes = [vp, vp, vp, vp, C.c_int64, vp]
        L.cdb_info.argtypes = [vp, i64p, i64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.cdb_prefix_directory.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i64p]
        L.cdb_listing_info.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), i64p, C.POINTER(C.c_double)]
        L.cdb_export_sa.argtypes = [vp, vp, C.c_int64]
        L.cdb_sa_device_ptr.argtypes = [vp, C.POINTER(vp)]
        L.cdb_locate_batch.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(Result)]
        L.cdb_result_free.argtypes = [C.POINTER(Result)]
        L.cdb_query.argtypes = [vp, vp, C.c_int64, C.POINTER(Result)]
        L.cdb_query_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.cdb_result_free.restype = None
        L.cdb_locate_batch_device.argtypes = [vp, vp, vp, C.c_int64, vp, C.POINTER(DeviceResult)]
        L.cdb_locate_batch_device_ex.argtypes = [vp, vp, vp, C.c_int64, vp, ROWS_READY_FN, vp, C.POINTER(DeviceResult)]
        L.cdb_device_result_free.argtypes = [C.POINTER(DeviceResult)]
        L.cdb_device_result_free.restype = None
        L.cdb_locate_spans.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, C.POINTER(Spans)]
        L.cdb_locate_spans_batch.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, C.POINTER(Spans)]
        L.cdb_locate_spans_batch_device.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp,
                                                    C.POINTER(DeviceSpans)]
        L.cdb_device_spans_free.argtypes = [C.POINTER(DeviceSpans)]
        L.cdb_device_spans_free.restype = None
        L.cdb_spans_free.argtypes = [C.POINTER(Spans)]
        L.cdb_spans_free.restype = None
        L.cdb_splice.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64, vp, C.c_int64]
        L.cdb_splice.restype = C.c_int64
        L.cdb_verify_sa.argtypes = [vp, i64p]
        L.cdb_compare_sa.argtypes = [vp, vp, C.c_int64, i64p]
        L.cdb_build_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), i64p, i64p]
        L.cdb_last_locate_stats.argtypes = [C.POINTER(C.c_double), i64p]
        L.cdb_last_locate_stats.restype = None
        L.cdb_launch_count.restype = C.c_uint64
        L.cdb_trim.restype = None
        L.cdb_numeric_create.argtypes = [C.c_int32, vp, vp, C.c_int64, C.c_int32, C.POINTER(vp)]
        L.cdb_numeric_destroy.argtypes = [vp]
        L.cdb_numeric_destroy.restype = None
        L.cdb_numeric_query.argtypes = [vp, i64p, i64p, C.POINTER(Result)]
        L.cdb_filter.argtypes = [C.POINTER(FilterBatch), C.POINTER(FilterResult)]
        L.cdb_filter_result_free.argtypes = [C.POINTER(FilterResult)]
        L.cdb_filter_result_free.restype = None
        L.cdb_sharded_create.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(Options), C.POINTER(vp)]
        L.cdb_sharded_destroy.argtypes = [vp]
        L.cdb_sharded_destroy.restype = None
        L.cdb_sharded_add.argtypes = [vp, C.c_int64, vp, C.c_int64]
        L.cdb_sharded_add_many.argtypes = [vp, vp, vp, vp, C.c_int64]
        L.cdb_sharded_build.argtypes = [vp]
        L.cdb_sharded_locate_batch.argtypes = [vp, vp, vp, C.c_int64, C.POINTER(Result)]
        L.cdb_sharded_shard.argtypes = [vp, C.c_int32, C.POINTER(vp), i64p, i64p]
        L.cdb_sharded_count.argtypes = [vp]
        L.cdb_sharded_count.restype = C.c_int32
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().cdb_launch_count())


def last_locate_stats() -> dict:
    ms = (C.c_double * 8)()
    cn = (C.c_int64 * 8)()
    lib().cdb_last_locate_stats_ex(ms, cn)
    return {"search_ms": ms[0], "gather_ms": ms[1], "large_ms": ms[2], "tail_ms": ms[3], "translate_ms": ms[4],
            "total_ms": ms[5], "listing_ms": ms[6], "npat": cn[0], "pairs": cn[1], "occurrences": cn[2], "nlarge": cn[3],
            "nlisted": cn[4], "listed_pairs": cn[5]}


def _check(rc: int):
    if rc != CDB_OK:
        raise RuntimeError(lib().cdb_last_error().decode())


def _u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8)


def pack(items) -> tuple[np.ndarray, np.ndarray]:
    """list of bytes -> (bytes uint8[total], off int64[len+1])"""
    off = np.zeros(len(items) + 1, np.int64)
    if len(items):
        off[1:] = np.cumsum([len(x) for x in items])
    data = np.frombuffer(b"".join(bytes(x) for x in items), dtype=np.uint8) if off[-1] else np.zeros(0, np.uint8)
    return data, off


class StringIndex:
    """Drop-in for the reference's ``string_index`` (src/index.h:54-86)."""

    def __init__(self, device: int = -1, compat_signed: bool = True, workspace_bytes: int = 0,
                 keep_host_copy: bool = False):
        self._L = lib()
        self._h = C.c_void_p()
        opt = Options(device, 1 if compat_signed else 0, workspace_bytes, 1 if keep_host_copy else 0, 0)
        _check(self._L.cdb_create(C.byref(opt), C.byref(self._h)))
        self._keep = []  # device tensors borrowed by build_device

    # -- reference surface ----------------------------------------------------------------------------
    def add(self, id_: int, value: bytes):
        """string_index::add (src/index.cpp:174-177)"""
        v = _u8(value)
        _check(self._L.cdb_add(self._h, int(id_), v.ctypes.data if len(v) else None, len(v)))

    def add_many(self, ids, text, doc_off):
        ids = np.ascontiguousarray(ids, np.int64)
        text = _u8(text)
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        _check(self._L.cdb_add_many(self._h, ids.ctypes.data, text.ctypes.data if len(text) else None,
                                    doc_off.ctypes.data, len(ids)))

    def staging_stats(self) -> dict:
        """Bytes of text staged by add() and how many of them already have a device copy."""
        a, b = C.c_int64(0), C.c_int64(0)
        _check(self._L.cdb_staging_stats(self._h, C.byref(a), C.byref(b)))
        return {"staged": a.value, "on_device": b.value}

    def build(self):
        """string_index::build (src/index.cpp:178-236)"""
        _check(self._L.cdb_build(self._h))

    def save(self, path: str):
        """Writes the built suffix array to `path`, keyed by a hash of the corpus."""
        _check(self._L.cdb_save(self._h, os.fsencode(path)))

    def build_or_load(self, path: str) -> bool:
        """build(), reading the suffix array back from `path` when the file matches the staged corpus.  -> loaded?"""
        loaded = C.c_int32(0)
        _check(self._L.cdb_build_or_load(self._h, os.fsencode(path), C.byref(loaded)))
        return bool(loaded.value)

    def query(self, keyword: bytes) -> list[tuple[int, int]]:
        """string_index::query (src/index.cpp:237-326): [(id, count)] in ascending doc index.  One keyword through
        cdb_query: calls from concurrent threads (ctypes drops the GIL) share device batches."""
        return [(int(a), int(b)) for a, b in self.query_array(keyword)]

    def query_array(self, keyword: bytes) -> np.ndarray:
        """query() as an int64 [d, 2] array."""
        kw = bytes(keyword)
        res = Result()
        _check(self._L.cdb_query(self._h, kw, len(kw), C.byref(res)))
        try:
            tp = res.total_pairs
            return np.ctypeslib.as_array(res.pairs, shape=(tp, 2)).copy() if tp else np.zeros((0, 2), np.int64)
        finally:
            self._L.cdb_result_free(C.byref(res))

    def query_stats(self) -> dict:
        """Coalescing counters of query(): keywords submitted, device batches issued, largest batch."""
        q, b, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._L.cdb_query_stats(self._h, C.byref(q), C.byref(b), C.byref(m)))
        return {"queries": q.value, "batches": b.value, "largest": m.value}

    # -- batched / device forms ------------------------------------------------------------------------
    def locate_batch(self, patterns, pat_off=None):
        """patterns: list of bytes, or (uint8 array, int64 offsets).  -> (row_off int64[npat+1], pairs int64[total,2])"""
        if pat_off is None:
            pat, pat_off = pack(list(patterns))
        else:
            pat, pat_off = _u8(patterns), np.ascontiguousarray(pat_off, np.int64)
        npat = len(pat_off) - 1
        res = Result()
        _check(self._L.cdb_locate_batch(self._h, pat.ctypes.data if len(pat) else None, pat_off.ctypes.data, npat,
                                        C.byref(res)))
        try:
            row_off = np.ctypeslib.as_array(res.row_off, shape=(npat + 1,)).copy()
            tp = res.total_pairs
            pairs = (np.ctypeslib.as_array(res.pairs, shape=(tp, 2)).copy() if tp else np.zeros((0, 2), np.int64))
            self.last_total_occurrences = res.total_occurrences
        finally:
            self._L.cdb_result_free(C.byref(res))
        return row_off, pairs

    def locate_batch_raw(self, pat: np.ndarray, pat_off: np.ndarray) -> Result:
        """Host-buffer call without copying the result out (caller must result_free)."""
        res = Result()
        _check(self._L.cdb_locate_batch(self._h, pat.ctypes.data, pat_off.ctypes.data, len(pat_off) - 1, C.byref(res)))
        return res

    def result_free(self, res):
        self._L.cdb_result_free(C.byref(res))

    def build_device(self, d_text_ptr: int, d_doc_off_ptr: int, d_ids_ptr: int, nd: int, stream: int = 0, keep=()):
        """Build from a corpus resident in device memory (pointers are borrowed; `keep` holds their owners).
        The text buffer must be padded with at least 64 readable bytes after the last document."""
        self._keep = list(keep)
        _check(self._L.cdb_build_device(self._h, d_text_ptr, d_doc_off_ptr, d_ids_ptr, nd, stream))

    def locate_batch_device(self, d_pat_ptr: int, d_pat_off_ptr: int, npat: int, stream: int = 0) -> DeviceResult:
        res = DeviceResult()
        _check(self._L.cdb_locate_batch_device(self._h, d_pat_ptr, d_pat_off_ptr, npat, stream, C.byref(res)))
        return res

    def locate_batch_device_ex(self, d_pat_ptr: int, d_pat_off_ptr: int, npat: int, stream: int, rows_ready) -> DeviceResult:
        """locate_batch_device with the sharded caller's hook: rows_ready(stats32_ptr, npat, stream_handle) runs once the
        per-pattern (row length, occurrences) are enqueued; they are ready on `stream_handle` (the launching stream, or a
        side stream of the library when the rows are being streamed from the document listing at that moment)."""
        res = DeviceResult()
        cb = ROWS_READY_FN(lambda _user, stats, n, st: rows_ready(stats, n, st))
        _check(self._L.cdb_locate_batch_device_ex(self._h, d_pat_ptr, d_pat_off_ptr, npat, stream, cb, None, C.byref(res)))
        return res

    def device_result_free(self, res: DeviceResult):
        self._L.cdb_device_result_free(C.byref(res))

    def info(self) -> dict:
        n, nd, w, bits, mask = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32(), C.c_uint64()
        _check(self._L.cdb_info(self._h, C.byref(n), C.byref(nd), C.byref(w), C.byref(bits), C.byref(mask)))
        return {"n": n.value, "nd": nd.value, "width": w.value, "bits": bits.value, "mask": mask.value}

    def prefix_directory(self) -> dict:
        k, b, e = C.c_int32(), C.c_int32(), C.c_int64()
        _check(self._L.cdb_prefix_directory(self._h, C.byref(k), C.byref(b), C.byref(e)))
        return {"symbols": k.value, "bits_per_symbol": b.value, "entries": e.value}

    def listing_info(self, order: int = 0) -> dict:
        """The document listing of the directory's buckets (order 0: doc order, 1: id order), see cdb_listing_info."""
        p, hw, b, ms = C.c_int32(), C.c_int32(), C.c_int64(), C.c_double()
        _check(self._L.cdb_listing_info(self._h, order, C.byref(p), C.byref(hw), C.byref(b), C.byref(ms)))
        return {"present": bool(p.value), "hi_bytes": hw.value, "bytes": b.value, "build_ms": ms.value}

    def build_stats(self) -> dict:
        t, s, r, c = C.c_double(), C.c_double(), C.c_int64(), C.c_int64()
        _check(self._L.cdb_build_stats(self._h, C.byref(t), C.byref(s), C.byref(r), C.byref(c)))
        return {"total_ms": t.value, "sort_ms": s.value, "rounds": r.value, "chunks": c.value}

    def verify_sa(self) -> dict:
        """Independent device-side check of the suffix array (adjacent-pair order under the reference's comparator,
        src/index.cpp:92-93 / note N1; permutation; ties).  `ok` is True iff nothing is out of place."""
        out = (C.c_int64 * 8)()
        _check(self._L.cdb_verify_sa(self._h, out))
        keys = ["inversions", "invalid", "duplicates", "ties", "ties_unordered", "signed_queue", "unchecked", "signed_rule_pairs"]
        d = {k: int(out[i]) for i, k in enumerate(keys)}
        d["ok"] = not (d["inversions"] or d["invalid"] or d["duplicates"] or d["ties_unordered"] or d["unchecked"])
        return d

    def compare_sa(self, other: np.ndarray) -> dict:
        """Element-wise comparison with another packed suffix array of the same corpus (uint32/uint64 of the index's
        width, host memory): identical elements, ties (different element, byte-identical suffix), real differences."""
        other = np.ascontiguousarray(other)
        out = (C.c_int64 * 3)()
        _check(self._L.cdb_compare_sa(self._h, other.ctypes.data, other.nbytes, out))
        return {"identical": int(out[0]), "ties": int(out[1]), "different": int(out[2])}

    def export_sa(self) -> np.ndarray:
        """The packed suffix array widened to uint64 (element = (offset << bits) | doc, src/index.cpp:209-215)."""
        inf = self.info()
        buf = np.zeros(max(inf["n"], 1), np.uint32 if inf["width"] == 4 else np.uint64)
        _check(self._L.cdb_export_sa(self._h, buf.ctypes.data, buf.nbytes))
        return buf[: inf["n"]].astype(np.uint64)

    def sa_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(self._L.cdb_sa_device_ptr(self._h, C.byref(p)))
        return p.value or 0

    def spans(self, keywords, docs) -> list[np.ndarray]:
        """Merged highlight spans (inclusive [begin,end]) of `keywords` inside each doc index in `docs`
        (replaces the span loop of ac_automaton::render, src/database.cpp:58-77)."""
        kw, kw_off = pack(list(keywords))
        docs = np.ascontiguousarray(docs, np.int64)
        sp = Spans()
        _check(self._L.cdb_locate_spans(self._h, kw.ctypes.data if len(kw) else None, kw_off.ctypes.data, len(kw_off) - 1,
                                        docs.ctypes.data if len(docs) else None, len(docs), C.byref(sp)))
        try:
            off = np.ctypeslib.as_array(sp.span_off, shape=(len(docs) + 1,)).copy() if len(docs) else np.zeros(1, np.int64)
            allsp = (np.ctypeslib.as_array(sp.spans, shape=(sp.total_spans, 2)).copy() if sp.total_spans
                     else np.zeros((0, 2), np.int64))
        finally:
            self._L.cdb_spans_free(C.byref(sp))
        return [allsp[off[i]:off[i + 1]] for i in range(len(docs))]

    def spans_batch(self, requests, texts) -> list[np.ndarray]:
        """Batched highlight.  requests: list of keyword lists; texts: list of (request index, doc index).  -> one int64
        [k, 2] array of merged inclusive [begin, end] spans per text (cdb_locate_spans_batch)."""
        flat = [k for r in requests for k in r]
        kw, kw_off = pack(flat)
        rko = np.zeros(len(requests) + 1, np.int64)
        if len(requests):
            rko[1:] = np.cumsum([len(r) for r in requests])
        treq = np.ascontiguousarray([t[0] for t in texts], np.int64)
        tdoc = np.ascontiguousarray([t[1] for t in texts], np.int64)
        sp = Spans()
        _check(self._L.cdb_locate_spans_batch(self._h, kw.ctypes.data if len(kw) else None, kw_off.ctypes.data, len(flat),
                                              rko.ctypes.data, len(requests), treq.ctypes.data if len(texts) else None,
                                              tdoc.ctypes.data if len(texts) else None, len(texts), C.byref(sp)))
        try:
            off = np.ctypeslib.as_array(sp.span_off, shape=(len(texts) + 1,)).copy() if len(texts) else np.zeros(1, np.int64)
            allsp = (np.ctypeslib.as_array(sp.spans, shape=(sp.total_spans, 2)).copy() if sp.total_spans
                     else np.zeros((0, 2), np.int64))
        finally:
            self._L.cdb_spans_free(C.byref(sp))
        return [allsp[off[i]:off[i + 1]] for i in range(len(texts))]

    def spans_batch_device(self, d_kw, d_kw_off, nkw, d_req_kw_off, nreq, d_text_req, d_text_doc, ntext, stream=0) -> DeviceSpans:
        """Device pointers in, device CSR out (caller frees with device_spans_free)."""
        out = DeviceSpans()
        _check(self._L.cdb_locate_spans_batch_device(self._h, d_kw, d_kw_off, nkw, d_req_kw_off, nreq, d_text_req, d_text_doc,
                                                     ntext, stream, C.byref(out)))
        return out

    def device_spans_free(self, sp: DeviceSpans):
        self._L.cdb_device_spans_free(C.byref(sp))

    def close(self):
        if self._h:
            if not getattr(self, "_borrowed", False):
                self._L.cdb_destroy(self._h)
            self._h = C.c_void_p()
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def splice(text: bytes, spans: np.ndarray, left: bytes, right: bytes) -> bytes:
    """Marker splicing (src/database.cpp:78-90) — host-side, C-ABI cdb_splice."""
    L = lib()
    t, l, r = _u8(text), _u8(left), _u8(right)
    sp = np.ascontiguousarray(spans, np.int64).reshape(-1)
    ns = len(sp) // 2
    need = L.cdb_splice(t.ctypes.data if len(t) else None, len(t), sp.ctypes.data if ns else None, ns,
                        l.ctypes.data if len(l) else None, len(l), r.ctypes.data if len(r) else None, len(r), None, 0)
    out = np.zeros(max(need, 1), np.uint8)
    got = L.cdb_splice(t.ctypes.data if len(t) else None, len(t), sp.ctypes.data if ns else None, ns,
                       l.ctypes.data if len(l) else None, len(l), r.ctypes.data if len(r) else None, len(r),
                       out.ctypes.data, need)
    return out[:got].tobytes()


# ---- filter() on the device (include/coffeedb_b200.h: cdb_filter) --------------------------------------
INT64_MAX = (1 << 63) - 1
_RANGE = re.compile(r"\s*(\[|\()\s*(.+)\s*,\s*(.+)(\]|\))\s*")  # range_pattern, src/utility.h:68


def parse_range(text: str, kind: int) -> tuple[int, int, int, int]:
    """parse_range<T> (src/utility.h:69-86) -> (lo value bits, lo id, hi value bits, hi id).  kind 0: int64, 1: double.
    "-inf" / "inf" become numeric_limits<T>::min() / max() as value_conv does (for double min() is the smallest
    positive normal — the reference's behaviour, kept)."""
    m = _RANGE.fullmatch(text)
    if not m:
        raise ValueError("Invalid range: " + text)

    def conv(sv: str) -> int:
        low = sv.lower()
        if kind == 0:
            if low == "-inf":
                return -(1 << 63)
            if low == "inf":
                return INT64_MAX
            if not re.fullmatch(r"-?[0-9]+", sv):
                raise ValueError("Invalid value: " + low)
            return int(sv)
        if low == "-inf":
            v = np.finfo(np.float64).tiny
        elif low == "inf":
            v = np.finfo(np.float64).max
        else:
            v = float(sv)
        return int(np.array([v], np.float64).view(np.int64)[0])

    return (conv(m.group(2)), INT64_MAX if m.group(1) == "(" else 0, conv(m.group(3)), INT64_MAX if m.group(4) == "]" else 0)


def parse_uint_range(text: str) -> tuple[int, int]:
    """parse_uint_range (src/utility.h:87-104): "$correlation" and "span" -> half-open [L, R)."""
    m = _RANGE.fullmatch(text)
    lo, hi = 1, 0
    if m:
        lo, hi = int(m.group(2)), int(m.group(3))
        if m.group(1) == "(":
            lo += 1
        if m.group(4) == "]":
            hi += 1
    if lo > hi or lo < 0:
        raise ValueError("Invalid range: " + text)
    return lo, hi


class NumericIndex:
    """integer_index / double_index (src/index.h:29-53) on the device: kind 0 = int64 values, 1 = double."""

    def __init__(self, kind: int, ids, values, device: int = -1):
        self._L = lib()
        self.kind = kind
        ids = np.ascontiguousarray(ids, np.int64)
        vals = np.ascontiguousarray(values, np.int64 if kind == 0 else np.float64)
        assert len(ids) == len(vals)
        self._h = C.c_void_p()
        _check(self._L.cdb_numeric_create(kind, ids.ctypes.data, vals.ctypes.data, len(ids), device, C.byref(self._h)))

    def query(self, range_text: str) -> np.ndarray:
        """integer_index::query / double_index::query (src/index.cpp:159-161, 170-172): [(id, 0)] in (value, id) order."""
        lo0, lo1, hi0, hi1 = parse_range(range_text, self.kind)
        lo = (C.c_int64 * 2)(lo0, lo1)
        hi = (C.c_int64 * 2)(hi0, hi1)
        res = Result()
        _check(self._L.cdb_numeric_query(self._h, lo, hi, C.byref(res)))
        try:
            return np.ctypeslib.as_array(res.pairs, shape=(max(res.total_pairs, 1), 2))[: res.total_pairs].copy()
        finally:
            self._L.cdb_result_free(C.byref(res))

    def close(self):
        if self._h:
            self._L.cdb_numeric_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def filter_raw(keys, kw: np.ndarray, ranges: np.ndarray, terms: np.ndarray, req_term_off: np.ndarray, corr=None, span=None):
    """cdb_filter on packed arrays (no per-request Python work): keys = list of StringIndex | NumericIndex | None (a key the
    database does not have); terms = TERM_DTYPE array; ranges = int64 [nranges, 4]; corr / span = int64 [nreq, 2] or None.
    -> FilterResult (free with filter_result_free); row_off / pairs / matched are views of pinned host memory."""
    L = lib()
    fk = (FilterKey * max(len(keys), 1))()
    for i, k in enumerate(keys):
        fk[i].kind = -1 if k is None else (0 if isinstance(k, StringIndex) else 1)
        fk[i].index = None if k is None else k._h
    b = FilterBatch()
    b.keys, b.nkeys = fk, len(keys)
    b.kw, b.kw_len = (kw.ctypes.data if len(kw) else None), len(kw)
    b.ranges, b.nranges = (ranges.ctypes.data if len(ranges) else None), len(ranges)
    assert terms.dtype == TERM_DTYPE and req_term_off.dtype == np.int64
    b.terms = terms.ctypes.data if len(terms) else None
    b.req_term_off, b.nreq = req_term_off.ctypes.data, len(req_term_off) - 1
    b.corr_range = corr.ctypes.data if corr is not None else None
    b.span = span.ctypes.data if span is not None else None
    res = FilterResult()
    _check(L.cdb_filter(C.byref(b), C.byref(res)))
    return res


def filter_result_free(res: FilterResult):
    lib().cdb_filter_result_free(C.byref(res))


def filter_batch(keys: dict, requests: list) -> list:
    """filter() + span of src/interface.cpp:46-147, 196-209 for a batch of requests.

    keys      {name: StringIndex | NumericIndex | None}
    requests  [{"constraints": {name: "range" | ["range", ...], "$correlation": "[L,R)"}, "span": "[b,e)"}, ...]
    ->        per request (pairs int64 [m, 2] = (id, $correlation) in the reference's order, matched)"""
    names = list(keys)
    slot = {n: i for i, n in enumerate(names)}
    objs = [keys[n] for n in names]
    terms, ranges, rto, kws = [], [], [0], bytearray()
    corr = np.zeros((len(requests), 2), np.int64)
    span = np.zeros((len(requests), 2), np.int64)
    corr[:, 0], corr[:, 1] = -(1 << 63), INT64_MAX
    span[:, 1] = INT64_MAX
    for r, req in enumerate(requests):
        for name, val in req.get("constraints", {}).items():
            if name == "$correlation":
                corr[r] = parse_uint_range(val)
                continue
            vals = [val] if isinstance(val, str) else list(val)
            if name not in slot:  # a key the database does not have (src/database.cpp:389-391)
                slot[name] = len(objs)
                objs.append(None)
            k = slot[name]
            for v in vals:
                if isinstance(objs[k], NumericIndex):
                    ranges.append(parse_range(v, objs[k].kind))
                    terms.append((k, len(ranges) - 1, 0, 0))
                else:
                    b = v.encode() if isinstance(v, str) else bytes(v)
                    terms.append((k, -1, len(kws), len(kws) + len(b)))
                    kws += b
        rto.append(len(terms))
        if req.get("span") is not None:
            span[r] = parse_uint_range(req["span"])
    res = filter_raw(objs, np.frombuffer(bytes(kws), np.uint8), np.array(ranges, np.int64).reshape(-1, 4),
                     np.array(terms, TERM_DTYPE), np.array(rto, np.int64), corr, span)
    try:
        n = res.nreq
        ro = np.ctypeslib.as_array(res.row_off, shape=(n + 1,)).copy()
        pr = np.ctypeslib.as_array(res.pairs, shape=(max(res.total_pairs, 1), 2))[: res.total_pairs].copy()
        mt = np.ctypeslib.as_array(res.matched, shape=(max(n, 1),))[:n].copy()
    finally:
        filter_result_free(res)
    return [(pr[ro[r]:ro[r + 1]], int(mt[r])) for r in range(n)]


class MultiDeviceIndex:
    """One string_index over several GPUs of this process (cdb_sharded_*): same add / build / query
    surface as StringIndex; `devices` lists one CUDA ordinal per shard (an ordinal may repeat)."""

    def __init__(self, devices, compat_signed: bool = True, keep_host_copy: bool = False):
        self._L = lib()
        self._h = C.c_void_p()
        dv = (C.c_int32 * len(devices))(*devices)
        opt = Options(-1, 1 if compat_signed else 0, 0, 1 if keep_host_copy else 0, 0)
        _check(self._L.cdb_sharded_create(dv, len(devices), C.byref(opt), C.byref(self._h)))

    def add(self, id_: int, value: bytes):
        v = _u8(value)
        _check(self._L.cdb_sharded_add(self._h, id_, v.ctypes.data if len(v) else None, len(v)))

    def add_many(self, ids, text, doc_off):
        ids = np.ascontiguousarray(ids, np.int64)
        text = _u8(text)
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        _check(self._L.cdb_sharded_add_many(self._h, ids.ctypes.data, text.ctypes.data if len(text) else None, doc_off.ctypes.data,
                                            len(ids)))

    def build(self):
        _check(self._L.cdb_sharded_build(self._h))

    def locate_batch(self, patterns, pat_off=None):
        """-> (row_off int64 [npat+1], pairs int64 [total, 2]); row q = string_index::query(pattern q) on the whole corpus."""
        if pat_off is None:
            pat, pat_off = pack(list(patterns))
        else:
            pat, pat_off = _u8(patterns), np.ascontiguousarray(pat_off, np.int64)
        res = Result()
        _check(self._L.cdb_sharded_locate_batch(self._h, pat.ctypes.data if len(pat) else None, pat_off.ctypes.data, len(pat_off) - 1,
                                                C.byref(res)))
        try:
            n = res.npat
            ro = np.ctypeslib.as_array(res.row_off, shape=(n + 1,)).copy()
            pr = np.ctypeslib.as_array(res.pairs, shape=(max(res.total_pairs, 1), 2))[: res.total_pairs].copy()
        finally:
            self._L.cdb_result_free(C.byref(res))
        return ro, pr

    def query(self, keyword: bytes):
        ro, pr = self.locate_batch([keyword])
        return [(int(a), int(b)) for a, b in pr]

    def shards(self):
        """[(borrowed StringIndex view, doc_begin, doc_end)] — for parity checks of the individual shards."""
        out = []
        for g in range(self._L.cdb_sharded_count(self._h)):
            h, b, e = C.c_void_p(), C.c_int64(0), C.c_int64(0)
            _check(self._L.cdb_sharded_shard(self._h, g, C.byref(h), C.byref(b), C.byref(e)))
            view = StringIndex.__new__(StringIndex)
            view._L, view._h, view._keep, view._borrowed = self._L, h, [], True
            out.append((view, b.value, e.value))
        return out

    def close(self):
        if self._h:
            self._L.cdb_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
