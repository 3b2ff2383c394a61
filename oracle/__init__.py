"""TEST INFRASTRUCTURE ONLY — Python bindings of the parity oracle.

Two checkers live here, both CPU-only:

* ``port``  – the plain-C restatement (``coffee_oracle.c``), built on demand with gcc.
* ``Ref``   – the UNMODIFIED reference (``/root/reference/src/index.cpp`` + the highlighter of
  ``src/database.cpp``) compiled by ``oracle/Makefile`` into ``oracle/_ref/libcoffeeref.so``.
  It can only be (re)built where ``/root/reference`` exists; the built file travels to the GPU box.

Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and
``__graft_entry__.smoke()`` may import this package.  The product (``coffeedb_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libcoffeeoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcoffeeref.so")
REFERENCE_ROOT = os.environ.get("COFFEEDB_REFERENCE", "/root/reference")

ERR_MESSAGES = {
    1: "The amount of data exceeds the maximum range that CoffeeDB can handle",
    2: "The number of objects exceeds the maximum range that CoffeeDB can handle",
    3: "Empty keywords are not allowed",
}


def _newer(a: str, b: str) -> bool:
    return os.path.exists(a) and os.path.exists(b) and os.path.getmtime(a) > os.path.getmtime(b)


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "coffee_oracle.c")
    if force or not os.path.exists(PORT_SO) or _newer(src, PORT_SO):
        subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    return PORT_SO


def build_ref(force: bool = False) -> str | None:
    """Builds oracle/_ref when the reference sources are present; otherwise returns the prebuilt
    file if there is one, else None."""
    have_src = os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
    stale = any(_newer(os.path.join(HERE, f), REF_SO) for f in ("ref_harness.cpp", "ref_highlight.cpp"))
    if have_src and (force or stale or not os.path.exists(REF_SO)):
        subprocess.run(["make", "-s", "-C", HERE, "ref", f"REF={REFERENCE_ROOT}"], check=True)
    return REF_SO if os.path.exists(REF_SO) else None


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _bytes_arr(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if len(b) else np.zeros(0, np.uint8)


def pack_docs(docs) -> tuple[np.ndarray, np.ndarray]:
    """list of bytes -> (text uint8[n], doc_off int64[nd+1])."""
    off = np.zeros(len(docs) + 1, np.int64)
    if docs:
        off[1:] = np.cumsum([len(d) for d in docs])
    text = np.frombuffer(b"".join(docs), dtype=np.uint8).copy() if off[-1] else np.zeros(0, np.uint8)
    return text, off


def pack_patterns(pats) -> tuple[np.ndarray, np.ndarray]:
    return pack_docs(pats)


# --------------------------------------------------------------------------------------------------
class _Port:
    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            lib = C.CDLL(build_port())
            u8p, i64p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)
            lib.co_widths.argtypes = [i64p, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
            lib.co_widths.restype = C.c_int
            lib.co_build_sa.argtypes = [u8p, i64p, C.c_int64, u64p]
            lib.co_build_sa.restype = C.c_int
            lib.co_canonicalise_sa.argtypes = [u8p, i64p, C.c_int64, u64p, C.c_int64, C.c_int]
            lib.co_canonicalise_sa.restype = None
            lib.co_search.argtypes = [u8p, i64p, u64p, C.c_int64, C.c_int, u8p, C.c_int64, i64p, i64p]
            lib.co_search.restype = None
            lib.co_query.argtypes = [u8p, i64p, i64p, u64p, C.c_int64, C.c_int, u8p, C.c_int64, C.POINTER(i64p)]
            lib.co_query.restype = C.c_int64
            lib.co_spans.argtypes = [u8p, i64p, C.c_int64, u8p, C.c_int64, C.POINTER(i64p)]
            lib.co_spans.restype = C.c_int64
            lib.co_splice.argtypes = [u8p, C.c_int64, i64p, C.c_int64, u8p, C.c_int64, u8p, C.c_int64, C.POINTER(u8p)]
            lib.co_splice.restype = C.c_int64
            lib.co_free.argtypes = [C.c_void_p]
            lib.co_free.restype = None
            self._lib = lib
        return self._lib

    def widths(self, doc_off: np.ndarray):
        """-> (rc, bits1, bits2, width)"""
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        b1, b2, w = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = self.lib.co_widths(_p(doc_off, C.c_int64), len(doc_off) - 1, C.byref(b1), C.byref(b2), C.byref(w))
        return rc, b1.value, b2.value, w.value

    def build_sa(self, text: np.ndarray, doc_off: np.ndarray) -> tuple[np.ndarray, int, int]:
        """-> (sa uint64[n] canonical, bits1, width); raises RuntimeError with the reference's message."""
        text = _bytes_arr(text)
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        rc, b1, _b2, w = self.widths(doc_off)
        if rc:
            raise RuntimeError(ERR_MESSAGES[rc])
        n = int(doc_off[-1] - doc_off[0]) if len(doc_off) > 1 else 0
        sa = np.zeros(max(n, 1), np.uint64)
        textp = np.concatenate([text, np.zeros(1, np.uint8)])
        rc = self.lib.co_build_sa(_p(textp, C.c_uint8), _p(doc_off, C.c_int64), len(doc_off) - 1, _p(sa, C.c_uint64))
        if rc:
            raise RuntimeError(ERR_MESSAGES.get(rc, f"oracle error {rc}"))
        return sa[:n], b1, w

    def canonicalise(self, text, doc_off, sa: np.ndarray, bits1: int) -> np.ndarray:
        text = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)])
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        out = np.ascontiguousarray(sa, np.uint64).copy()
        self.lib.co_canonicalise_sa(_p(text, C.c_uint8), _p(doc_off, C.c_int64), len(doc_off) - 1,
                                    _p(out, C.c_uint64), len(out), bits1)
        return out

    def search(self, text, doc_off, sa, bits1: int, kw: bytes) -> tuple[int, int]:
        text = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)])
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        sa = np.ascontiguousarray(sa, np.uint64)
        k = np.concatenate([_bytes_arr(kw), np.zeros(1, np.uint8)])
        left, right = C.c_int64(0), C.c_int64(0)
        self.lib.co_search(_p(text, C.c_uint8), _p(doc_off, C.c_int64), _p(sa, C.c_uint64), len(sa), bits1,
                           _p(k, C.c_uint8), len(kw), C.byref(left), C.byref(right))
        return left.value, right.value

    def query(self, text, doc_off, ids, sa, bits1: int, kw: bytes) -> np.ndarray:
        """-> int64[npairs, 2] of (id, count) in ascending doc index; raises on empty keyword."""
        text = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)])
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        ids = np.ascontiguousarray(ids, np.int64)
        sa = np.ascontiguousarray(sa, np.uint64)
        k = np.concatenate([_bytes_arr(kw), np.zeros(1, np.uint8)])
        out = C.POINTER(C.c_int64)()
        npairs = self.lib.co_query(_p(text, C.c_uint8), _p(doc_off, C.c_int64), _p(ids, C.c_int64),
                                   _p(sa, C.c_uint64), len(sa), bits1, _p(k, C.c_uint8), len(kw), C.byref(out))
        if npairs < 0:
            raise RuntimeError(ERR_MESSAGES.get(-npairs, f"oracle error {npairs}"))
        res = np.ctypeslib.as_array(out, shape=(max(npairs, 1), 2))[:npairs].copy() if npairs else np.zeros((0, 2), np.int64)
        self.lib.co_free(out)
        return res

    def spans(self, keywords, text: bytes) -> np.ndarray:
        """-> int64[nspans, 2] inclusive [begin, end]."""
        kb, ko = pack_patterns(list(keywords))
        kb = np.concatenate([kb, np.zeros(1, np.uint8)])
        t = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)])
        out = C.POINTER(C.c_int64)()
        ns = self.lib.co_spans(_p(kb, C.c_uint8), _p(ko, C.c_int64), len(ko) - 1, _p(t, C.c_uint8), len(text), C.byref(out))
        if ns < 0:
            raise MemoryError("co_spans")
        res = np.ctypeslib.as_array(out, shape=(max(ns, 1), 2))[:ns].copy() if ns else np.zeros((0, 2), np.int64)
        self.lib.co_free(out)
        return res

    def splice(self, text: bytes, spans: np.ndarray, left: bytes, right: bytes) -> bytes:
        t = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)])
        sp = np.ascontiguousarray(spans, np.int64).reshape(-1)
        sp = np.concatenate([sp, np.zeros(2, np.int64)])
        l = np.concatenate([_bytes_arr(left), np.zeros(1, np.uint8)])
        r = np.concatenate([_bytes_arr(right), np.zeros(1, np.uint8)])
        out = C.POINTER(C.c_uint8)()
        w = self.lib.co_splice(_p(t, C.c_uint8), len(text), _p(sp, C.c_int64), (len(sp) - 2) // 2, _p(l, C.c_uint8),
                               len(left), _p(r, C.c_uint8), len(right), C.byref(out))
        res = bytes(bytearray(out[:w]))
        self.lib.co_free(out)
        return res


port = _Port()


# --------------------------------------------------------------------------------------------------
_ref_lib = None


def ref_available() -> bool:
    return build_ref() is not None


def _ref():
    global _ref_lib
    if _ref_lib is None:
        path = build_ref()
        if path is None:
            raise RuntimeError("oracle/_ref/libcoffeeref.so is not built and /root/reference is absent")
        lib = C.CDLL(path)
        i64p = C.POINTER(C.c_int64)
        lib.ref_create.restype = C.c_void_p
        lib.ref_destroy.argtypes = [C.c_void_p]
        lib.ref_add.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]
        lib.ref_add_many.argtypes = [C.c_void_p, i64p, C.c_void_p, i64p, C.c_int64]
        lib.ref_add_many_borrowed.argtypes = [C.c_void_p, i64p, C.c_void_p, i64p, C.c_int64]
        lib.ref_adopt_sa.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
        lib.ref_build.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.ref_build.restype = C.c_int
        lib.ref_export_sa.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64), C.c_void_p]
        lib.ref_query.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(i64p), C.c_char_p, C.c_int]
        lib.ref_query.restype = C.c_int64
        lib.ref_free.argtypes = [C.c_void_p]
        lib.ref_query_batch.argtypes = [C.c_void_p, C.c_void_p, i64p, C.c_int64, C.c_int, i64p, i64p]
        lib.ref_query_batch.restype = C.c_double
        lib.ref_query_batch_csr.argtypes = [C.c_void_p, C.c_void_p, i64p, C.c_int64, C.c_int, i64p, C.POINTER(i64p)]
        lib.ref_query_batch_csr.restype = C.c_int64
        lib.ref_filter_span_batch.argtypes = [C.c_void_p, C.c_void_p, i64p, C.c_int64, C.c_int, C.c_int64, C.c_int64, i64p,
                                              C.POINTER(i64p), i64p]
        lib.ref_filter_span_batch.restype = C.c_double
        lib.ref_hardware_threads.restype = C.c_int
        lib.ref_render.argtypes = [C.c_void_p, i64p, C.c_int64, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64,
                                   C.c_char_p, C.c_int64, C.POINTER(C.c_void_p)]
        lib.ref_render.restype = C.c_int64
        _ref_lib = lib
    return _ref_lib


class Ref:
    """The reference's own string_index (add / build / query), plus a peek at its private SA."""

    def __init__(self):
        self._lib = _ref()
        self._h = C.c_void_p(self._lib.ref_create())

    def close(self):
        if self._h:
            self._lib.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, id_: int, value: bytes):
        self._lib.ref_add(self._h, id_, value, len(value))

    def add_many(self, ids: np.ndarray, text: np.ndarray, doc_off: np.ndarray):
        ids = np.ascontiguousarray(ids, np.int64)
        text = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)]) if len(text) < (1 << 28) else _bytes_arr(text)
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        self._lib.ref_add_many(self._h, _p(ids, C.c_int64), text.ctypes.data, _p(doc_off, C.c_int64), len(ids))

    def add_many_borrowed(self, ids: np.ndarray, text: np.ndarray, doc_off: np.ndarray):
        """add() for every document without copying the text (the arrays are kept alive by this object)."""
        ids = np.ascontiguousarray(ids, np.int64)
        doc_off = np.ascontiguousarray(doc_off, np.int64)
        assert text.dtype == np.uint8 and text.flags.c_contiguous
        self._borrowed = (ids, text, doc_off)
        self._lib.ref_add_many_borrowed(self._h, _p(ids, C.c_int64), text.ctypes.data, _p(doc_off, C.c_int64), len(ids))

    def adopt_sa(self, sa: np.ndarray, bits1: int):
        """Makes query() run on a caller-supplied packed suffix array (uint32 or uint64, kept alive here)."""
        assert sa.dtype in (np.uint32, np.uint64) and sa.flags.c_contiguous
        self._sa = sa
        self._lib.ref_adopt_sa(self._h, sa.ctypes.data, sa.dtype.itemsize, bits1, len(sa))

    def build(self):
        err = C.create_string_buffer(512)
        if self._lib.ref_build(self._h, err, 512):
            raise RuntimeError(err.value.decode())

    def export_sa(self) -> tuple[np.ndarray, int, int, int]:
        """-> (raw sa widened to uint64, bits1, mask, width)"""
        w, bits, mask, size = C.c_int(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._lib.ref_export_sa(self._h, C.byref(w), C.byref(bits), C.byref(mask), C.byref(size), None)
        buf = np.zeros(max(size.value, 1), np.uint32 if w.value == 4 else np.uint64)
        self._lib.ref_export_sa(self._h, None, None, None, None, buf.ctypes.data)
        return buf[: size.value].astype(np.uint64), int(bits.value), int(mask.value), int(w.value)

    def export_sa_raw(self) -> tuple[np.ndarray, int, int, int]:
        """-> (raw sa in its own element width, bits1, mask, width) — no widening copy (for arrays of many GB)"""
        w, bits, mask, size = C.c_int(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._lib.ref_export_sa(self._h, C.byref(w), C.byref(bits), C.byref(mask), C.byref(size), None)
        buf = np.zeros(max(size.value, 1), np.uint32 if w.value == 4 else np.uint64)
        self._lib.ref_export_sa(self._h, None, None, None, None, buf.ctypes.data)
        return buf[: size.value], int(bits.value), int(mask.value), int(w.value)

    def query(self, kw: bytes) -> np.ndarray:
        out = C.POINTER(C.c_int64)()
        err = C.create_string_buffer(512)
        n = self._lib.ref_query(self._h, kw, len(kw), C.byref(out), err, 512)
        if n < 0:
            raise RuntimeError(err.value.decode())
        res = np.ctypeslib.as_array(out, shape=(max(n, 1), 2))[:n].copy() if n else np.zeros((0, 2), np.int64)
        self._lib.ref_free(out)
        return res

    def query_batch_timed(self, pat: np.ndarray, pat_off: np.ndarray, nthreads: int):
        """-> (seconds, total_pairs, total_occurrences)"""
        pat = np.concatenate([_bytes_arr(pat), np.zeros(1, np.uint8)])
        pat_off = np.ascontiguousarray(pat_off, np.int64)
        tp, to = C.c_int64(0), C.c_int64(0)
        s = self._lib.ref_query_batch(self._h, pat.ctypes.data, _p(pat_off, C.c_int64), len(pat_off) - 1, nthreads,
                                      C.byref(tp), C.byref(to))
        return s, tp.value, to.value

    def query_batch(self, pat: np.ndarray, pat_off: np.ndarray, nthreads: int = 0):
        """-> (row_off int64[npat+1], pairs int64[total,2])"""
        pat = np.concatenate([_bytes_arr(pat), np.zeros(1, np.uint8)])
        pat_off = np.ascontiguousarray(pat_off, np.int64)
        npat = len(pat_off) - 1
        row_off = np.zeros(npat + 1, np.int64)
        out = C.POINTER(C.c_int64)()
        nthreads = nthreads or hardware_threads()
        total = self._lib.ref_query_batch_csr(self._h, pat.ctypes.data, _p(pat_off, C.c_int64), npat, nthreads,
                                              _p(row_off, C.c_int64), C.byref(out))
        pairs = np.ctypeslib.as_array(out, shape=(max(total, 1), 2))[:total].copy() if total else np.zeros((0, 2), np.int64)
        self._lib.ref_free(out)
        return row_off, pairs


    def filter_span_batch(self, pat: np.ndarray, pat_off: np.ndarray, span0: int, span1: int, nthreads: int = 0, keep: bool = True):
        """One-keyword requests through query() + the sorts and the span of src/interface.cpp:82,143-146,196-209.
        -> (seconds, row_off, pairs[total,2], matched) — row_off / pairs / matched are None when keep is False."""
        pat = np.concatenate([_bytes_arr(pat), np.zeros(1, np.uint8)])
        pat_off = np.ascontiguousarray(pat_off, np.int64)
        npat = len(pat_off) - 1
        nthreads = nthreads or hardware_threads()
        if not keep:
            s = self._lib.ref_filter_span_batch(self._h, pat.ctypes.data, _p(pat_off, C.c_int64), npat, nthreads, span0, span1,
                                                None, None, None)
            return s, None, None, None
        row_off = np.zeros(npat + 1, np.int64)
        matched = np.zeros(max(npat, 1), np.int64)
        out = C.POINTER(C.c_int64)()
        s = self._lib.ref_filter_span_batch(self._h, pat.ctypes.data, _p(pat_off, C.c_int64), npat, nthreads, span0, span1,
                                            _p(row_off, C.c_int64), C.byref(out), _p(matched, C.c_int64))
        total = int(row_off[-1])
        pairs = np.ctypeslib.as_array(out, shape=(max(total, 1), 2))[:total].copy() if total else np.zeros((0, 2), np.int64)
        self._lib.ref_free(out)
        return s, row_off, pairs, matched[:npat]


def hardware_threads() -> int:
    return int(_ref().ref_hardware_threads())


def ref_render(keywords, text: bytes, left: bytes, right: bytes) -> bytes:
    """ac_automaton(keywords).render(text, left, right) of the reference (database.cpp:58-91)."""
    lib = _ref()
    kb, ko = pack_patterns(list(keywords))
    kb = np.concatenate([kb, np.zeros(1, np.uint8)])
    t = np.concatenate([_bytes_arr(text), np.zeros(1, np.uint8)])
    out = C.c_void_p()
    n = lib.ref_render(kb.ctypes.data, _p(ko, C.c_int64), len(ko) - 1, t.ctypes.data, len(text), left, len(left),
                       right, len(right), C.byref(out))
    res = C.string_at(out, n)
    lib.ref_free(out)
    return res
