"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C-ABI library
(coffeedb_b200 is a ctypes binding); expected values are the golden vectors generated from the compiled
reference, the live oracle, and size-independent properties at larger sizes."""
import hashlib
import os

import numpy as np
import pytest

import coffeedb_b200 as cdb
import oracle
from tests import cases, corpora

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint64).tobytes()).hexdigest()


def build(text, off, ids, **kw):
    ix = cdb.StringIndex(**kw)
    ix.add_many(ids, text, off)
    ix.build()
    return ix


ALL_CASES = list(cases.CASES)  # including the note-N1 cases (signed-radix / unsigned-leaf layout)


@pytest.mark.parametrize("name", ALL_CASES)
def test_suffix_array_matches_reference(name, golden):
    text, off, ids, _ = cases.CASES[name]()
    ix = build(text, off, ids)
    inf = ix.info()
    assert inf["bits"] == int(golden[f"{name}/bits1"]) and inf["width"] == int(golden[f"{name}/width"])
    assert inf["n"] == int(golden[f"{name}/n"]) and inf["mask"] == (1 << inf["bits"]) - 1
    sa = ix.export_sa()
    if f"{name}/sa" in golden:
        assert np.array_equal(sa, golden[f"{name}/sa"])
    assert sha(sa) == str(golden[f"{name}/sa_sha"])
    ix.close()


@pytest.mark.parametrize("small", ["0", "256"])
@pytest.mark.parametrize("name", ALL_CASES)
def test_locate_matches_reference(name, small, golden, monkeypatch):
    monkeypatch.setenv("CDB_SMALL_BATCH", small)  # the general path and the small-batch path
    text, off, ids, pats = cases.CASES[name]()
    ix = build(text, off, ids)
    row_off, pairs = ix.locate_batch(golden[f"{name}/pat"], golden[f"{name}/pat_off"])
    assert np.array_equal(row_off, golden[f"{name}/row_off"])
    assert np.array_equal(pairs, golden[f"{name}/pairs"])
    # single-keyword form == string_index::query
    for i in (0, len(pats) // 2, len(pats) - 1):
        g = golden[f"{name}/pairs"][golden[f"{name}/row_off"][i]:golden[f"{name}/row_off"][i + 1]]
        assert ix.query(pats[i]) == [tuple(r) for r in g.tolist()]
    ix.close()


def test_known_answers_and_errors():
    ix = cdb.StringIndex()
    ix.add(100, b"3010103")
    ix.add(101, b"301022")
    ix.build()
    assert ix.query(b"010") == [(100, 2), (101, 1)]      # README.md:91
    assert ix.query(b"0") == [(100, 3), (101, 2)]
    assert ix.query(b"9") == [] and ix.query(b"30101034") == []
    with pytest.raises(RuntimeError, match="^Empty keywords are not allowed$"):
        ix.query(b"")
    with pytest.raises(RuntimeError, match="^Empty keywords are not allowed$"):
        ix.locate_batch([b"a", b"", b"b"])
    ix.close()


def test_empty_index_and_empty_batch():
    ix = cdb.StringIndex()
    ix.build()
    assert ix.info()["n"] == 0 and ix.info()["bits"] == 1 and ix.info()["width"] == 4
    assert ix.query(b"a") == []
    ro, pr = ix.locate_batch([])
    assert ro.tolist() == [0] and pr.shape == (0, 2)
    ix.close()
    ix = cdb.StringIndex()
    for i in range(5):
        ix.add(i, b"")
    ix.build()
    assert ix.info()["n"] == 0 and ix.query(b"a") == []
    ix.close()


def test_against_live_oracle_random():
    """Random corpora the golden file does not hold: compare with the oracle port run here."""
    for seed in range(4):
        text, off, ids = corpora.ragged(2000 + 500 * seed, 40 + 30 * seed, seed=100 + seed,
                                        alphabet=[b"ab", b"abc", b"acgt", b"abcdefgh"][seed])
        ix = build(text, off, ids)
        sa, b1, w = oracle.port.build_sa(text, off)
        assert np.array_equal(ix.export_sa(), sa)
        pat, poff = corpora.sampled_patterns(text, off, 200, 1, 12, seed=200 + seed)
        row_off, pairs = ix.locate_batch(pat, poff)
        for q in range(200):
            kw = bytes(pat[poff[q]:poff[q + 1]])
            assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), kw
        ix.close()


def test_chunked_build_equals_single_chunk():
    """A workspace cap forces the multi-chunk path (the 10 GB configuration's path); result must not change."""
    text, off, ids = corpora.uniform(3000, 200, seed=41)
    a = build(text, off, ids)
    b = build(text, off, ids, workspace_bytes=3_000_000)
    assert b.build_stats()["chunks"] > 1 and a.build_stats()["chunks"] == 1
    assert np.array_equal(a.export_sa(), b.export_sa())
    a.close()
    b.close()


@pytest.mark.parametrize("limit", [None, "40000", "1"])
def test_large_interval_path(limit, monkeypatch):
    """Intervals longer than the warp path's capacity go through the device radix sort, in sub-batches of bounded
    size (CDB_LARGE_LIMIT occurrences; "1" = one pattern per sub-batch)."""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")  # these batches are small: keep them on the general path
    if limit:
        monkeypatch.setenv("CDB_LARGE_LIMIT", limit)
    text, off, ids = corpora.uniform(20000, 50, seed=43, lo=ord("a"), hi=ord("c"))
    ix = build(text, off, ids)
    sa, b1, _w = oracle.port.build_sa(text, off)
    pats = [b"a", b"ab", b"abc", b"abca", b"c", b"cc", b"abcabcabc", b"bbbbbbbbbbbbbbbbbbbbbbbb", b"ba", b"cab", b"bcb"]
    row_off, pairs = ix.locate_batch(pats)
    for q, kw in enumerate(pats):
        assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), kw
    ix.close()


def test_brute_force_property_medium():
    """test-string.py:52-56 property at a size the oracle does not need to see."""
    text, off, ids = corpora.uniform(2000, 2000, seed=47)
    ix = build(text, off, ids)
    pat, poff = corpora.uniform_patterns(40, 3, seed=48)
    row_off, pairs = ix.locate_batch(pat, poff)
    for q in range(40):
        kw = bytes(pat[poff[q]:poff[q + 1]])
        assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], corpora.brute_count(text, off, ids, kw))
    ix.close()


def test_highlight_spans_match_reference(golden):
    """K8: device spans == ac_automaton::render's spans (database.cpp:58-77)."""
    hl = cases.highlight_cases()
    rend, roff = golden["highlight/rendered"], golden["highlight/rendered_off"]
    texts = [t for _k, t in hl]
    ix = cdb.StringIndex()
    for i, t in enumerate(texts):
        ix.add(i, t)
    ix.build()
    for i, (kws, text) in enumerate(hl):
        spans = ix.spans(kws, [i])[0]
        assert np.array_equal(spans, oracle.port.spans(kws, text)), (kws, text)
        assert cdb.splice(text, spans, b"<b>", b"</b>") == rend[roff[i]:roff[i + 1]].tobytes()
    # several documents at once, duplicates and arbitrary order
    kws = [b"ab", b"b"]
    docs = [5, 3, 5, 0, 40]
    got = ix.spans(kws, docs)
    for d, sp in zip(docs, got):
        assert np.array_equal(sp, oracle.port.spans(kws, texts[d]))
    ix.close()


def test_translate_many_doc_ranges(monkeypatch):
    """translate_kernel walks ids[] by doc range (32 MB slices at production sizes); CDB_RANGE_BITS shrinks the
    ranges so that a small corpus exercises the multi-range path, including rows that miss most ranges."""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")  # these batches are small: keep them on the general path
    text, off, ids = corpora.uniform(20000, 60, seed=51, lo=ord("a"), hi=ord("e"))
    ix = build(text, off, ids)
    sa, b1, _w = oracle.port.build_sa(text, off)
    pat, poff = corpora.uniform_patterns(300, 4, seed=52, lo=ord("a"), hi=ord("e"))
    spat, soff = corpora.sampled_patterns(text, off, 100, 2, 9, seed=53)
    pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(300)] + [bytes(spat[soff[i]:soff[i + 1]]) for i in range(100)]
    pats += [b"zzzz", b"a", b"ab"]  # no hit / large-path rows mixed into the batch
    want = [oracle.port.query(text, off, ids, sa, b1, kw) for kw in pats]
    for bits in ("8", "11", "22"):
        monkeypatch.setenv("CDB_RANGE_BITS", bits)
        row_off, pairs = ix.locate_batch(pats)
        for q in range(len(pats)):
            assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], want[q]), (bits, pats[q])
    ix.close()


@pytest.mark.parametrize("buckets", ["1", "0"])
def test_gather_distribution_sort_and_fallback(buckets, monkeypatch):
    """gather_kernel sorts the documents of an interval with a distribution sort when they are spread over the corpus
    and with the sorting network when they cluster (or with CDB_GATHER_BUCKETS=0); both must give the reference's rows.
    Corpus A: 200 000 short documents, 6- and 7-byte patterns with ~730 / ~170 hits in unrelated documents (a few
    documents twice).  Corpus B: every hit falls into the first 300 of 100 300 documents, most of them repeatedly."""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")  # these batches are small: keep them on the general path
    monkeypatch.setenv("CDB_GATHER_BUCKETS", buckets)
    text, off, ids = corpora.uniform(200000, 20, seed=61, lo=ord("a"), hi=ord("d"))
    ix = build(text, off, ids)
    sa, b1, _w = oracle.port.build_sa(text, off)
    p6, o6 = corpora.uniform_patterns(150, 6, seed=62, lo=ord("a"), hi=ord("d"))
    p7, o7 = corpora.uniform_patterns(150, 7, seed=63, lo=ord("a"), hi=ord("d"))
    pats = [bytes(p6[o6[i]:o6[i + 1]]) for i in range(150)] + [bytes(p7[o7[i]:o7[i + 1]]) for i in range(150)]
    pats += [b"abcd", b"dddddddd", b"zz"]  # large path / few hits / none, mixed into the batch
    row_off, pairs = ix.locate_batch(pats)
    lens = np.diff(row_off)
    assert lens[:150].min() > 512 and lens[150:300].min() > 128  # the R = 32 and R = 8 code paths
    for q, kw in enumerate(pats):
        assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), kw
    ix.close()

    rng = np.random.default_rng(64)
    head = rng.integers(ord("a"), ord("d") + 1, size=300 * 2000, dtype=np.uint8)
    tail = rng.integers(ord("x"), ord("z") + 1, size=100000 * 8, dtype=np.uint8)
    text = np.concatenate([head, tail])
    off = np.concatenate([np.arange(300, dtype=np.int64) * 2000, 600000 + np.arange(100001, dtype=np.int64) * 8])
    ids = np.arange(100300, dtype=np.int64)[::-1].copy() + 7
    ix = build(text, off, ids)
    sa, b1, _w = oracle.port.build_sa(text, off)
    p5, o5 = corpora.uniform_patterns(200, 5, seed=65, lo=ord("a"), hi=ord("d"))
    pats = [bytes(p5[o5[i]:o5[i + 1]]) for i in range(200)] + [b"xyz", b"xyzxyz", b"ab"]
    row_off, pairs = ix.locate_batch(pats)
    for q, kw in enumerate(pats):
        assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), kw
    ix.close()


def test_query_coalesces_concurrent_callers():
    """cdb_query (what string_index::query maps to): one keyword per call, as the reference's server issues it from
    its worker pool (src/database.cpp:387-393); concurrent callers share device batches and each one gets exactly the
    row locate_batch gives for its keyword.  An empty keyword fails its own caller only, with the reference's text."""
    import threading

    text, off, ids = corpora.uniform(3000, 120, seed=71, lo=ord("a"), hi=ord("f"))
    ix = build(text, off, ids)
    pat, poff = corpora.uniform_patterns(240, 3, seed=72, lo=ord("a"), hi=ord("f"))
    pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(240)] + [b"zz", b"a"]  # no hit / large-interval path
    row_off, pairs = ix.locate_batch(pats)
    want = [pairs[row_off[q]:row_off[q + 1]] for q in range(len(pats))]
    assert np.array_equal(ix.query_array(pats[0]), want[0])  # a lone caller
    bad, errs = [], []

    def worker(w):
        try:
            for rep in range(3):
                for q in range(w, len(pats), 4):
                    if not np.array_equal(ix.query_array(pats[q]), want[q]):
                        bad.append((w, q))
                if w == 2:
                    with pytest.raises(RuntimeError, match="Empty keywords are not allowed"):
                        ix.query(b"")
        except Exception as e:  # noqa: BLE001 - reported below
            errs.append(repr(e))

    threads = [threading.Thread(target=worker, args=(w,)) for w in range(16)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs and not bad, (errs[:3], bad[:5])
    st = ix.query_stats()
    assert st["queries"] == 1 + 3 * sum(len(range(w, len(pats), 4)) for w in range(16))
    assert 1 <= st["batches"] <= st["queries"] and 1 <= st["largest"] <= 16
    ix.close()


def test_small_batch_path_matches_general_path(monkeypatch):
    """CDB_SMALL_BATCH (default 256; 0 = off): batches of up to 256 keywords take one upload, two launches and one synchronisation; the
    result must equal the general path's, and batches the path cannot take (a long interval, too many occurrences)
    must fall back silently."""
    text, off, ids = corpora.uniform(20000, 60, seed=81, lo=ord("a"), hi=ord("e"))
    ix = build(text, off, ids)
    pat, poff = corpora.sampled_patterns(text, off, 400, 5, 9, seed=82)  # <= ~370 occurrences each
    pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(400)]
    p5, o5 = corpora.uniform_patterns(250, 5, seed=83, lo=ord("a"), hi=ord("e"))
    many = [bytes(p5[o5[i]:o5[i + 1]]) for i in range(250)]  # 250 x ~370 occurrences: over the 65 536-pair capacity
    batches = [pats[:1], pats[1:8], pats[8:72], pats[72:328], [b"zzzz"], [b"zz", pats[0], b"q"],
               pats[:5] + [b"a"],  # b"a": ~240 000 occurrences, the long-interval path -> fallback
               many]
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    want = [ix.locate_batch(b) for b in batches]
    assert cdb.last_locate_stats()["total_ms"] > 0.0  # the general path ran
    monkeypatch.delenv("CDB_SMALL_BATCH")  # default: on
    for b, (wro, wpairs) in zip(batches, want):
        ro, pairs = ix.locate_batch(b)
        assert np.array_equal(ro, wro) and np.array_equal(pairs, wpairs), b[:3]
    took_small = []
    for b in batches:
        ix.locate_batch(b)
        took_small.append(cdb.last_locate_stats()["total_ms"] == 0.0)
    assert took_small[:6] == [True] * 6 and took_small[6:] == [False, False]
    with pytest.raises(RuntimeError, match="Empty keywords are not allowed"):
        ix.locate_batch([b"ab", b""])
    assert np.array_equal(ix.query_array(pats[3]), want[1][1][want[1][0][2]:want[1][0][3]])  # cdb_query on top of it
    ix.close()


def test_gather_mid_size_variants(monkeypatch):
    """CDB_GATHER_VARIANTS (default 1; 0 = only the 128- and 1024-key kernels): batches whose longest interval is <= 256 / <= 512 occurrences run gather_kernel compiled
    for 8 / 16 keys per lane (more warps per SM); rows must not change."""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")  # these batches are small: keep them on the general path
    text, off, ids = corpora.uniform(200000, 20, seed=91, lo=ord("a"), hi=ord("d"))
    ix = build(text, off, ids)
    for m in (7, 6):  # ~130-210 occurrences (the 256-key variant) and ~650-810 (the full-size kernel)
        p, o = corpora.uniform_patterns(200, m, seed=92 + m, lo=ord("a"), hi=ord("d"))
        pats = [bytes(p[o[i]:o[i + 1]]) for i in range(200)]
        monkeypatch.setenv("CDB_GATHER_VARIANTS", "0")
        want = ix.locate_batch(pats)
        monkeypatch.delenv("CDB_GATHER_VARIANTS")
        got = ix.locate_batch(pats)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), m
    # 512-key variant: half of the 6-byte patterns' documents (a corpus of 100 000 documents)
    text, off, ids = corpora.uniform(100000, 20, seed=95, lo=ord("a"), hi=ord("d"))
    ix2 = build(text, off, ids)
    p, o = corpora.uniform_patterns(200, 6, seed=96, lo=ord("a"), hi=ord("d"))
    pats = [bytes(p[o[i]:o[i + 1]]) for i in range(200)]
    monkeypatch.setenv("CDB_GATHER_VARIANTS", "0")
    want = ix2.locate_batch(pats)
    assert 256 < np.diff(want[0]).max() <= 512
    monkeypatch.delenv("CDB_GATHER_VARIANTS")
    got = ix2.locate_batch(pats)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    ix.close()
    ix2.close()


def test_n1_layout_against_live_oracle():
    """Note N1 at a size with several radix levels above chuck_size: the suffix array must equal the oracle's
    signed-radix / unsigned-leaf layout and every query must return the reference's (sometimes non-brute-force)
    answer, because the reference's own binary-search recurrences run on that layout."""
    for seed, (text, off, ids) in enumerate([corpora.utf8ish(3000, 300, seed=61),
                                             corpora.uniform(2000, 150, seed=62, lo=0x70, hi=0x90)]):
        ix = build(text, off, ids)
        sa, b1, _w = oracle.port.build_sa(text, off)
        assert np.array_equal(ix.export_sa(), sa)
        spat, soff = corpora.sampled_patterns(text, off, 300, 1, 6, seed=63 + seed)
        row_off, pairs = ix.locate_batch(spat, soff)
        nbrute = 0
        for q in range(300):
            kw = bytes(spat[soff[q]:soff[q + 1]])
            got = pairs[row_off[q]:row_off[q + 1]]
            assert np.array_equal(got, oracle.port.query(text, off, ids, sa, b1, kw)), kw
            nbrute += int(not np.array_equal(got, corpora.brute_count(text, off, ids, kw))) if q < 60 else 0
        ix.close()
        # compat off: plain unsigned order, prefix directory on, brute-force answers
        ix = build(text, off, ids, compat_signed=False)
        row_off, pairs = ix.locate_batch(spat, soff)
        for q in range(60):
            kw = bytes(spat[soff[q]:soff[q + 1]])
            assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], corpora.brute_count(text, off, ids, kw)), kw
        ix.close()


def test_prefix_directory_on_and_off(monkeypatch):
    """Same answers with the prefix directory disabled (reference recurrences) and at several directory depths."""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")  # these batches are small: keep them on the general path
    text, off, ids = corpora.ragged(4000, 80, seed=71, alphabet=b"abcdefg")
    sa, b1, _w = oracle.port.build_sa(text, off)
    spat, soff = corpora.sampled_patterns(text, off, 200, 1, 14, seed=72)
    pat, poff = corpora.uniform_patterns(100, 3, seed=73, lo=ord("a"), hi=ord("i"))  # some bytes absent from the corpus
    pats = [bytes(spat[soff[i]:soff[i + 1]]) for i in range(200)] + [bytes(pat[poff[i]:poff[i + 1]]) for i in range(100)]
    want = [oracle.port.query(text, off, ids, sa, b1, kw) for kw in pats]
    for bits in ("0", "3", "9", "16", "24"):
        monkeypatch.setenv("CDB_PTAB_BITS", bits)
        ix = build(text, off, ids)
        assert np.array_equal(ix.export_sa(), sa)
        row_off, pairs = ix.locate_batch(pats)
        for q in range(len(pats)):
            assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], want[q]), (bits, pats[q])
        ix.close()


def test_highlight_spans_single_request_entry():
    """cdb_locate_spans (one keyword set, several documents, duplicates and arbitrary order) == the reference
    highlighter's spans (database.cpp:58-77) and its rendered text."""
    text, off, ids = corpora.ragged(1500, 120, seed=81, alphabet=b"abc")
    ix = build(text, off, ids)
    kwsets = [[b"ab"], [b"a", b"bca"], [b"abc", b"cab", b"bb"], [b"c" * 5], [b"zz"], [b"abcabc", b"b"], [b"abcabcabca", b"ca"]]
    docs = [0, 7, 7, 1499, 3, 250, 1000]
    for kws in kwsets:
        got = ix.spans(kws, docs)
        for d, sp in zip(docs, got):
            t = text[off[d]:off[d + 1]].tobytes()
            assert np.array_equal(sp, oracle.port.spans(kws, t)), (kws, d)
            if oracle.ref_available():
                assert cdb.splice(t, sp, b"<", b">") == oracle.ref_render(kws, t, b"<", b">")
    with pytest.raises(RuntimeError, match="Empty keywords are not allowed"):
        ix.spans([b"a", b""], [0])
    with pytest.raises(RuntimeError, match="out of range"):
        ix.spans([b"a"], [1500])
    assert [len(x) for x in ix.spans([], [0, 1])] == [0, 0] and ix.spans([b"a"], []) == []
    ix.close()


def test_highlight_spans_batch_matches_reference():
    """cdb_locate_spans_batch: many requests, each with its own keyword set, each highlighted in its own documents, in one
    launch sequence — short documents (one thread each), documents longer than 1024 bytes (one warp each, spans that
    straddle the 32-position steps), keywords longer than the 8-byte compare window, overlapping / touching / nested
    occurrences.  Every text must equal ac_automaton::render's spans (src/database.cpp:58-77)."""
    rng = np.random.default_rng(91)
    docs_b = [bytes(rng.integers(ord("a"), ord("c") + 1, size=int(n), dtype=np.uint8))
              for n in list(rng.integers(0, 200, size=600)) + [1024, 1025, 3000, 20000, 70000]]
    docs_b += [b"ab" * 700, b"a" * 2100, b"abcabcabcabc" * 300, b""]
    text, off, ids = corpora.from_docs(docs_b, id_base=10)
    ix = build(text, off, ids)
    nd = len(docs_b)
    reqs = [[b"ab"], [b"a", b"bca"], [b"abc", b"cab", b"bb"], [b"ccccc"], [b"zz"], [b"abcabc", b"b"], [b"abcabcabca", b"ca"],
            [b"a"], [b"aa", b"aaa"], [b"abab", b"ba"], [b"abcabcabcabcabc"], [b"c", b"cc", b"ccc", b"bccb"]]
    texts = []
    for r in range(len(reqs)):
        for d in list(rng.integers(0, 600, size=40)) + list(range(600, nd)):
            texts.append((r, int(d)))
    rng.shuffle(texts)
    got = ix.spans_batch(reqs, texts)
    assert len(got) == len(texts)
    for (r, d), sp in zip(texts, got):
        assert np.array_equal(sp, oracle.port.spans(reqs[r], docs_b[d])), (reqs[r], d, len(docs_b[d]))
    # spot check against the reference's own renderer
    if oracle.ref_available():
        for (r, d), sp in list(zip(texts, got))[:60]:
            assert cdb.splice(docs_b[d], sp, b"<b>", b"</b>") == oracle.ref_render(reqs[r], docs_b[d], b"<b>", b"</b>")
    with pytest.raises(RuntimeError, match="out of range"):
        ix.spans_batch(reqs, [(0, nd)])
    with pytest.raises(RuntimeError, match="out of range"):
        ix.spans_batch(reqs, [(len(reqs), 0)])
    ix.close()


def test_concurrent_queries_and_rebuild():
    """Boundary threading contract (test/test-concurrency.py): query() is re-entrant — several host
    threads query one index at once while another index is being built (database.cpp builds the new index on locals
    while queries run against the previous one)."""
    import threading
    text, off, ids = corpora.uniform(4000, 120, seed=95, lo=ord("a"), hi=ord("f"))
    ix = build(text, off, ids)
    sa, b1, _w = oracle.port.build_sa(text, off)
    pat, poff = corpora.uniform_patterns(64, 3, seed=96, lo=ord("a"), hi=ord("f"))
    pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(64)]
    want = [oracle.port.query(text, off, ids, sa, b1, kw) for kw in pats]
    errors = []

    def querier(tid):
        try:
            for it in range(20):
                q = (tid * 7 + it) % len(pats)
                got = np.array(ix.query(pats[q]), np.int64).reshape(-1, 2)
                assert np.array_equal(got, want[q]), pats[q]
                ro, pr = ix.locate_batch(pats[q:q + 5])
                for j in range(len(ro) - 1):
                    assert np.array_equal(pr[ro[j]:ro[j + 1]], want[q + j])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    def builder():
        try:
            for seed in range(3):
                t2, o2, i2 = corpora.uniform(3000, 100, seed=200 + seed)
                other = build(t2, o2, i2)
                assert other.info()["n"] == 300000
                other.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=querier, args=(t,)) for t in range(8)] + [threading.Thread(target=builder)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    ix.close()


def test_config5_flavour_utf8_long_documents():
    """BASELINE configs[4] flavour at test scale (10.7 MB): valid UTF-8, ragged documents up to ~70 KB among 300 000
    short ones (bits1 + bits2 = 36 -> 64-bit suffix-array elements), bytes on both sides of 0x80 with n >> 4096 ->
    several signed-radix levels of the note-N1 layout (chuck_size = n/256 = 41 621).  Bit-exact suffix array and
    (id, count) rows against the oracle."""
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 40, size=300000)
    lens[rng.integers(0, 300000, size=40)] = rng.integers(20000, 50000, size=40)
    text, off, ids = corpora.utf8_fast(lens, seed=78)
    ix = build(text, off, ids)
    sa, b1, w = oracle.port.build_sa(text, off)
    inf = ix.info()
    assert (inf["bits"], inf["width"]) == (b1, w) and w == 8
    assert np.array_equal(ix.export_sa(), sa)
    spat, soff = corpora.sampled_patterns(text, off, 400, 1, 9, seed=79)
    row_off, pairs = ix.locate_batch(spat, soff)
    for q in range(400):
        kw = bytes(spat[soff[q]:soff[q + 1]])
        assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), kw
    # highlight on this layout goes through the direct document scan and must find every occurrence
    kws = [bytes(spat[soff[q]:soff[q + 1]]) for q in (3, 17, 101)]
    docs = [int(d) for d in np.argsort(-np.diff(off))[:3]] + [0, 5]
    for d, sp in zip(docs, ix.spans(kws, docs)):
        assert np.array_equal(sp, oracle.port.spans(kws, text[off[d]:off[d + 1]].tobytes()))
    ix.close()


def test_add_after_build_and_rebuild():
    """keep_host_copy = 0 (default): the staging copy is released by build(), later add()/build() fail loudly.
    keep_host_copy = 1: documents can be appended and the index rebuilt on the same handle."""
    ix = cdb.StringIndex()
    ix.add(1, b"abcabc")
    ix.build()
    with pytest.raises(RuntimeError, match="has been built"):
        ix.add(2, b"abc")
    with pytest.raises(RuntimeError, match="has been built"):
        ix.build()
    assert ix.query(b"abc") == [(1, 2)]
    ix.close()
    ix = cdb.StringIndex(keep_host_copy=True)
    ix.add(1, b"abcabc")
    ix.build()
    ix.add(2, b"xabc")
    ix.build()
    assert ix.query(b"abc") == [(1, 2), (2, 1)]
    ix.close()


def _fuzz_corpus(rng):
    """Small adversarial corpora: tiny alphabets, empty and one-byte documents, repeated documents, NUL and 0xFF
    bytes, documents shorter than a radix key, a few long runs."""
    kind = int(rng.integers(0, 6))
    nd = int(rng.integers(1, 400))
    if kind == 0:    # binary alphabet, many ties
        alpha = np.frombuffer(b"ab", np.uint8)
    elif kind == 1:  # bytes around the sign boundary and the extremes
        alpha = np.array([0x00, 0x01, 0x7F, 0x80, 0xFE, 0xFF], np.uint8)
    elif kind == 2:  # single symbol
        alpha = np.frombuffer(b"z", np.uint8)
    elif kind == 3:
        alpha = np.arange(97, 123, dtype=np.uint8)
    elif kind == 4:  # high bytes only
        alpha = np.arange(0x80, 0x90, dtype=np.uint8)
    else:
        alpha = np.frombuffer(b"acgt", np.uint8)
    maxlen = int(rng.choice([1, 2, 9, 13, 40, 300]))
    docs = []
    for _ in range(nd):
        ln = int(rng.integers(0, maxlen + 1)) if rng.random() > 0.15 else 0
        docs.append(bytes(alpha[rng.integers(0, len(alpha), size=ln)]))
    if nd > 3 and rng.random() < 0.5:  # exact duplicates of whole documents
        for _ in range(int(rng.integers(1, 10))):
            docs[int(rng.integers(0, nd))] = docs[int(rng.integers(0, nd))]
    if rng.random() < 0.3:
        docs[int(rng.integers(0, nd))] = bytes(alpha[:1]) * int(rng.integers(300, 3000))
    return corpora.from_docs(docs, id_base=int(rng.integers(-1000, 1000)))


def test_fuzz_small_corpora():
    rng = np.random.default_rng(20261017)
    for case in range(60):
        text, off, ids = _fuzz_corpus(rng)
        ix = build(text, off, ids)
        sa, b1, w = oracle.port.build_sa(text, off)
        inf = ix.info()
        assert (inf["n"], inf["bits"], inf["width"]) == (len(sa), b1, w), case
        assert np.array_equal(ix.export_sa(), sa), case
        pats = []
        if len(text):
            spat, soff = corpora.sampled_patterns(text, off, 30, 1, 12, seed=case)
            pats = [bytes(spat[soff[i]:soff[i + 1]]) for i in range(30)]
        alpha = np.unique(text) if len(text) else np.array([97], np.uint8)
        for _ in range(20):  # random keywords over (alphabet + one byte that may be absent)
            m = int(rng.integers(1, 6))
            pool = np.concatenate([alpha, np.array([int(rng.integers(0, 256))], np.uint8)])
            pats.append(bytes(pool[rng.integers(0, len(pool), size=m)]))
        row_off, pairs = ix.locate_batch(pats)
        for q, kw in enumerate(pats):
            assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], oracle.port.query(text, off, ids, sa, b1, kw)), (case, kw)
        nd = len(ids)
        docs = [int(d) for d in rng.integers(0, nd, size=min(nd, 5))]
        kws = pats[:3] if pats else [b"a"]
        for d, sp in zip(docs, ix.spans(kws, docs)):
            assert np.array_equal(sp, oracle.port.spans(kws, text[off[d]:off[d + 1]].tobytes())), (case, d)
        ix.close()


def test_reference_size_limit_errors():
    """The two limits string_index::build enforces (src/index.cpp:195-200), with the reference's exact messages.
    They only depend on the document-boundary array, so crafted boundaries reach them without a huge corpus."""
    import torch
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    text = torch.zeros(1024, dtype=torch.uint8, device=dev)
    # bits1 + bits2 > 64: three documents, one of them 2^62 bytes "long"
    doc_off = torch.tensor([0, 1 << 62, (1 << 62) + 1, (1 << 62) + 2], dtype=torch.int64, device=dev)
    ids = torch.arange(3, dtype=torch.int64, device=dev)
    ix = cdb.StringIndex(device=0)
    with pytest.raises(RuntimeError, match="^The amount of data exceeds the maximum range that CoffeeDB can handle$"):
        ix.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), 3, stream)
    ix.close()
    # the same check in the oracle's restatement of the width rule
    rc, b1, b2, _w = oracle.port.widths(doc_off.cpu().numpy())
    assert rc == 1 and b1 + b2 > 64
    # bits1 > 32: 2^32 (empty) documents
    if torch.cuda.mem_get_info()[0] > 60 * (1 << 30):
        nd = 1 << 32
        doc_off = torch.zeros(nd + 1, dtype=torch.int64, device=dev)
        ix = cdb.StringIndex(device=0)
        with pytest.raises(RuntimeError, match="^The number of objects exceeds the maximum range that CoffeeDB can handle$"):
            ix.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), nd, stream)
        ix.close()
        del doc_off
        torch.cuda.empty_cache()


# ---- cdb_verify_sa: the independent device-side order / permutation check ------------------------------------------
class _DevView:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


@pytest.mark.parametrize("name", ALL_CASES)
def test_verifier_accepts_every_golden_case(name, golden):
    """The verifier (adjacent pairs under the reference's comparator, signed rule in the note-N1 layout; permutation)
    accepts the arrays that are bit-identical to the compiled reference's, and counts exactly the ties the canonical
    array holds (adjacent elements with byte-identical suffixes)."""
    text, off, ids, _ = cases.CASES[name]()
    ix = build(text, off, ids)
    v = ix.verify_sa()
    assert v["ok"], v
    sa = ix.export_sa()
    inf = ix.info()
    docs = (sa & np.uint64(inf["mask"])).astype(np.int64)
    offs = (sa >> np.uint64(inf["bits"])).astype(np.int64)
    t = text.tobytes()
    suf = [t[off[d] + o: off[d + 1]] for d, o in zip(docs.tolist(), offs.tolist())] if inf["n"] <= 200_000 else None
    if suf is not None:
        assert v["ties"] == sum(1 for a, b in zip(suf, suf[1:]) if a == b)
    if name.startswith("n1"):
        assert v["signed_rule_pairs"] > 0  # the signed rule really decided pairs in the note-N1 cases
    ix.close()


def test_verifier_detects_corruption():
    import torch
    text, off, ids = corpora.uniform(300, 200, seed=77)
    ix = build(text, off, ids)
    inf = ix.info()
    assert ix.verify_sa()["ok"]
    sa = torch.as_tensor(_DevView(ix.sa_device_ptr(), inf["n"], "<i4" if inf["width"] == 4 else "<i8"), device="cuda:0")
    a, b = int(sa[1000]), int(sa[40000])
    sa[1000], sa[40000] = b, a          # two elements swapped: still a permutation, no longer sorted
    torch.cuda.synchronize()
    v = ix.verify_sa()
    assert not v["ok"] and v["inversions"] >= 2 and v["duplicates"] == 0 and v["invalid"] == 0
    sa[1000], sa[40000] = a, a          # one element twice: a duplicate position (and one missing)
    torch.cuda.synchronize()
    v = ix.verify_sa()
    assert not v["ok"] and v["duplicates"] == 1
    sa[40000] = b
    sa[5] = (199 + 50) << inf["bits"]   # offset beyond the end of document 0
    torch.cuda.synchronize()
    v = ix.verify_sa()
    assert not v["ok"] and v["invalid"] == 1
    ix.close()


def test_verifier_rejects_plain_order_where_reference_is_signed():
    """On a mixed-byte corpus the plainly sorted array (compat_signed = 0) is NOT the reference's layout: a verifier
    applying the reference's rules must flag it, and accept it when told the index is in plain order."""
    text, off, ids, _ = cases.CASES["n1_mixed"]()
    plain = build(text, off, ids, compat_signed=False)
    assert plain.verify_sa()["ok"]      # checked under its own (unsigned) rule
    ref = build(text, off, ids)
    assert ref.verify_sa()["ok"] and ref.verify_sa()["signed_rule_pairs"] > 0
    assert not np.array_equal(plain.export_sa(), ref.export_sa())
    plain.close()
    ref.close()


# ---- suffix-array persistence ---------------------------------------------------------------------------
def test_save_and_build_or_load(tmp_path):
    """cdb_save + cdb_build_or_load: the same corpus reads the array back (bit-identical, same answers, no sort rounds);
    a changed corpus, a changed compat flag or a truncated file falls back to a normal build."""
    text, off, ids = corpora.utf8ish(2500, 300, seed=95)  # note-N1 layout: the saved array includes the rotations
    path = str(tmp_path / "val.cdbsa")
    a = build(text, off, ids)
    a.save(path)
    sa = a.export_sa()
    pats = [b"a", b"ab", b"\xc3", b"the", b" "]
    want = a.locate_batch(pats)
    a.close()
    b = cdb.StringIndex()
    b.add_many(ids, text, off)
    assert b.build_or_load(path) is True
    assert np.array_equal(b.export_sa(), sa) and b.verify_sa()["ok"]
    got = b.locate_batch(pats)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    b.close()
    # one byte of the corpus changed -> stale file -> normal build of the new corpus
    text2 = text.copy()
    text2[len(text2) // 2] ^= 1
    c = cdb.StringIndex()
    c.add_many(ids, text2, off)
    assert c.build_or_load(path) is False
    sa2, _b1, _w = oracle.port.build_sa(text2, off)
    assert np.array_equal(c.export_sa(), sa2)
    c.close()
    # different ids (same text) also change the key
    d = cdb.StringIndex()
    d.add_many(ids + 1, text, off)
    assert d.build_or_load(path) is False
    d.close()
    # other compat flag -> different layout -> not loaded
    e = cdb.StringIndex(compat_signed=False)
    e.add_many(ids, text, off)
    assert e.build_or_load(path) is False
    e.close()
    # truncated file -> normal build, same result
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[: len(raw) // 2])
    f = cdb.StringIndex()
    f.add_many(ids, text, off)
    assert f.build_or_load(path) is False and np.array_equal(f.export_sa(), sa)
    f.close()
    # no file at all
    g = cdb.StringIndex()
    g.add_many(ids, text, off)
    assert g.build_or_load(str(tmp_path / "missing.cdbsa")) is False and np.array_equal(g.export_sa(), sa)
    g.close()


# ---- loader staging ------------------------------------------------------------------------------------
def test_loader_staging_uploads_full_chunks_during_add():
    """cdb_add streams into page-locked chunks (1, 2, 4 ... MB); full chunks are on the device before build() is called,
    the build result is the reference's, and a rejected add_many leaves the staged text untouched."""
    text, off, ids = corpora.uniform(60000, 100, seed=97)  # 6 MB: chunks of 1 + 2 MB fill up, the 4 MB one does not
    ix = cdb.StringIndex()
    half = 30000
    for d in range(0, half):  # document by document, like the reference's loader (src/database.cpp:255-264)
        ix.add(int(ids[d]), text[off[d]:off[d + 1]].tobytes())
    st = ix.staging_stats()
    assert st["staged"] == off[half] and st["on_device"] == (1 << 20)  # the 2 MB chunk is still filling
    bad = off[half:].copy() - off[half]
    bad[5] = bad[4] - 1  # decreasing offsets: rejected before anything is staged
    with pytest.raises(RuntimeError, match="non-decreasing"):
        ix.add_many(ids[half:], text[off[half]:], bad)
    assert ix.staging_stats()["staged"] == off[half]
    ix.add_many(ids[half:], text[off[half]:], off[half:] - off[half])
    assert ix.staging_stats()["staged"] == len(text)
    ix.build()
    assert ix.staging_stats() == {"staged": len(text), "on_device": (1 << 20) + (2 << 20)}
    sa, bits1, _w = oracle.port.build_sa(text, off)
    assert np.array_equal(ix.export_sa(), sa)
    kw = bytes(text[off[41000] + 3: off[41000] + 7])
    assert np.array_equal(ix.query_array(kw), oracle.port.query(text, off, ids, sa, bits1, kw))
    ix.close()


# ---- long repeats: rank doubling after the extension rounds (sa_build.cu finish_by_doubling) -----------------------------
def _repetitive_corpora():
    rng = np.random.default_rng(123)
    x = rng.integers(ord("a"), ord("e"), size=12000, dtype=np.uint8)
    yield "two identical documents", [x, x.copy()]
    yield "three copies and a prefix", [x, x[:7000].copy(), x.copy(), x.copy()]
    yield "periodic", [np.frombuffer(b"abcab" * 2500, np.uint8), np.frombuffer(b"abcab" * 1800 + b"x", np.uint8)]
    yield "runs of one byte", [np.full(9000, ord("z"), np.uint8), np.full(5000, ord("z"), np.uint8), np.frombuffer(b"zzzy", np.uint8)]
    y = rng.integers(0x20, 0xF0, size=9000, dtype=np.uint8)  # bytes on both sides of 0x80: note-N1 layout on top
    yield "identical mixed-byte documents", [y, y.copy(), rng.integers(0x20, 0xF0, size=3000, dtype=np.uint8)]
    shared = rng.integers(ord("a"), ord("c"), size=6000, dtype=np.uint8)
    yield "long shared substrings", [np.concatenate([rng.integers(ord("a"), ord("c"), size=int(k), dtype=np.uint8), shared])
                                     for k in rng.integers(1, 50, size=12)]


@pytest.mark.parametrize("name,docs", list(_repetitive_corpora()), ids=[n for n, _ in _repetitive_corpora()])
@pytest.mark.parametrize("chunked", [False, True])
def test_long_repeats_take_rank_doubling(name, docs, chunked, monkeypatch):
    """Corpora whose suffixes share thousands of leading bytes: key extension alone would need hundreds of rounds (S
    symbols each); after 4 of them the remaining ties are refined by rank doubling.  Same array as the reference port,
    few rounds, with one chunk and with a chunked build."""
    if chunked:
        monkeypatch.setenv("CDB_BUILD_WORKSPACE_MB", "1")
    text = np.concatenate(docs)
    off = np.zeros(len(docs) + 1, np.int64)
    off[1:] = np.cumsum([len(d) for d in docs])
    ids = np.arange(len(docs), dtype=np.int64) * 3 + 11
    ix = build(text, off, ids)
    st = ix.build_stats()
    sa, bits1, _w = oracle.port.build_sa(text, off)
    assert np.array_equal(ix.export_sa(), sa), name
    assert ix.verify_sa()["ok"]
    assert st["rounds"] <= 4 * max(st["chunks"], 1) + 20, st  # not hundreds
    kw = bytes(text[off[1] + 5: off[1] + 5 + 40])
    assert np.array_equal(ix.query_array(kw), oracle.port.query(text, off, ids, sa, bits1, kw))
    ix.close()


def test_two_identical_64k_documents_build_in_milliseconds():
    """two byte-identical ~64 KB UTF-8 documents needed ~8 000 extension rounds (8 symbols each).
    Timed, compared with the reference port and checked by the independent device verifier."""
    import time
    docs = []
    for seed in (3, 4):  # two different long documents, each present twice, plus a short one
        text, _off, _ids = corpora.utf8ish(1, 65536, seed=seed)
        docs += [text, text.copy()]
    docs.append(docs[0][:3000].copy())
    t = np.concatenate(docs)
    off = np.zeros(len(docs) + 1, np.int64)
    off[1:] = np.cumsum([len(d) for d in docs])
    ids = np.arange(len(docs), dtype=np.int64) + 5
    ix = cdb.StringIndex()
    ix.add_many(ids, t, off)
    t0 = time.perf_counter()
    ix.build()
    dt = time.perf_counter() - t0
    st = ix.build_stats()
    v = ix.verify_sa()
    assert v["ok"] and v["ties"] >= len(docs[0]), v  # every suffix of a copy ties with its twin
    sa, bits1, _w = oracle.port.build_sa(t, off)
    assert np.array_equal(ix.export_sa(), sa)
    assert st["rounds"] <= 30 and dt < 5.0, (st, dt)
    kw = bytes(docs[0][100:140])
    assert np.array_equal(ix.query_array(kw), oracle.port.query(t, off, ids, sa, bits1, kw))
    print(f"\nlong identical documents ({[len(d) for d in docs]} bytes): {st['rounds']} rounds, {dt * 1e3:.1f} ms")
    ix.close()
