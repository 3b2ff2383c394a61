"""The document listing of the prefix directory's buckets (locate.cu: build_listing / listing_emit_kernel): a keyword of
exactly `symbols` symbols is streamed from its bucket instead of being gathered, sorted and translated.  Rows must be
the rows of string_index::query (src/index.cpp:237-326) bit for bit — compared with the oracle port, and with the same
library run with CDB_LISTING=0 (the suffix-array path) — in doc order and, through cdb_filter, in id order; for ids of
every width, negative ids, documents that repeat inside a row, buckets longer than the warp path, batches that mix
listed and unlisted keywords, duplicate ids (the listing must refuse), and the swap of the two orders when memory is short."""
import numpy as np
import pytest

import coffeedb_b200 as cdb
import oracle
from tests import corpora

pytestmark = pytest.mark.gpu


def build(text, off, ids):
    ix = cdb.StringIndex()
    ix.add_many(ids, text, off)
    ix.build()
    return ix


def all_lengths_patterns(text, off, k, seed, n_each=120):
    """keywords of 1 .. k + 3 bytes: sampled from the corpus (they hit), plus a few random ones (mostly misses)"""
    pats = []
    for m in range(1, k + 4):
        p, po = corpora.sampled_patterns(text, off, n_each, m, m, seed=seed + m)
        pats += [bytes(p[po[i]:po[i + 1]]) for i in range(len(po) - 1)]
    rng = np.random.default_rng(seed)
    alphabet = np.unique(text)
    for _ in range(100):
        pats.append(bytes(rng.choice(alphabet, size=k).astype(np.uint8)))
    pats.append(b"\xff" * k)  # a byte the corpus does not have
    return pats


def check_against_oracle(ix, text, off, ids, pats, sample):
    sa, b1, _w = oracle.port.build_sa(text, off)
    row_off, pairs = ix.locate_batch(pats)
    for q in sample:
        want = oracle.port.query(text, off, ids, sa, b1, pats[q])
        assert np.array_equal(pairs[row_off[q]:row_off[q + 1]], want), pats[q]
    return row_off, pairs


ID_KINDS = {
    "narrow": lambda nd, rng: rng.permutation(nd).astype(np.int64) * 7 + 1_000_000,                     # 4 bytes per suffix
    "40 bits": lambda nd, rng: (np.arange(nd, dtype=np.int64) * 2654435761) % (1 << 40) + 10 ** 12,     # 4 + 1
    "48 bits": lambda nd, rng: rng.permutation(nd).astype(np.int64) * ((1 << 47) // nd) - (1 << 46),    # 4 + 2, negative ids
    "62 bits": lambda nd, rng: rng.permutation(nd).astype(np.int64) * ((1 << 62) // nd) - (1 << 61),    # 4 + 4
    "ascending": lambda nd, rng: np.arange(nd, dtype=np.int64) * 3 + 11,
}


@pytest.mark.parametrize("id_kind", list(ID_KINDS))
def test_listed_rows_equal_reference_rows(id_kind, monkeypatch):
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    rng = np.random.default_rng(3)
    # few symbols and short documents: many documents repeat inside a row; the directory is 5-6 symbols deep
    text, off, _ = corpora.ragged(4000, 90, seed=17, alphabet=b"abcd")
    ids = ID_KINDS[id_kind](len(off) - 1, rng)
    ix = build(text, off, ids)
    k = ix.prefix_directory()["symbols"]
    assert k >= 3
    info = ix.listing_info(0)
    assert info["present"] and info["hi_bytes"] == {"narrow": 0, "40 bits": 1, "48 bits": 2, "62 bits": 4, "ascending": 0}[id_kind]
    pats = all_lengths_patterns(text, off, k, seed=5)
    sample = range(0, len(pats), 3)
    row_off, pairs = check_against_oracle(ix, text, off, ids, pats, sample)
    st = cdb.last_locate_stats()
    n_k = sum(1 for p in pats if len(p) == k)
    assert 0 < st["nlisted"] <= n_k and st["listed_pairs"] > 0
    # the same batch on the suffix-array path only
    monkeypatch.setenv("CDB_LISTING", "0")
    ix0 = build(text, off, ids)
    assert not ix0.listing_info(0)["present"]
    row_off0, pairs0 = ix0.locate_batch(pats)
    assert cdb.last_locate_stats()["nlisted"] == 0
    assert np.array_equal(row_off, row_off0) and np.array_equal(pairs, pairs0)
    ix.close()
    ix0.close()


@pytest.mark.parametrize("bits", ["6", "12", "20"])
def test_only_listed_keywords_skip_the_gather(bits, monkeypatch):
    """a batch of k-symbol keywords alone: phases A and B are skipped altogether (the benchmark's case)"""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    monkeypatch.setenv("CDB_PTAB_BITS", bits)
    text, off, ids = corpora.uniform(3000, 200, seed=23)  # a-z: 5 bits per symbol -> 1, 2, 4 symbols
    ix = build(text, off, ids)
    k = ix.prefix_directory()["symbols"]
    assert k == int(bits) // 5
    p, po = corpora.sampled_patterns(text, off, 700, k, k, seed=9)
    pats = [bytes(p[po[i]:po[i + 1]]) for i in range(len(po) - 1)] + [b"~" * k, b"a" * k]
    row_off, pairs = check_against_oracle(ix, text, off, ids, pats, range(0, len(pats), 7))
    st = cdb.last_locate_stats()
    # k = 1: every bucket holds ~23 000 suffixes — not listed, the large-interval path answers; otherwise all rows are listed
    if k == 1:
        assert st["nlisted"] == 0 and st["nlarge"] > 0
    else:
        assert st["nlisted"] >= 700 and st["listed_pairs"] == st["pairs"]
    ix.close()


def test_long_buckets_are_not_listed_and_mix_with_listed_ones(monkeypatch):
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    # 'a' dominates: the buckets of "aaa.." are far longer than 1024 suffixes, the rare ones are listed
    rng = np.random.default_rng(8)
    docs = []
    for _ in range(3000):
        n = int(rng.integers(20, 120))
        d = rng.choice(np.frombuffer(b"aaaaaaaabcd", np.uint8), size=n)
        docs.append(d.tobytes())
    text, off, ids = corpora.from_docs(docs)
    ix = build(text, off, ids)
    k = ix.prefix_directory()["symbols"]
    pats = all_lengths_patterns(text, off, k, seed=31) + [b"a" * k, b"a" * (k - 1) + b"b"]
    check_against_oracle(ix, text, off, ids, pats, range(0, len(pats), 2))
    st = cdb.last_locate_stats()
    assert st["nlisted"] > 0 and st["nlarge"] > 0
    ix.close()


def test_duplicate_ids_refuse_the_listing(monkeypatch):
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    text, off, ids = corpora.ragged(3000, 80, seed=4, alphabet=b"abc")
    ids = ids // 2  # pairs of documents share an id: equal neighbours in a listed row would be ambiguous
    ix = build(text, off, ids)
    assert ix.prefix_directory()["symbols"] > 0 and not ix.listing_info(0)["present"]
    k = ix.prefix_directory()["symbols"]
    pats = all_lengths_patterns(text, off, k, seed=2)
    check_against_oracle(ix, text, off, ids, pats, range(0, len(pats), 3))
    assert cdb.last_locate_stats()["nlisted"] == 0
    ix.close()


def _requests(kws):
    return [{"constraints": {"t": kw.decode()}, "span": "[0,32)"} for kw in kws] + \
           [{"constraints": {"t": [kws[i].decode(), kws[-1 - i].decode()]}} for i in range(40)] + \
           [{"constraints": {"t": kws[i].decode(), "$correlation": "[2,9)"}} for i in range(40)]


@pytest.mark.parametrize("budget_mb", [None, "1"])
def test_id_order_listing_through_filter(budget_mb, monkeypatch):
    """cdb_filter merges rows in id order: with ids that do not ascend with the doc index that is a listing of its own.
    With the budget of 1 MB there is room for one listing only: the two orders swap, answers stay the same."""
    text, off, ids = corpora.ragged(5000, 100, seed=12, alphabet=b"abcde")  # ids: a permutation
    ix = build(text, off, ids)
    k = ix.prefix_directory()["symbols"]
    p, po = corpora.sampled_patterns(text, off, 400, k, k, seed=77)
    kws = [bytes(p[po[i]:po[i + 1]]) for i in range(len(po) - 1)]
    reqs = _requests(kws)
    if budget_mb:
        monkeypatch.setenv("CDB_LISTING_BUDGET_MB", budget_mb)
        assert ix.listing_info(0)["bytes"] > (1 << 19)
    got = cdb.filter_batch({"t": ix}, reqs)
    if budget_mb:
        # the other order's listing is in the way: the first call takes the suffix-array path, the second one swaps
        assert not ix.listing_info(1)["present"] and ix.listing_info(0)["present"]
        again = cdb.filter_batch({"t": ix}, reqs)
        for (gp, gm), (wp, wm) in zip(got, again):
            assert gm == wm and np.array_equal(gp, wp)
    assert ix.listing_info(1)["present"]
    assert ix.listing_info(0)["present"] == (budget_mb is None)
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    row_off, pairs = ix.locate_batch(kws)  # doc order again
    if budget_mb:
        assert cdb.last_locate_stats()["nlisted"] == 0 and ix.listing_info(1)["present"]
        # alternating orders never swap: the listing in place stays, the other order keeps the suffix-array path
        cdb.filter_batch({"t": ix}, reqs[:50])
        ix.locate_batch(kws)
        assert cdb.last_locate_stats()["nlisted"] == 0 and ix.listing_info(1)["present"] and not ix.listing_info(0)["present"]
        ix.locate_batch(kws)  # twice in a row: now it swaps
    assert cdb.last_locate_stats()["nlisted"] > 0
    assert ix.listing_info(0)["present"] and ix.listing_info(1)["present"] == (budget_mb is None)
    monkeypatch.delenv("CDB_LISTING_BUDGET_MB", raising=False)
    monkeypatch.setenv("CDB_LISTING", "0")
    ix0 = build(text, off, ids)
    want = cdb.filter_batch({"t": ix0}, reqs)
    assert not ix0.listing_info(1)["present"]
    assert len(got) == len(want)
    for (gp, gm), (wp, wm) in zip(got, want):
        assert gm == wm and np.array_equal(gp, wp)
    row_off0, pairs0 = ix0.locate_batch(kws)
    assert np.array_equal(row_off, row_off0) and np.array_equal(pairs, pairs0)
    ix.close()
    ix0.close()


@pytest.mark.parametrize("range_bits", ["8", "11", "22"])
def test_two_phase_build_equals_one_phase(range_bits, monkeypatch):
    """Indexes whose ids[] outgrow L2 build the listing in two phases (sorted keys + split points, then ids range by range:
    listing_translate_kernel); forced here at small size, with many / a few / one doc range, in both orders."""
    monkeypatch.setenv("CDB_SMALL_BATCH", "0")
    rng = np.random.default_rng(5)
    text, off, _ = corpora.ragged(6000, 90, seed=29, alphabet=b"abcd")
    ids = ID_KINDS["40 bits"](len(off) - 1, rng)
    monkeypatch.setenv("CDB_LISTING_TWO_PHASE", "0")
    one = build(text, off, ids)
    k = one.prefix_directory()["symbols"]
    pats = all_lengths_patterns(text, off, k, seed=41)
    row_off1, pairs1 = one.locate_batch(pats)
    assert cdb.last_locate_stats()["nlisted"] > 0
    kws = [p for p in pats if len(p) == k and max(p) < 0x80][:300]
    reqs = _requests(kws)
    want = cdb.filter_batch({"t": one}, reqs)
    monkeypatch.setenv("CDB_LISTING_TWO_PHASE", "1")
    monkeypatch.setenv("CDB_RANGE_BITS", range_bits)
    two = build(text, off, ids)
    assert two.listing_info(0)["present"] and two.listing_info(0)["hi_bytes"] == 1
    row_off2, pairs2 = two.locate_batch(pats)
    assert cdb.last_locate_stats()["nlisted"] > 0
    assert np.array_equal(row_off1, row_off2) and np.array_equal(pairs1, pairs2)
    got = cdb.filter_batch({"t": two}, reqs)
    assert two.listing_info(1)["present"]
    for (gp, gm), (wp, wm) in zip(got, want):
        assert gm == wm and np.array_equal(gp, wp)
    # duplicate ids are found by the two-phase build's own check
    dup = build(text, off, np.repeat(ids[::2], 2)[: len(ids)].copy())  # pairs of documents share an id
    assert not dup.listing_info(0)["present"]
    for ix in (one, two, dup):
        ix.close()
