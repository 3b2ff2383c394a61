"""Pins the oracle: the plain-C restatement (oracle/coffee_oracle.c) must reproduce
 (a) the golden vectors generated from the compiled, unmodified reference (tests/golden/golden.npz),
 (b) the live compiled reference where oracle/_ref is present, and
 (c) the known answers of README.md / examples/example.py and the brute-force property of test-string.py."""
import hashlib

import numpy as np
import pytest

import oracle
from tests import cases, corpora


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, np.uint64).tobytes()).hexdigest()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_port_matches_golden(name, golden):
    text, off, ids, pats = cases.CASES[name]()
    sa, bits1, width = oracle.port.build_sa(text, off)
    assert bits1 == int(golden[f"{name}/bits1"]) and width == int(golden[f"{name}/width"])
    assert len(sa) == int(golden[f"{name}/n"])
    assert sha(sa) == str(golden[f"{name}/sa_sha"]), "canonical suffix array differs from the reference's"
    if f"{name}/sa" in golden:
        assert np.array_equal(sa, golden[f"{name}/sa"])
    row_off, pairs = golden[f"{name}/row_off"], golden[f"{name}/pairs"]
    for i, p in enumerate(pats):
        got = oracle.port.query(text, off, ids, sa, bits1, p)
        assert np.array_equal(got, pairs[row_off[i]:row_off[i + 1]]), (name, p)


@pytest.mark.parametrize("name", ["readme", "abab", "ragged", "n1_mixed", "wide64"])
def test_port_matches_live_reference(name):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    text, off, ids, pats = cases.CASES[name]()
    r = oracle.Ref()
    r.add_many(ids, text, off)
    r.build()
    raw, bits1, mask, width = r.export_sa()
    sa, b1, w = oracle.port.build_sa(text, off)
    assert (b1, w) == (bits1, width) and mask == (1 << bits1) - 1
    assert np.array_equal(oracle.port.canonicalise(text, off, raw, bits1), sa)
    for p in pats:
        assert np.array_equal(r.query(p), oracle.port.query(text, off, ids, sa, bits1, p))
        # note N2: the answer does not depend on the order inside runs of equal suffixes
        assert np.array_equal(r.query(p), oracle.port.query(text, off, ids, raw, bits1, p))
    r.close()


def test_known_answers():
    text, off, ids, _ = cases.CASES["readme"]()
    sa, b1, _w = oracle.port.build_sa(text, off)
    q = lambda kw: oracle.port.query(text, off, ids, sa, b1, kw).tolist()
    assert q(b"010") == [[100, 2], [101, 1]]
    assert q(b"3") == [[100, 2], [101, 1]]
    assert q(b"0") == [[100, 3], [101, 2]]
    assert q(b"3010103") == [[100, 1]]
    assert q(b"30101034") == [] and q(b"9") == []
    with pytest.raises(RuntimeError, match="Empty keywords are not allowed"):
        q(b"")
    text, off, ids, _ = cases.CASES["aaaa"]()
    sa, b1, _w = oracle.port.build_sa(text, off)
    assert [oracle.port.query(text, off, ids, sa, b1, b"a" * k).tolist() for k in (1, 2, 4, 5)] == [
        [[100, 4]], [[100, 3]], [[100, 1]], []]


def test_empty_corpus_and_widths():
    text, off, ids = cases.empty_corpus()
    sa, b1, w = oracle.port.build_sa(text, off)
    assert len(sa) == 0 and b1 == 1 and w == 4
    assert oracle.port.query(text, off, ids, sa, b1, b"a").shape == (0, 2)
    # nd=0/1 -> 1 bit, 2 -> 2, 4 -> 3, 1000 -> 10, 70000 -> 17
    for nd, want in [(1, 1), (2, 2), (4, 3), (1000, 10), (70000, 17)]:
        rc, b1, _b2, _w = oracle.port.widths(np.arange(nd + 1, dtype=np.int64))
        assert rc == 0 and b1 == want
    # index.cpp:195-200 limits
    rc, *_ = oracle.port.widths(np.array([0, 1 << 40] + [1 << 40] * ((1 << 25) - 1), dtype=np.int64))
    assert rc == 1


def test_brute_force_property():
    # test-string.py:52-56 at reduced scale, on the restatement
    text, off, ids = corpora.uniform(200, 300, seed=5)
    sa, b1, _ = oracle.port.build_sa(text, off)
    pat, poff = corpora.uniform_patterns(30, 2, seed=6)
    for i in range(30):
        kw = bytes(pat[poff[i]:poff[i + 1]])
        assert np.array_equal(oracle.port.query(text, off, ids, sa, b1, kw), corpora.brute_count(text, off, ids, kw))


def test_highlight_matches_golden(golden):
    hl = cases.highlight_cases()
    rend, roff = golden["highlight/rendered"], golden["highlight/rendered_off"]
    for i, (kws, text) in enumerate(hl):
        spans = oracle.port.spans(kws, text)
        got = oracle.port.splice(text, spans, b"<b>", b"</b>")
        assert got == rend[roff[i]:roff[i + 1]].tobytes(), (kws, text)
        if oracle.ref_available():
            assert got == oracle.ref_render(kws, text, b"<b>", b"</b>")
