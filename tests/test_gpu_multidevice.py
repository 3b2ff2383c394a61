"""cdb_sharded_*: one string_index over several devices of ONE process (SURVEY.md 8e behind the drop-in boundary).  On a
one-GPU box the shards share device 0 — the doc-range split, the concurrent builds, the device-to-device broadcast and
exchange and the scatter of the row parts into the result are the same code; with more GPUs every shard gets its own."""
import numpy as np
import pytest

import coffeedb_b200 as cdb
import oracle
from tests import corpora

pytestmark = pytest.mark.gpu


def device_list(nshards):
    n = cdb.lib().cdb_device_count()
    return [g % n for g in range(nshards)]


@pytest.mark.parametrize("nshards", [1, 2, 3, 8])
def test_rows_equal_whole_corpus_reference(nshards):
    text, off, ids = corpora.ragged(4000, 90, seed=31)
    mi = cdb.MultiDeviceIndex(device_list(nshards))
    mi.add_many(ids[:1500], text[: off[1500]], off[:1501])
    for d in range(1500, 1600):  # document by document, as the loader does
        mi.add(int(ids[d]), text[off[d]:off[d + 1]].tobytes())
    mi.add_many(ids[1600:], text, off[1600:])
    mi.build()
    sa, bits1, _w = oracle.port.build_sa(text, off)
    pats = [b"a", b"ab", b"abc", b"dd", b"abcd", b"ca", b"zz"]
    spat, soff = corpora.sampled_patterns(text, off, 300, 2, 7, seed=5)
    pats += [bytes(spat[soff[i]:soff[i + 1]]) for i in range(300)]
    ro, pr = mi.locate_batch(pats)
    assert len(ro) == len(pats) + 1
    for q, kw in enumerate(pats):
        want = oracle.port.query(text, off, ids, sa, bits1, kw)
        assert np.array_equal(pr[ro[q]:ro[q + 1]], want), (nshards, kw)
    # the shards are contiguous doc ranges of about equal bytes, each an ordinary index of its own documents
    shards = mi.shards()
    assert len(shards) == nshards and shards[0][1] == 0 and shards[-1][2] == len(ids)
    for (view, b, e), (_v2, b2, _e2) in zip(shards, shards[1:] + [(None, len(ids), None)]):
        assert e == b2 and b <= e
        sub_sa, _b1, _w2 = oracle.port.build_sa(text[off[b]:off[e]], off[b:e + 1] - off[b])
        assert np.array_equal(view.export_sa(), sub_sa)
    if nshards > 1:
        sizes = [off[e] - off[b] for _v, b, e in shards]
        assert max(sizes) - min(sizes) <= 2 * 90 * nshards
    assert mi.query(b"abcd") == [tuple(map(int, r)) for r in oracle.port.query(text, off, ids, sa, bits1, b"abcd")]
    mi.close()


def test_errors_and_empty_cases():
    mi = cdb.MultiDeviceIndex(device_list(2))
    with pytest.raises(RuntimeError, match="has not been built"):
        mi.locate_batch([b"a"])
    mi.add(7, b"hello")  # fewer documents than shards: one shard stays empty
    mi.build()
    assert mi.query(b"ell") == [(7, 1)]
    assert mi.query(b"x") == []
    with pytest.raises(RuntimeError, match="Empty keywords are not allowed"):
        mi.locate_batch([b"a", b""])
    ro, pr = mi.locate_batch([])
    assert ro.tolist() == [0] and len(pr) == 0
    with pytest.raises(RuntimeError, match="has been built"):
        mi.add(8, b"more")
    mi.close()


def test_large_batch_matches_single_device_index():
    text, off, ids = corpora.uniform(200000, 100, seed=77)
    pat, poff = corpora.uniform_patterns(20000, 3, seed=78)
    one = cdb.StringIndex()
    one.add_many(ids, text, off)
    one.build()
    want_off, want = one.locate_batch(pat, poff)
    one.close()
    mi = cdb.MultiDeviceIndex(device_list(4))
    mi.add_many(ids, text, off)
    mi.build()
    got_off, got = mi.locate_batch(pat, poff)
    assert np.array_equal(got_off, want_off) and np.array_equal(got, want)
    mi.close()
