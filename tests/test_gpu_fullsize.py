"""BASELINE configs[1] at full size (10^7 docs x 100 B = 1 GB, 10^5 patterns) on the device, checked through
size-independent properties and an independent brute-force count written in torch (no oracle at this size):

 * the suffix array is a permutation of all (doc, offset) pairs (sum / sum-of-squares checksums of the positions),
 * every sampled pattern's (id, count) row equals a brute-force scan of the whole text,
 * total pairs / occurrences of the 10^5-pattern batch equal the sum over two half batches (split invariance) and a
   second run returns identical bytes (idempotence),
 * rows are in ascending doc index (ids here are strictly increasing with the doc index)."""
import numpy as np
import pytest
import torch

import coffeedb_b200 as cdb
from tests import corpora

pytestmark = pytest.mark.gpu

ND, L = 10_000_000, 100


class _Dev:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


@pytest.fixture(scope="module")
def big():
    if torch.cuda.mem_get_info()[0] < 40 * (1 << 30):
        pytest.skip("needs 40 GB of free device memory")
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(20261017)
    text = torch.zeros(ND * L + 64, dtype=torch.uint8, device=dev)
    text[: ND * L] = torch.randint(97, 123, (ND * L,), dtype=torch.uint8, device=dev, generator=g)
    doc_off = torch.arange(ND + 1, dtype=torch.int64, device=dev) * L
    ids = torch.arange(ND, dtype=torch.int64, device=dev) * 3 + 7  # strictly increasing with the doc index
    ix = cdb.StringIndex(device=0)
    ix.build_device(text.data_ptr(), doc_off.data_ptr(), ids.data_ptr(), ND, torch.cuda.current_stream().cuda_stream,
                    keep=(text, doc_off, ids))
    yield ix, text, ids
    ix.close()


def brute(text, ids, kw: bytes):
    n = ND * L
    m = len(kw)
    hit = text[: n - m + 1] == kw[0]
    for j in range(1, m):
        hit &= text[j: n - m + 1 + j] == kw[j]
    pos = torch.nonzero(hit).flatten()
    pos = pos[(pos % L) <= L - m]  # an occurrence never crosses a document end
    docs, counts = torch.unique(pos // L, return_counts=True)
    return torch.stack([ids[docs], counts], dim=1).cpu().numpy()


def test_suffix_array_is_a_permutation(big):
    ix, _text, _ids = big
    inf = ix.info()
    assert inf["n"] == ND * L and inf["width"] == 4 and inf["bits"] == 24
    sa = torch.as_tensor(_Dev(ix.sa_device_ptr(), inf["n"], "<i4"), device="cuda:0")
    n = inf["n"]
    s1 = torch.zeros((), dtype=torch.int64, device="cuda:0")
    s2 = torch.zeros((), dtype=torch.int64, device="cuda:0")
    step = 1 << 27
    for lo in range(0, n, step):
        e = sa[lo: lo + step].to(torch.int64) & 0xFFFFFFFF
        pos = (e & inf["mask"]) * L + (e >> inf["bits"])
        assert int(pos.max()) < n
        s1 += pos.sum()
        s2 += (pos * pos).sum()  # wraps modulo 2^64, like the expectation below
    want1 = n * (n - 1) // 2
    want2 = (n - 1) * n * (2 * n - 1) // 6
    assert int(s1) == want1
    assert int(s2) % (1 << 64) == want2 % (1 << 64) or int(s2) == ((want2 + (1 << 63)) % (1 << 64)) - (1 << 63)


def test_suffix_array_order_verified_on_device(big):
    """n adjacent-pair compares with the comparator of src/index.cpp:92-93 + the permutation bit map, on the device,
    independent of the build and of the oracle: 0 inversions, 0 invalid, 0 duplicates."""
    ix, _text, _ids = big
    v = ix.verify_sa()
    assert v["ok"] and v["inversions"] == 0 and v["invalid"] == 0 and v["duplicates"] == 0, v


def test_rows_equal_brute_force_scan(big):
    ix, text, ids = big
    pat, poff = corpora.uniform_patterns(24, 5, seed=5)
    pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(24)]
    h = text[: 4096].cpu().numpy().tobytes()
    pats += [h[100:108], h[250:262], h[1300:1303], h[95:100], b"zzzzzzzzzz", b"q"]
    row_off, pairs = ix.locate_batch(pats[:-1])
    for q, kw in enumerate(pats[:-1]):
        got = pairs[row_off[q]:row_off[q + 1]]
        assert np.array_equal(got, brute(text, ids, kw)), kw
        assert np.all(np.diff(got[:, 0]) > 0)  # ascending doc index
    # a one-byte keyword hits every document: the large-interval path at full size
    ro, pr = ix.locate_batch([pats[-1]])
    assert np.array_equal(pr, brute(text, ids, pats[-1]))


def test_batch_split_invariance_and_idempotence(big):
    ix, _text, _ids = big
    pat, poff = corpora.uniform_patterns(100_000, 5, seed=1002)
    ro, pr = ix.locate_batch(pat, poff)
    ro2, pr2 = ix.locate_batch(pat, poff)
    assert np.array_equal(ro, ro2) and np.array_equal(pr, pr2)
    k = 37_123
    roa, pra = ix.locate_batch(pat[: poff[k]], poff[: k + 1])
    rob, prb = ix.locate_batch(pat[poff[k]:], poff[k:] - poff[k])
    assert np.array_equal(np.concatenate([pra, prb]), pr)
    assert np.array_equal(np.concatenate([roa, rob[1:] + roa[-1]]), ro)
    assert ix.last_total_occurrences == int(prb[:, 1].sum())
