"""Generates tests/golden/golden.npz by running the UNMODIFIED compiled reference
(oracle/_ref/libcoffeeref.so, built from /root/reference by oracle/Makefile) on tests/cases.py.

Run in the build container (where /root/reference exists):   python -m tests.golden.make_golden

Per case the file stores: bits1, element width, n, the reference's suffix array canonicalised per note N2
(full array when small, sha256 always), the raw SA sha256 for the N1 cases (whose exact layout matters),
and the reference's query() answer for every pattern as one CSR (row_off, pairs).  Highlight cases store
the reference's rendered string.  The known answers of README.md / examples/example.py are asserted here
so a wrong build of the reference cannot become "golden".
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

import oracle
from tests import cases, corpora

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
FULL_SA_MAX = 1 << 15


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, np.uint64).tobytes()).hexdigest()


def run_case(name, fn, store):
    text, off, ids, pats = fn()
    r = oracle.Ref()
    r.add_many(ids, text, off)
    r.build()
    raw, bits1, mask, width = r.export_sa()
    canon = oracle.port.canonicalise(text, off, raw, bits1)
    pb, po = corpora.pack_patterns(pats)
    row_off, pairs = r.query_batch(pb, po, nthreads=4)
    r.close()
    store[f"{name}/bits1"] = np.int64(bits1)
    store[f"{name}/width"] = np.int64(width)
    store[f"{name}/n"] = np.int64(len(raw))
    store[f"{name}/sa_sha"] = np.array(sha(canon))
    store[f"{name}/raw_sha"] = np.array(sha(raw))
    if len(raw) <= FULL_SA_MAX:
        store[f"{name}/sa"] = canon.astype(np.uint64)
        store[f"{name}/raw"] = raw.astype(np.uint64)
    store[f"{name}/pat"] = pb
    store[f"{name}/pat_off"] = po
    store[f"{name}/row_off"] = row_off
    store[f"{name}/pairs"] = pairs
    # everything but the N1 cases must equal brute force (test-string.py:52-56)
    if name not in cases.N1_CASES:
        for i, p in enumerate(pats[:25]):
            want = corpora.brute_count(text, off, ids, p)
            got = pairs[row_off[i]:row_off[i + 1]]
            assert np.array_equal(want, got), (name, p)
    print(f"{name:18s} n={len(raw):8d} bits1={bits1:2d} w={width} patterns={len(pats):4d} pairs={len(pairs)}")


def main():
    assert oracle.ref_available(), "needs /root/reference (build container)"
    store = {}
    for name, fn in cases.CASES.items():
        run_case(name, fn, store)
    # known answers, README.md:64-109 / SURVEY.md §8c
    ro, pr = store["readme/row_off"], store["readme/pairs"]
    assert pr[ro[0]:ro[1]].tolist() == [[100, 2], [101, 1]]          # "010"
    assert pr[ro[2]:ro[3]].tolist() == [[100, 3], [101, 2]]          # "0"
    assert pr[ro[4]:ro[5]].tolist() == [] and pr[ro[5]:ro[6]].tolist() == []
    b = int(store["readme/bits1"])
    assert [(int(x) >> b, int(x) & 3) for x in store["readme/sa"]] == [
        (1, 0), (1, 1), (3, 0), (3, 1), (5, 0), (2, 0), (2, 1), (4, 0), (5, 1), (4, 1), (6, 0), (0, 0), (0, 1)]
    assert store["abab/sa"].tolist() == [1, 8, 0, 5, 12, 4] and store["abab/raw"].tolist() == [8, 1, 0, 12, 5, 4]
    # highlight
    hl = cases.highlight_cases()
    rendered = [oracle.ref_render(kws, text, b"<b>", b"</b>") for kws, text in hl]
    assert rendered[0] == b"3<b>01010</b>3"
    store["highlight/rendered"] = np.frombuffer(b"".join(rendered), np.uint8)
    store["highlight/rendered_off"] = np.concatenate([[0], np.cumsum([len(x) for x in rendered])]).astype(np.int64)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
