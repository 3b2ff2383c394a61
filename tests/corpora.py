"""Seeded synthetic corpora and pattern workloads shared by the tests, the golden-vector generator and
bench.py.  Pure numpy; nothing here touches the oracle or the CUDA engine.

A corpus is (text uint8[n], doc_off int64[nd+1], ids int64[nd]); document d = text[doc_off[d]:doc_off[d+1]].
"""
from __future__ import annotations

import numpy as np


def _pack(docs):
    off = np.zeros(len(docs) + 1, np.int64)
    if docs:
        off[1:] = np.cumsum([len(d) for d in docs])
    text = np.frombuffer(b"".join(docs), dtype=np.uint8).copy() if off[-1] else np.zeros(0, np.uint8)
    return text, off


def from_docs(docs, id_base=100):
    text, off = _pack([bytes(d) for d in docs])
    ids = np.arange(len(docs), dtype=np.int64) + id_base
    return text, off, ids


def uniform(nd: int, doclen: int, seed: int, lo=ord("a"), hi=ord("z")):
    """nd documents of exactly doclen bytes, uniform in [lo, hi] (configs 1-4 use a-z)."""
    rng = np.random.default_rng(seed)
    text = rng.integers(lo, hi + 1, size=nd * doclen, dtype=np.uint8)
    off = np.arange(nd + 1, dtype=np.int64) * doclen
    ids = rng.permutation(nd).astype(np.int64) * 7 + 1_000_000  # ids are NOT in doc order (SURVEY §8b "Doc order")
    return text, off, ids


def ragged(nd: int, maxlen: int, seed: int, alphabet=None, p_empty=0.1):
    """Variable-length documents, some empty, bytes drawn from `alphabet` (default a-d: many ties)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(alphabet if alphabet is not None else b"abcd", dtype=np.uint8)
    lens = rng.integers(1, maxlen + 1, size=nd)
    lens[rng.random(nd) < p_empty] = 0
    off = np.zeros(nd + 1, np.int64)
    off[1:] = np.cumsum(lens)
    text = alphabet[rng.integers(0, len(alphabet), size=int(off[-1]))]
    ids = rng.permutation(nd).astype(np.int64) + 5_000
    return text, off, ids


def utf8ish(nd: int, maxlen: int, seed: int):
    """Config-5 flavour: valid UTF-8 mixing 1-, 2- and 3-byte code points; exercises note N1 (bytes on both
    sides of 0x80)."""
    rng = np.random.default_rng(seed)
    docs = []
    for _ in range(nd):
        ln = int(rng.integers(0, maxlen + 1))
        out = bytearray()
        while len(out) < ln:
            r = rng.random()
            if r < 0.70:
                out += bytes([int(rng.integers(0x20, 0x7F))])
            elif r < 0.85:
                out += chr(int(rng.integers(0x80, 0x800))).encode()
            else:
                cp = int(rng.integers(0x800, 0xD800))
                out += chr(cp).encode()
        docs.append(bytes(out))
    text, off = _pack(docs)
    ids = np.arange(nd, dtype=np.int64) * 3 + 11
    return text, off, ids


def repetitive(seed: int):
    """Long runs and periodic documents: deep radix recursion / many doubling rounds."""
    rng = np.random.default_rng(seed)
    docs = [b"z" * 5000, b"ab" * 1500, b"abc" * 700, b"z" * 4999, b"ab" * 1500, b"", b"a", b"zz"]
    docs += [bytes(rng.integers(ord("a"), ord("c") + 1, size=int(rng.integers(1, 40)), dtype=np.uint8)) for _ in range(200)]
    return from_docs(docs, id_base=-50)


def uniform_patterns(npat: int, m: int, seed: int, lo=ord("a"), hi=ord("z")):
    """W{m}: npat patterns of m uniform bytes (test-string.py:40-41, benchmark.py:35)."""
    rng = np.random.default_rng(seed)
    pat = rng.integers(lo, hi + 1, size=npat * m, dtype=np.uint8)
    off = np.arange(npat + 1, dtype=np.int64) * m
    return pat, off


def sampled_patterns(text, doc_off, npat: int, mmin: int, mmax: int, seed: int):
    """W{m}s: substrings sampled from inside documents (every pattern hits at least once)."""
    rng = np.random.default_rng(seed)
    nd = len(doc_off) - 1
    lens = np.diff(doc_off)
    cand = np.nonzero(lens >= 1)[0]
    pats = []
    for _ in range(npat):
        d = int(cand[rng.integers(0, len(cand))])
        ln = int(lens[d])
        m = int(min(ln, rng.integers(mmin, mmax + 1)))
        s = int(rng.integers(0, ln - m + 1))
        pats.append(bytes(text[doc_off[d] + s: doc_off[d] + s + m]))
    return _pack(pats)


def pack_patterns(pats):
    return _pack([bytes(p) for p in pats])


def brute_count(text, doc_off, ids, kw: bytes):
    """Overlapping-occurrence count per document, the property test-string.py:14-19 checks."""
    out = []
    raw = text.tobytes()
    for d in range(len(doc_off) - 1):
        doc = raw[doc_off[d]:doc_off[d + 1]]
        c, s = 0, doc.find(kw)
        while s >= 0:
            c += 1
            s = doc.find(kw, s + 1)
        if c:
            out.append((int(ids[d]), c))
    return np.array(out, np.int64).reshape(-1, 2)


def utf8_fast(char_lens, seed: int):
    """Vectorised config-5 flavour corpus: document d holds char_lens[d] code points, 70 % one-byte (0x20..0x7E),
    15 % two-byte, 15 % three-byte UTF-8 sequences (valid UTF-8; bytes on both sides of 0x80 -> note N1)."""
    rng = np.random.default_rng(seed)
    char_lens = np.asarray(char_lens, np.int64)
    nchar = int(char_lens.sum())
    cls = rng.choice(np.array([1, 2, 3], np.int64), size=nchar, p=[0.70, 0.15, 0.15])
    start = np.zeros(nchar + 1, np.int64)
    np.cumsum(cls, out=start[1:])
    text = np.zeros(int(start[-1]), np.uint8)
    s = start[:-1]
    one, two, three = cls == 1, cls == 2, cls == 3
    text[s[one]] = rng.integers(0x20, 0x7F, size=int(one.sum()), dtype=np.uint8)
    cp2 = rng.integers(0x80, 0x800, size=int(two.sum()))
    text[s[two]] = 0xC0 | (cp2 >> 6)
    text[s[two] + 1] = 0x80 | (cp2 & 0x3F)
    cp3 = rng.integers(0x800, 0xD800, size=int(three.sum()))
    text[s[three]] = 0xE0 | (cp3 >> 12)
    text[s[three] + 1] = 0x80 | ((cp3 >> 6) & 0x3F)
    text[s[three] + 2] = 0x80 | (cp3 & 0x3F)
    cend = np.zeros(len(char_lens) + 1, np.int64)
    np.cumsum(char_lens, out=cend[1:])
    off = start[cend]
    ids = np.arange(len(char_lens), dtype=np.int64) * 5 + 3
    return text, off, ids


MAX_CHARS = 44_000  # 1.45 bytes per code point on average: the longest documents reach ~64 000 bytes, never 65 536
SHORT_CHARS = 2_000  # three quarters of the documents are short, so that 1 GiB already holds > 65 535 documents


def utf8_corpus_on_device(target_bytes: int, seed: int, device_index: int = 0):
    """Documents of up to 64 KB: 70 % one-byte (0x20..0x7E), 15 % two-byte, 15 % three-byte code points (SURVEY.md 8d
    config 5), generated with torch on the GPU -> (text uint8 [n + 64], doc_off int64 [nd + 1], ids, nd, n >= target).
    Lengths are skewed (75 % of the documents have 1..SHORT_CHARS code points, the rest 1..MAX_CHARS) so that a 1 GiB
    corpus has more than 2^16 documents AND documents longer than 2^15 bytes: bits1 + bits2 > 32, 64-bit elements."""
    import torch
    dev = torch.device("cuda", device_index)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    avg_chars = 0.75 * (SHORT_CHARS + 1) / 2 + 0.25 * (MAX_CHARS + 1) / 2
    nd = int(1.03 * target_bytes / (1.45 * avg_chars)) + 4
    short = torch.rand(nd, device=dev, generator=g) < 0.75
    char_lens = torch.where(short, torch.randint(1, SHORT_CHARS + 1, (nd,), device=dev, generator=g),
                            torch.randint(1, MAX_CHARS + 1, (nd,), device=dev, generator=g))
    del short
    nchar = int(char_lens.sum())
    u = torch.rand(nchar, device=dev, generator=g)
    cls = (1 + (u >= 0.70).to(torch.int8) + (u >= 0.85).to(torch.int8))
    del u
    start = torch.zeros(nchar + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cls, 0, out=start[1:])
    n = int(start[-1])
    text = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
    s = start[:-1]
    one = s[cls == 1]
    text[one] = torch.randint(0x20, 0x7F, (one.numel(),), device=dev, generator=g).to(torch.uint8)
    del one
    two = s[cls == 2]
    cp2 = torch.randint(0x80, 0x800, (two.numel(),), device=dev, generator=g)
    text[two] = (0xC0 | (cp2 >> 6)).to(torch.uint8)
    text[two + 1] = (0x80 | (cp2 & 0x3F)).to(torch.uint8)
    del two, cp2
    three = s[cls == 3]
    cp3 = torch.randint(0x800, 0xD800, (three.numel(),), device=dev, generator=g)
    text[three] = (0xE0 | (cp3 >> 12)).to(torch.uint8)
    text[three + 1] = (0x80 | ((cp3 >> 6) & 0x3F)).to(torch.uint8)
    text[three + 2] = (0x80 | (cp3 & 0x3F)).to(torch.uint8)
    del three, cp3, cls, s
    cend = torch.zeros(nd + 1, dtype=torch.int64, device=dev)
    torch.cumsum(char_lens, 0, out=cend[1:])
    doc_off = start[cend].contiguous()
    del start, cend
    ids = torch.arange(nd, dtype=torch.int64, device=dev) * 7 + 1_000_003
    torch.cuda.empty_cache()
    return text, doc_off, ids, nd, n
