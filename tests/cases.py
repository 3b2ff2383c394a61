"""Named parity cases: corpus + patterns.  Used by tests/golden/make_golden.py (which runs the compiled
reference on them) and by the CPU and GPU parity tests (which check the oracle port / the CUDA engine
against the committed golden file).  Sizes are chosen so the whole CPU suite runs in a few minutes."""
from __future__ import annotations

import numpy as np

from . import corpora as C


def _kat(docs, pats):
    text, off, ids = C.from_docs(docs)
    return text, off, ids, [bytes(p) for p in pats]


def _readme():  # README.md:64-109 / examples/example.py, SURVEY.md §8c KAT 1
    return _kat([b"3010103", b"301022"], [b"010", b"3", b"0", b"3010103", b"30101034", b"9", b"1", b"22", b"103"])


def _aaaa():
    return _kat([b"aaaa"], [b"a", b"aa", b"aaa", b"aaaa", b"aaaaa", b"b"])


def _empties():
    return _kat([b"", b"ab", b"", b"ab"], [b"ab", b"a", b"b", b"ba", b"abc"])


def _abab():
    return _kat([b"abab", b"ab"], [b"ab", b"ba", b"b", b"a", b"abab", b"bab"])


def _single_byte_docs():
    return _kat([b"x"] * 5 + [b"y"] + [b""] * 3 + [b"x"], [b"x", b"y", b"xy", b"z"])


def _nul_bytes():  # end-of-document sorts before 0x00 (index.h:66-73)
    return _kat([b"a\x00", b"a", b"a\x00\x00", b"\x00a", b"\x00"], [b"a", b"a\x00", b"\x00", b"\x00\x00", b"\x00a"])


def _deep():  # SURVEY.md §8c KAT 5, reduced: radix depth = document length
    docs = [b"xy"] * 6999 + [b"z" * 7000]
    return _kat(docs, [b"zz", b"xy", b"y", b"z" * 7000, b"z" * 6999, b"yz"])


def _cfg1():  # BASELINE configs[0]: 1k docs x 1 KB a-z, 3-char keywords (test-string.py at reduced scale)
    text, off, ids = C.uniform(1000, 1024, seed=1)
    pat, poff = C.uniform_patterns(100, 3, seed=101)
    pats = [bytes(pat[poff[i]:poff[i + 1]]) for i in range(100)]
    pat5, poff5 = C.uniform_patterns(20, 5, seed=105)
    pats += [bytes(pat5[poff5[i]:poff5[i + 1]]) for i in range(20)]
    spat, soff = C.sampled_patterns(text, off, 40, 1, 12, seed=108)
    pats += [bytes(spat[soff[i]:soff[i + 1]]) for i in range(40)]
    return text, off, ids, pats


def _ragged():
    text, off, ids = C.ragged(3000, 60, seed=7)
    spat, soff = C.sampled_patterns(text, off, 80, 1, 10, seed=8)
    pats = [bytes(spat[soff[i]:soff[i + 1]]) for i in range(80)] + [b"a", b"b", b"dddd", b"e", b"abcdabcdabcd"]
    return text, off, ids, pats


def _repetitive():
    text, off, ids = C.repetitive(seed=3)
    pats = [b"z", b"zz", b"z" * 4999, b"z" * 5000, b"z" * 5001, b"ab", b"ba", b"abab", b"abc", b"cab", b"bca" * 30,
            b"ab" * 1500, b"ab" * 1501, b"a", b"c", b"ca"]
    return text, off, ids, pats


def _highbytes():  # all bytes >= 0x80: signed and unsigned order agree, no N1 effect
    text, off, ids = C.uniform(400, 64, seed=11, lo=0x80, hi=0xFF)
    spat, soff = C.sampled_patterns(text, off, 60, 1, 4, seed=12)
    return text, off, ids, [bytes(spat[soff[i]:soff[i + 1]]) for i in range(60)]


def _mixed_small():  # bytes on both sides of 0x80 but n <= 4096: single leaf, plain unsigned order
    text, off, ids = C.uniform(40, 100, seed=13, lo=1, hi=255)
    spat, soff = C.sampled_patterns(text, off, 60, 1, 3, seed=14)
    return text, off, ids, [bytes(spat[soff[i]:soff[i + 1]]) for i in range(60)]


def _n1_mixed():  # SURVEY.md §8 note N1: 300 docs x 300 B of bytes 1..255, n > 4096 -> signed/unsigned quirk
    text, off, ids = C.uniform(300, 300, seed=15, lo=1, hi=255)
    spat, soff = C.sampled_patterns(text, off, 150, 1, 3, seed=16)
    pats = [bytes(spat[soff[i]:soff[i + 1]]) for i in range(150)]
    pat, poff = C.uniform_patterns(50, 1, seed=17, lo=1, hi=255)
    pats += [bytes(pat[poff[i]:poff[i + 1]]) for i in range(50)]
    return text, off, ids, pats


def _n1_utf8():
    text, off, ids = C.utf8ish(400, 200, seed=19)
    spat, soff = C.sampled_patterns(text, off, 150, 1, 6, seed=20)
    return text, off, ids, [bytes(spat[soff[i]:soff[i + 1]]) for i in range(150)]


def _wide64():  # bits1 + bits2 > 32 -> 64-bit suffix-array elements (index.cpp:203-208)
    rng = np.random.default_rng(23)
    docs = [bytes(rng.integers(97, 101, size=3, dtype=np.uint8)) for _ in range(70_000)]
    docs[12345] = bytes(rng.integers(97, 101, size=40_000, dtype=np.uint8))
    text, off, ids = C.from_docs(docs, id_base=1 << 40)
    spat, soff = C.sampled_patterns(text, off, 40, 1, 9, seed=24)
    return text, off, ids, [bytes(spat[soff[i]:soff[i + 1]]) for i in range(40)] + [b"a", b"abcd"]


CASES = {
    "readme": _readme,
    "aaaa": _aaaa,
    "empties": _empties,
    "abab": _abab,
    "single_byte_docs": _single_byte_docs,
    "nul_bytes": _nul_bytes,
    "deep": _deep,
    "cfg1": _cfg1,
    "ragged": _ragged,
    "repetitive": _repetitive,
    "highbytes": _highbytes,
    "mixed_small": _mixed_small,
    "n1_mixed": _n1_mixed,
    "n1_utf8": _n1_utf8,
    "wide64": _wide64,
}

# cases on which the reference's answers differ from brute force (note N1); everywhere else they agree
N1_CASES = {"n1_mixed", "n1_utf8"}


def empty_corpus():
    return np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.int64)


# Highlight cases (database.cpp:58-91): (keywords, text)
def highlight_cases():
    rng = np.random.default_rng(31)
    out = [
        ([b"010"], b"3010103"),                     # README.md:109 -> 3<b>01010</b>3
        ([b"ab", b"bc"], b"abcabxbc"),
        ([b"aa"], b"aaaa aa a"),
        ([b"ab", b"cd"], b"abcdab"),                # adjacent, not merged
        ([b"abcd", b"bc"], b"xabcdx"),              # nested
        ([b"a", b"aaa"], b"aaaaa"),
        ([b"xyz"], b""),
        ([b"\xc3\xa9", b"\xa9\xc3"], b"\xc3\xa9\xc3\xa9 e\xcc\x81"),
    ]
    for _ in range(200):
        alpha = rng.integers(1, 256, size=int(rng.integers(2, 5)), dtype=np.uint8)
        text = bytes(alpha[rng.integers(0, len(alpha), size=int(rng.integers(0, 80)))])
        kws = [bytes(alpha[rng.integers(0, len(alpha), size=int(rng.integers(1, 5)))]) for _ in range(int(rng.integers(1, 5)))]
        out.append((kws, text))
    return out
