"""Host model of the distribution sort gather_kernel uses for intervals whose documents are spread over the corpus
(coffeedb_b200/csrc/locate.cu, warp_bucket_sort): same bucket function, same padded shared-memory indexing, same
ownership of counters by lanes, atomics applied in an arbitrary order, same odd-even transposition clean-up.  It pins
the three facts the kernel relies on: the bucket index is monotone in the doc index and stays below N for every doc
< nd; `largest bucket` alternating phases finish the job regardless of where a bucket starts; and the padded layouts
never collide.  The kernel itself is checked against the reference on the GPU
(tests/test_gpu_parity.py::test_gather_distribution_sort_and_fallback)."""
import itertools
import random

import numpy as np
import pytest

K_BUCKET_MAX = 24
SENTINEL = 0xFFFFFFFF


def pad(i):
    return i + (i >> 5)


def bucket_mul(nd):
    return min(0xFFFFFFFF, (1024 << 32) // nd)


def warp_bucket_sort(docs, R, nd):
    """Returns (keys in blocked layout or None when the kernel would fall back, largest bucket)."""
    N, sh = 32 * R, {32: 0, 16: 1, 8: 2}[R]
    occ, mul = len(docs), bucket_mul(nd)
    x = [[docs[r * 32 + lane] if r * 32 + lane < occ else SENTINEL for r in range(R)] for lane in range(32)]
    s_cnt, s_out = [None] * (33 * R), [None] * (33 * R)
    for lane in range(32):  # uint4 stores
        for t in range((33 * R // 4 + 31) // 32):
            if t * 32 + lane < 33 * R // 4:
                for k in range(4):
                    s_cnt[4 * (t * 32 + lane) + k] = 0
    assert None not in s_cnt

    def bucket(d):
        return pad(min(((d * mul) >> 32) >> sh, N - 1))

    for lane in range(32):
        for r in range(R):
            if r * 32 + lane < occ:
                s_cnt[bucket(x[lane][r])] += 1
    sums, mx = [], 0
    for lane in range(32):
        base = lane * R + ((lane * R) >> 5)
        c = [s_cnt[base + j] for j in range(R)]
        sums.append(sum(c))
        mx = max(mx, max(c))
    assert sum(sums) == occ  # every counter is owned by exactly one lane
    if mx > K_BUCKET_MAX:
        return None, mx
    incl = np.cumsum(sums)
    for lane in range(32):
        base, run = lane * R + ((lane * R) >> 5), int(incl[lane] - sums[lane])
        for j in range(R):
            c = s_cnt[base + j]
            s_cnt[base + j] = run
            run += c
    order = [(lane, r) for r in range(R) for lane in range(32)]
    random.shuffle(order)  # the order in which the atomics of a warp land is not defined
    for lane, r in order:
        if r * 32 + lane < occ:
            b = bucket(x[lane][r])
            p = s_cnt[b]
            s_cnt[b] += 1
            assert s_out[pad(p)] is None
            s_out[pad(p)] = x[lane][r]
    for lane in range(32):
        base = lane * R + ((lane * R) >> 5)
        for r in range(R):
            x[lane][r] = s_out[base + r] if lane * R + r < occ else SENTINEL
    if mx >= 2:
        for _ in range(0, mx, 2):
            for lane in range(32):
                for r in range(0, R - 1, 2):
                    a, b = x[lane][r], x[lane][r + 1]
                    x[lane][r], x[lane][r + 1] = min(a, b), max(a, b)
            for lane in range(32):
                for r in range(1, R - 1, 2):
                    a, b = x[lane][r], x[lane][r + 1]
                    x[lane][r], x[lane][r + 1] = min(a, b), max(a, b)
            up = [x[l + 1][0] if l < 31 else None for l in range(32)]
            dn = [x[l - 1][R - 1] if l > 0 else None for l in range(32)]
            for l in range(32):
                last = min(x[l][R - 1], up[l]) if l < 31 else x[l][R - 1]
                first = max(x[l][0], dn[l]) if l > 0 else x[l][0]
                x[l][R - 1], x[l][0] = last, first
    return [x[l][r] for l in range(32) for r in range(R)], mx


def test_odd_even_transposition_needs_k_phases_for_k_elements_from_either_parity():
    def oets(a, start, phases):
        a = list(a)
        for ph in range(phases):
            for i in range((start + ph) & 1, len(a) - 1, 2):
                if a[i] > a[i + 1]:
                    a[i], a[i + 1] = a[i + 1], a[i]
        return a

    for k in range(1, 8):
        for start in (0, 1):
            for perm in itertools.permutations(range(k)):
                assert oets(perm, start, k) == sorted(perm)


@pytest.mark.parametrize("nd", [1, 5, 1000, 1024, 1025, 70000, 10 ** 8, 2 ** 32 - 2])
def test_bucket_index_is_monotone_and_in_range(nd):
    mul = bucket_mul(nd)
    rng = np.random.default_rng(nd % 1000)
    docs = np.unique(np.concatenate([rng.integers(0, nd, 2000), [0, nd - 1, nd // 2]])).astype(object)
    for sh, n in ((0, 1024), (1, 512), (2, 256)):
        b = [min(((int(d) * mul) >> 32) >> sh, n - 1) for d in docs]
        assert all(b[i] <= b[i + 1] for i in range(len(b) - 1))
        assert ((int(docs[-1]) * mul) >> 32) >> sh < n  # no clamping needed for a valid doc index


def test_model_sorts_like_the_reference_order():
    rng = np.random.default_rng(1)
    random.seed(1)
    fallbacks = sorted_rows = 0
    for trial in range(150):
        R = random.choice([8, 16, 32])
        N = 32 * R
        nd = random.choice([1, 5, 1000, 1024, 1025, 70000, 10 ** 6, 10 ** 8, 2 ** 32 - 2])
        occ = random.randint(N // 2 + 1, N)
        mode = random.random()
        if mode < 0.6:  # spread over the corpus
            docs = rng.integers(0, nd, occ)
        elif mode < 0.8:  # clustered
            w = max(1, nd // 100)
            docs = rng.integers(0, w, occ) + (nd - w) // 2
        else:  # many repeats
            docs = rng.choice(rng.integers(0, nd, max(1, occ // 3)), occ)
        docs = [int(d) for d in docs]
        flat, mx = warp_bucket_sort(docs, R, nd)
        if flat is None:
            fallbacks += 1
            continue
        sorted_rows += 1
        assert flat == sorted(docs) + [SENTINEL] * (N - occ), (R, nd, occ, mx)
    assert fallbacks > 10 and sorted_rows > 50  # both outcomes are exercised
