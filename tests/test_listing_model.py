"""Host models of two device-side pieces of the document listing (coffeedb_b200/csrc/listing.cu) — checked here without a
GPU, so that the index arithmetic of the kernels is pinned by the CPU suite too:

  * listing_emit_kernel's run-length path: a warp walks a listed row 32 entries at a time and writes one (id, count) pair per
    run from ballots (run heads at lanes <= lane, carried head position, heads before the chunk) — restated lane by lane and
    compared with the plain run-length encoding of the row, which is what string_index::query reports
    (src/index.cpp:305-322: equal neighbours after the sort are one document, the run length is its count);
  * the bucket tag in the high bits of a prefix-directory entry (locate.cuh: kPtRank, kPtCountShift, kPtCountMask) and the
    per-pattern word the search derives from it (kPreListed | kPreRepeat | row length)."""
import numpy as np

K_PT_RANK = (1 << 48) - 1
K_PT_COUNT_SHIFT = 48
K_PT_COUNT_MASK = 0x7FF
K_PRE_LISTED = 1 << 63
K_PRE_REPEAT = 1 << 62
K_PRE_COUNT = K_PRE_REPEAT - 1


def emit_rle_model(values):
    """listing_emit_kernel, rows with a repeated document: returns the (value, count) pairs it stores, by output slot."""
    occ = len(values)
    out = {}
    carry_head = 0
    heads_before = 0
    last_v = 0
    for j0 in range(0, occ, 32):
        lanes = range(32)
        v = [values[j0 + l] if j0 + l < occ else 0 for l in lanes]
        valid = [j0 + l < occ for l in lanes]
        prev = [last_v] + v[:31]                                  # shfl_up, lane 0 takes the previous chunk's last value
        nxt = v[1:] + [0]                                          # shfl_down
        nvalid = [j0 + l + 1 < occ for l in lanes]
        if nvalid[31]:
            nxt[31] = values[j0 + 32]                              # lane 31 loads the first entry of the next chunk
        head = [valid[l] and (j0 + l == 0 or v[l] != prev[l]) for l in lanes]
        tail = [valid[l] and (not nvalid[l] or nxt[l] != v[l]) for l in lanes]
        hm = sum(1 << l for l in lanes if head[l])
        for l in lanes:
            le = hm & (0xFFFFFFFF >> (31 - l))
            headpos = j0 + (le.bit_length() - 1) if le else carry_head
            idx = heads_before + bin(le).count("1") - 1
            if tail[l]:
                assert idx not in out, "two runs stored in one slot"
                out[idx] = (v[l], j0 + l - headpos + 1)
        if hm:
            carry_head = j0 + (hm.bit_length() - 1)
        heads_before += bin(hm).count("1")
        last_v = v[31]
    return [out[i] for i in range(len(out))]


def rle(values):
    pairs = []
    for x in values:
        if pairs and pairs[-1][0] == x:
            pairs[-1][1] += 1
        else:
            pairs.append([x, 1])
    return [tuple(p) for p in pairs]


def test_emit_run_length_path_equals_plain_run_length_encoding():
    rng = np.random.default_rng(7)
    rows = [
        [5], [5, 5], [5, 6], [1] * 31 + [2], [1] * 32 + [2], [1] * 33, [1] * 64 + [2] * 64, list(range(100)),
        [3] * 1024, sorted(rng.integers(0, 40, size=1024).tolist()), sorted(rng.integers(0, 900, size=1000).tolist()),
    ]
    for n in (2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 500, 1023, 1024):
        for spread in (2, n // 3 + 1, 4 * n):
            rows.append(sorted(rng.integers(0, spread, size=n).tolist()))
    # runs that begin in one chunk and end several chunks later, at every alignment
    for start in (0, 1, 30, 31, 32, 33):
        for length in (1, 2, 32, 33, 64, 70, 129):
            rows.append(list(range(start)) + [10 ** 6] * length + [10 ** 6 + 1 + i for i in range(7)])
    for row in rows:
        assert emit_rle_model(row) == rle(row), row[:40]


def test_directory_tag_round_trip():
    for rank in (0, 1, 12345, (1 << 40) + 17, K_PT_RANK):
        for d, occ in ((1, 1), (1, 2), (7, 7), (1000, 1024), (1024, 1024)):
            repeat = d != occ
            entry = rank | (1 << 63) | ((1 << 62) if repeat else 0) | (d << K_PT_COUNT_SHIFT)
            assert entry < (1 << 64)
            assert entry & K_PT_RANK == rank                                   # every reader masks the rank
            assert entry >> 63 == 1
            pre = K_PRE_LISTED | (K_PRE_REPEAT if (entry >> 62) & 1 else 0) | ((entry >> K_PT_COUNT_SHIFT) & K_PT_COUNT_MASK)
            assert pre & K_PRE_COUNT == d and bool(pre & K_PRE_REPEAT) == repeat and pre & K_PRE_LISTED
    # an untagged entry (long or empty bucket, or no listing built) reads as "not listed"
    assert (123456 >> 63) == 0 and ((123456 >> K_PT_COUNT_SHIFT) & K_PT_COUNT_MASK) == 0
    # the count field holds every row length a listed bucket can have (<= 1024 suffixes) and stays clear of bits 62 / 63
    assert K_PT_COUNT_MASK >= 1024 and (K_PT_COUNT_MASK << K_PT_COUNT_SHIFT) < (1 << 62)


def test_two_phase_build_visits_every_listed_entry_exactly_once():
    """listing_build_kernel (two-phase) leaves, per directory bucket, the split points of its sorted keys at the doc-range
    boundaries in seg[c / 8][r][c % 8]; listing_translate_kernel walks items (range r, 32 buckets), spreads its lanes over the
    concatenated segments through an exclusive scan and `pos0 - excl` bases (unsigned wrap-around included) and must touch
    every entry of every listed bucket exactly once, with a key inside the item's doc range."""
    rng = np.random.default_rng(11)
    M64 = (1 << 64) - 1
    tile_warps = 8
    for nentries, nd, rshift in ((64, 1000, 8), (96, 5000, 10), (32, 300, 40)):
        nranges = max(1, -(-nd // (1 << rshift)))
        sizes = rng.integers(0, 60, size=nentries)
        sizes[rng.integers(0, nentries, size=nentries // 3)] = 0          # empty buckets
        ptab = np.concatenate([[0], np.cumsum(sizes)])
        keys = np.zeros(int(ptab[-1]), np.int64)
        seg = np.zeros(-(-nentries // tile_warps) * tile_warps * (nranges + 1), np.int64)
        for c in range(nentries):
            row = np.sort(rng.integers(0, nd, size=sizes[c]))
            keys[ptab[c]:ptab[c + 1]] = row
            for r in range(nranges + 1):                                    # first rank whose key is >= r << rshift
                seg[(c // tile_warps) * (nranges + 1) * tile_warps + r * tile_warps + c % tile_warps] = np.searchsorted(row, r << rshift)
        visited = np.zeros(len(keys), np.int64)
        ntile = -(-nentries // 32)
        for item in range(ntile * nranges):
            r, blk = divmod(item, ntile)
            lens, pos0 = [], []
            for lane in range(32):
                c = blk * 32 + lane
                ln = p0 = 0
                if c < nentries:
                    base = ((c // tile_warps) * (nranges + 1) + r) * tile_warps + c % tile_warps
                    s, e = seg[base], seg[base + tile_warps]
                    ln = int(e - s)
                    if ln:
                        p0 = int(ptab[c] + s)
                lens.append(ln)
                pos0.append(p0)
            excl = np.concatenate([[0], np.cumsum(lens)])[:32]
            s_pos = [(pos0[l] - int(excl[l])) & M64 for l in range(32)]     # u64 arithmetic on the device
            tot = sum(lens)
            for idx in range(tot):
                j = 0
                for st in (16, 8, 4, 2, 1):                                 # largest j with excl[j] <= idx
                    if excl[j + st] <= idx:
                        j += st
                p = (s_pos[j] + idx) & M64
                assert lens[j] > 0 and pos0[j] <= p < pos0[j] + lens[j]
                assert (r << rshift) <= keys[p] < ((r + 1) << rshift)
                visited[p] += 1
        assert (visited == 1).all()
