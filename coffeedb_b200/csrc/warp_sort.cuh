// Warp-level sorts of up to 1024 32-bit keys held in registers (R per lane): the all-ascending bitonic network and the
// distribution sort used by gather_kernel (locate.cu: the doc indices of one suffix-array interval) and by the document
// listing's build (listing.cu: the doc indices / id ranks of one prefix-directory bucket).
#pragma once
#include "common.cuh"

namespace cdb {

constexpr int kWarpCap = 1024;    // occurrences one warp sorts in shared memory
constexpr int kTileWarps = 8;     // patterns per CTA tile


// Bitonic sorting network over 32*R keys held in registers, R per lane, in the all-ascending "flip + disperse"
// form: every compare-exchange puts the minimum at the lower index.  Two register layouts are used:
//   blocked  (B): lane holds ranks lane*R + r.  Strides below R are two VIMNMX per pair with no direction select,
//                 strides >= R cost one SHFL + a min/max chosen by a lane predicate per key (3 issue slots).
//   striped  (T): lane holds ranks r*32 + lane.  Strides >= 32 are register pairs, and a flip pairs register r of
//                 lane l with register r ^ (h/32-1) of lane l ^ 31: one SHFL + ONE VIMNMX per key.
// The merge phases with three or more lane-crossing strides (h >= 256) switch to T through shared memory (padded,
// conflict-free), do those strides there and come back for the strides below R.  The kernel is bound by the ALU pipe,
// and the round trip trades 2 ALU instructions per key and stride for 4 LDS/STS per key and phase.
// (MATCH.ANY-based multi-split radix sorting was measured first and is slower on B200: profiles/README.md.)
template <int R>
__device__ __forceinline__ int tr_addr(int i) { return (i / R) * (R + 1) + (i % R); }

template <int R>
__device__ __forceinline__ void warp_bitonic_regs(u32 (&x)[R], int lane, u32* tbuf) {
    constexpr int N = 32 * R;
#pragma unroll
    for (int h = 2; h <= N; h <<= 1) {
        const bool striped = R >= 16 && h >= 256;  // this phase does its strides >= 32 in layout T
        if (striped) {
            // ---- B -> T
#pragma unroll
            for (int r = 0; r < R; ++r) tbuf[lane * (R + 1) + r] = x[r];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = tbuf[tr_addr<R>(r * 32 + lane)];
            __syncwarp();
            // flip: (r, lane) <-> (r ^ (h/32-1), lane ^ 31); the lower rank is the one with the smaller register
            {
                const int rm = h / 32 - 1;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int rp = r ^ rm;
                    if (r < rp) {
                        const u32 a = x[r], b = x[rp];
                        const u32 ob = __shfl_xor_sync(0xffffffffu, b, 31);
                        const u32 oa = __shfl_xor_sync(0xffffffffu, a, 31);
                        x[r] = min(a, ob);
                        x[rp] = max(b, oa);
                    }
                }
            }
            // disperse strides h/4 .. 32: register pairs
#pragma unroll
            for (int j = h >> 2; j >= 32; j >>= 1) {
                const int rj = j / 32;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rj) == 0) {
                        const u32 a = x[r], b = x[r | rj];
                        x[r] = min(a, b);
                        x[r | rj] = max(a, b);
                    }
                }
            }
            // ---- T -> B
#pragma unroll
            for (int r = 0; r < R; ++r) tbuf[tr_addr<R>(r * 32 + lane)] = x[r];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < R; ++r) x[r] = tbuf[lane * (R + 1) + r];
            __syncwarp();
        } else {
            // flip: i <-> i ^ (h-1)
            if (h <= R) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int p = r ^ (h - 1);
                    if (r < p) {
                        const u32 a = x[r], b = x[p];
                        x[r] = min(a, b);
                        x[p] = max(a, b);
                    }
                }
            } else {
                const int lm = h / R - 1;
                const bool keep_min = (lane & (h / (2 * R))) == 0;
                if (R == 1) {
                    const u32 o = __shfl_xor_sync(0xffffffffu, x[0], lm);
                    x[0] = keep_min ? min(x[0], o) : max(x[0], o);
                } else {
#pragma unroll
                    for (int r = 0; r < R / 2; ++r) {
                        const u32 a = x[r], b = x[R - 1 - r];
                        const u32 oa = __shfl_xor_sync(0xffffffffu, b, lm);
                        const u32 ob = __shfl_xor_sync(0xffffffffu, a, lm);
                        x[r] = keep_min ? min(a, oa) : max(a, oa);
                        x[R - 1 - r] = keep_min ? min(b, ob) : max(b, ob);
                    }
                }
            }
        }
        // disperse: i <-> i ^ j for the remaining strides, layout B
#pragma unroll
        for (int j = h >> 2; j > 0; j >>= 1) {
            if (striped && j >= 32) continue;  // done above
            if (j >= R) {
                const int lj = j / R;
                const bool keep_min = (lane & lj) == 0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const u32 o = __shfl_xor_sync(0xffffffffu, x[r], lj);
                    x[r] = keep_min ? min(x[r], o) : max(x[r], o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & j) == 0) {
                        const u32 a = x[r], b = x[r | j];
                        x[r] = min(a, b);
                        x[r | j] = max(a, b);
                    }
                }
            }
        }
    }
}

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 5); }

// Distribution sort for intervals whose documents are spread over the corpus (the usual case: a pattern's hits fall
// into unrelated documents).  N = 32*R buckets by doc * N / nd (monotone in doc, so bucket order is doc order):
// count with shared-memory atomics, scan, scatter; the array is then sorted up to the order inside each bucket, and
// because neighbouring buckets are already in order, `largest bucket size` phases of an odd-even transposition over the
// whole array (unconditional compare-exchanges of neighbours, in registers) finish it.  About 650 warp instructions
// for 1024 keys whose largest bucket holds 6, against about 2 400 for the sorting network.  Returns false — with x[]
// untouched — when some bucket holds more than kBucketMax keys (clustered or repeated documents): the caller then
// runs the sorting network.  In: x[r] = key of element r*32 + lane (0xffffffff beyond occ).  Out: blocked layout,
// lane holds ranks lane*R .. lane*R + R-1, like warp_bitonic_regs.
//   s_out: 33*R words (pad_idx layout), s_cnt: 33*R words, 16-byte aligned.  bucket_mul = floor(2^32 * 1024 / nd), saturated.
constexpr u32 kBucketMax = 24;
#ifndef CDB_BUCKET_MIN_R
#define CDB_BUCKET_MIN_R 4
#endif
// intervals of up to 32 * R keys with R below this always take the sorting network.  4 (intervals of 65..128 keys sort by
// distribution too) took the shard-sized gather from 0.82 to 0.71 ms per 10^6 rows of ~84 entries; it was 8 in round 1.
constexpr int kBucketMinR = CDB_BUCKET_MIN_R;

template <int R>
__device__ __forceinline__ bool warp_bucket_sort(u32 (&x)[R], int occ, u32 bucket_mul, u32* s_out, u32* s_cnt, int lane) {
    constexpr int N = 32 * R;
    constexpr int SH = R == 32 ? 0 : R == 16 ? 1 : R == 8 ? 2 : R == 4 ? 3 : R == 2 ? 4 : 5;  // 1024 / N
    static_assert(R >= 2 && R <= 32, "bucket sort: 2 <= R <= 32");
    auto bucket = [&](u32 doc) { return pad_idx((int)min(__umulhi(doc, bucket_mul) >> SH, (u32)(N - 1))); };
    // counter of bucket b lives at pad_idx(b) < 33*R
#pragma unroll
    for (int t = 0; t < (33 * R / 4 + 31) / 32; ++t)
        if (t * 32 + lane < 33 * R / 4) reinterpret_cast<uint4*>(s_cnt)[t * 32 + lane] = make_uint4(0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r * 32 + lane < occ) atomicAdd(&s_cnt[bucket(x[r])], 1u);
    __syncwarp();
    // exclusive scan of the counters; lane owns buckets lane*R .. lane*R + R-1 (conflict-free through the padding)
    u32* mine = s_cnt + lane * R + ((lane * R) >> 5);
    u32 sum = 0, mx = 0;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const u32 c = mine[j];
        sum += c;
        mx = max(mx, c);
    }
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (mx > kBucketMax) return false;
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    u32 run = incl - sum;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const u32 c = mine[j];
        mine[j] = run;
        run += c;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r * 32 + lane < occ) {
            const u32 p = atomicAdd(&s_cnt[bucket(x[r])], 1u);
            s_out[pad_idx((int)p)] = x[r];
        }
    }
    __syncwarp();
    {
        const u32* src = s_out + lane * R + ((lane * R) >> 5);
#pragma unroll
        for (int r = 0; r < R; ++r) x[r] = lane * R + r < occ ? src[r] : 0xffffffffu;
    }
    __syncwarp();  // s_out is the caller's s_doc: all reads done before it is written again
    if (mx >= 2) {
        for (u32 ph = 0; ph < mx; ph += 2) {
            // even phase: pairs (2i, 2i+1), all inside a lane (R is even)
#pragma unroll
            for (int r = 0; r + 1 < R; r += 2) {
                const u32 a = x[r], b = x[r + 1];
                x[r] = min(a, b);
                x[r + 1] = max(a, b);
            }
            // odd phase: pairs (2i+1, 2i+2); the last key of a lane pairs with the first key of the next lane
#pragma unroll
            for (int r = 1; r + 1 < R; r += 2) {
                const u32 a = x[r], b = x[r + 1];
                x[r] = min(a, b);
                x[r + 1] = max(a, b);
            }
            const u32 up = __shfl_down_sync(0xffffffffu, x[0], 1);
            const u32 dn = __shfl_up_sync(0xffffffffu, x[R - 1], 1);
            const u32 last = lane < 31 ? min(x[R - 1], up) : x[R - 1];
            const u32 first = lane > 0 ? max(x[0], dn) : x[0];
            x[R - 1] = last;
            x[0] = first;
        }
    }
    return true;
}

// s_doc and s_pos of one warp (u32, padded) for intervals of up to 32 * MAXR occurrences
template <int MAXR>
constexpr size_t warp_smem_bytes() {
    return ((size_t)32 * MAXR + 32) * 4 + ((size_t)32 * MAXR + 64) * 4;
}

}  // namespace cdb
