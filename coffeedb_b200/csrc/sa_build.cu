// Suffix-array construction on the device — replaces string_index::build() + parallel_sort<T>()
// (src/index.cpp:178-236, 75-128; kernels K1-K3 of SURVEY.md §2a).
//
// Algorithm (B200-first, not the reference's work-queue MSD radix):
//   round 0   every suffix gets a key holding its first S0 symbols; the alphabet is re-coded to
//             b = ceil(log2(sigma+1)) bits with 0 = end-of-document, and S0 = ceil((log2 n + 10) / b) symbols make
//             random text already almost tie-free (9 symbols = 45 bits = 6 radix passes for a-z at n = 10^10; at most
//             S = 64/b).  The reference's 257-symbol order "end-of-document < every byte" (src/index.h:66-73) is
//             exactly integer order of these keys.  (key, packed) pairs are sorted by the onesweep radix sort.
//   round r   only suffixes still tied with a neighbour stay on a worklist.  A tied group whose members have
//             ended (their remaining length < compared depth) consists of byte-identical suffixes: their
//             final order is ascending packed value (the canonical order of note N2).  Every other tied
//             suffix gets the next S symbols as its key.  The worklist is sorted by key, then stably by
//             group number, and written back in place; new ties form the next worklist.
//   chunks    when 2*(8+w) bytes per suffix do not fit the workspace, suffixes are partitioned by the top
//             12 bits of their round-0 key and the partitions are sorted one after another; keys ping-pong
//             between two buffers, values between one buffer and the partition's own range of the suffix array,
//             where the last radix pass lands (a 10 GB corpus takes 4 chunks on one 180 GB B200).
//   note N1   on corpora mixing bytes < 0x80 and >= 0x80 the sorted array is then rotated into the reference's
//             signed-radix / unsigned-leaf layout (apply_signed_radix_layout).
// The payload carried through every sort is the reference's own packed element (offset << bits1) | doc, so the
// finished array is byte-for-byte what src/index.cpp:209-215 + the sort would hold (up to note N2 ties).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "index.cuh"
#include "primitives.cuh"
#include "radix_sort.cuh"

namespace cdb {

// ---- corpus statistics ----------------------------------------------------------------------------------
__global__ void doc_stats_kernel(const i64* __restrict__ doc_off, i64 nd, unsigned long long* maxlen) {
    u64 m = 0;
    for (i64 d = (i64)blockIdx.x * blockDim.x + threadIdx.x; d < nd; d += (i64)gridDim.x * blockDim.x) {
        u64 len = (u64)(doc_off[d + 1] - doc_off[d]);
        m = len > m ? len : m;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        u64 t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(maxlen, (unsigned long long)m);
}

__global__ void byte_presence_kernel(const u8* __restrict__ text, i64 n, u32* __restrict__ present) {
    __shared__ u32 sp[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sp[i] = 0;
    __syncthreads();
    const i64 nvec = n >> 4;
    const uint4* t4 = reinterpret_cast<const uint4*>(text);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (i64)gridDim.x * blockDim.x) {
        uint4 v = ld_stream_v4(t4 + i);
        u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sp[w[k] & 255] = 1;  // benign race: every writer stores 1
            sp[(w[k] >> 8) & 255] = 1;
            sp[(w[k] >> 16) & 255] = 1;
            sp[w[k] >> 24] = 1;
        }
    }
    if (blockIdx.x == 0)
        for (i64 i = (nvec << 4) + threadIdx.x; i < n; i += blockDim.x) sp[text[i]] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sp[i]) present[i] = 1;
}

// largest d in [lo, hi] with doc_off[d] <= g   (doc_off[lo] <= g is guaranteed by the caller)
__device__ __forceinline__ i64 doc_of(const i64* __restrict__ doc_off, i64 lo, i64 hi, i64 g) {
    while (lo < hi) {
        i64 mid = lo + (hi - lo + 1) / 2;
        if (__ldg(doc_off + mid) <= g)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// ---- K1 + key extraction --------------------------------------------------------------------------------
// tile_doc[t] = the document that holds text position t * EX_TILE (largest d with doc_off[d] <= position; empty
// documents share an offset with their successor and are skipped this way), for t = 0 .. ntiles.  One parallel
// lower bound per tile: neighbouring tiles follow the same path, the probes hit L1/L2.
constexpr int EX_THREADS = 256;
constexpr int EX_IPT = 8;
constexpr int EX_TILE = EX_THREADS * EX_IPT;
constexpr int EX_MAXS = 32;
constexpr int EX_HIST = 4096;  // chunk planning histogram over the top 12 key bits

__global__ void tile_doc_kernel(const i64* __restrict__ doc_off, i64 nd, i64 n, i64 ntiles, i64* __restrict__ tile_doc) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    const i64 pos = t * EX_TILE;
    tile_doc[t] = pos >= n ? nd - 1 : doc_of(doc_off, 0, nd - 1, pos);
}

// Persistent CTAs walk tiles of 1024 consecutive text positions.  Per tile: the bytes (plus S-1 lookahead) arrive as
// 16-byte vector loads and are re-coded into shared memory once; the document of every position comes from a
// block-wide scan over the document starts that fall into the tile (no per-position search); each thread then
// assembles the keys of 4 positions.  MODE 0: histogram of the top key bits (chunk planning; shared-memory
// histogram, flushed once per CTA).  MODE 1: all positions, output index = position.  MODE 2: positions whose
// bucket lies in [blo, bhi): compacted in shared memory, appended through a CTA-aggregated atomic cursor and
// written as contiguous runs.
template <typename P, int MODE>
__global__ void __launch_bounds__(EX_THREADS) extract_kernel(const u8* __restrict__ text, const i64* __restrict__ doc_off,
                                                             const i64* __restrict__ tile_doc, i64 nd, i64 n, i64 ntiles,
                                                             SymTab tab, int b, int S, int bits1, int cbshift, u32 blo,
                                                             u32 bhi, u64* __restrict__ keys, P* __restrict__ vals,
                                                             unsigned long long* cursor, unsigned long long* bucket_hist) {
    __shared__ u16 s_tab[256];
    __shared__ __align__(16) u16 s_sym[EX_TILE + 2 * EX_MAXS];
    __shared__ u32 s_cnt[EX_TILE];
    __shared__ u32 s_bits[(EX_TILE + 2 * EX_MAXS) / 32 + 2];  // bit p: a document (or the text) ends right before t0 + p
    __shared__ u32 s_ws[32];
    __shared__ u64 s_base;
    __shared__ u64 s_key[MODE == 2 ? EX_TILE : 1];
    __shared__ P s_val[MODE == 2 ? EX_TILE : 1];
    __shared__ u32 s_hist[MODE == 0 ? EX_HIST : 1];
    const int tid = threadIdx.x;
    s_tab[tid] = tab.sym[tid];
    if (MODE == 0)
        for (int i = tid; i < EX_HIST; i += EX_THREADS) s_hist[i] = 0;
    __syncthreads();
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 t0 = tile * EX_TILE;
        const i64 d0 = __ldg(tile_doc + tile), d1 = __ldg(tile_doc + tile + 1);
        // symbols of text[t0, t0 + EX_TILE + S): every thread re-codes one aligned 4-byte word (text is 16-byte aligned
        // and padded) and stores its 4 symbols with one 8-byte shared store
        for (int w = tid; w * 4 < EX_TILE + S; w += EX_THREADS) {
            u32 v = 0;  // bytes past the padded end are never part of a key
            if (t0 + (i64)w * 4 + 4 <= n + kTextPad) v = ld_stream_u32(reinterpret_cast<const u32*>(text + t0) + w);
            uint2 o;
            o.x = (u32)s_tab[v & 255] | ((u32)s_tab[(v >> 8) & 255] << 16);
            o.y = (u32)s_tab[(v >> 16) & 255] | ((u32)s_tab[v >> 24] << 16);
            *reinterpret_cast<uint2*>(&s_sym[w * 4]) = o;
        }
#pragma unroll
        for (int r = 0; r < EX_IPT; r += 4) *reinterpret_cast<uint4*>(&s_cnt[tid * EX_IPT + r]) = make_uint4(0, 0, 0, 0);
        if (tid < (int)(sizeof(s_bits) / 4)) s_bits[tid] = 0;
        __syncthreads();
        // document starts inside (t0, t0 + EX_TILE + S): documents d0+1 .. (doc_off is non-decreasing; entry nd = n is
        // the end of the text).  Counts feed the doc-index scan, the bit mask tells every position how far its
        // document reaches without touching global memory again.
        for (i64 d = d0 + 1 + tid; d <= nd; d += EX_THREADS) {
            const i64 p = __ldg(doc_off + d) - t0;
            if (p >= EX_TILE + S) break;
            if (p > 0) {
                if (p < EX_TILE && d <= d1) atomicAdd(&s_cnt[p], 1u);
                atomicOr(&s_bits[p >> 5], 1u << (p & 31));
            }
        }
        __syncthreads();
        // thread-contiguous scan: position li = tid * EX_IPT + r
        u32 c[EX_IPT];
        u32 csum = 0;
#pragma unroll
        for (int r = 0; r < EX_IPT; r += 4) {
            const uint4 cv = *reinterpret_cast<const uint4*>(&s_cnt[tid * EX_IPT + r]);
            c[r] = csum + cv.x;
            c[r + 1] = c[r] + cv.y;
            c[r + 2] = c[r + 1] + cv.z;
            c[r + 3] = c[r + 2] + cv.w;
            csum = c[r + 3];
        }
        u32 tot;
        const u64 dbase = (u64)d0 + prim::block_exclusive_scan_u32(csum, &tot, s_ws);
        u64 key[EX_IPT];
        P val[EX_IPT];
        bool sel[EX_IPT];
        u32 nsel = 0;
        // sliding window over the thread's 4 consecutive positions: W = the S symbols from the position on, ignoring
        // document ends; the key keeps the first min(S, remaining) of them
        const u64 keymask = b * S >= 64 ? ~0ull : ((1ull << (b * S)) - 1);
        u64 W = 0;
        for (int j = 0; j < S - 1; ++j) W = (W << b) | (u64)s_sym[tid * EX_IPT + j];
#pragma unroll
        for (int r = 0; r < EX_IPT; ++r) {
            const int li = tid * EX_IPT + r;
            const i64 g = t0 + li;
            W = ((W << b) | (u64)s_sym[li + S - 1]) & keymask;
            sel[r] = false;
            key[r] = 0;
            val[r] = 0;
            if (g < n) {
                // remaining length of the document from here, capped at 33: first set bit among positions li+1 .. li+32
                const int p1 = li + 1;
                const u32 win = __funnelshift_r(s_bits[p1 >> 5], s_bits[(p1 >> 5) + 1], p1 & 31);
                const int rem = win ? __ffs(win) : 33;
                u64 k = W;
                if (rem < S) {  // the document ends inside the window (rare): drop the symbols past its end
                    const int sh = b * (S - rem);
                    k = (W >> sh) << sh;
                }
                key[r] = k;
                if (MODE == 1) {
                    sel[r] = true;
                } else {
                    const u32 bucket = (u32)(k >> cbshift);
                    if (MODE == 0)
                        atomicAdd(&s_hist[bucket], 1u);
                    else
                        sel[r] = bucket >= blo && bucket < bhi;
                }
                if (sel[r]) {  // the packed element needs the document's start: only selected positions load it
                    const i64 d = (i64)(dbase + c[r]);
                    val[r] = (P)(((u64)(g - __ldg(doc_off + d)) << bits1) | (u64)d);
                    ++nsel;
                }
            }
        }
        if (MODE == 1 && t0 + EX_TILE <= n) {
            // thread-contiguous: 8 keys = 64 bytes, 8 values = 32/64 bytes per thread, 16-byte stores
            ulonglong2* kd = reinterpret_cast<ulonglong2*>(keys + t0 + tid * EX_IPT);
#pragma unroll
            for (int r = 0; r < EX_IPT; r += 2) kd[r / 2] = make_ulonglong2(key[r], key[r + 1]);
            if (sizeof(P) == 4) {
                uint4* vd = reinterpret_cast<uint4*>(vals + t0 + tid * EX_IPT);
#pragma unroll
                for (int r = 0; r < EX_IPT; r += 4) vd[r / 4] = make_uint4((u32)val[r], (u32)val[r + 1], (u32)val[r + 2], (u32)val[r + 3]);
            } else {
                ulonglong2* vd = reinterpret_cast<ulonglong2*>(vals + t0 + tid * EX_IPT);
#pragma unroll
                for (int r = 0; r < EX_IPT; r += 2) vd[r / 2] = make_ulonglong2((u64)val[r], (u64)val[r + 1]);
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int r = 0; r < EX_IPT; ++r) {
                const i64 g = t0 + tid * EX_IPT + r;
                if (sel[r]) {
                    keys[g] = key[r];
                    vals[g] = val[r];
                }
            }
        } else if (MODE == 2) {
            u32 stot;
            const u32 ex = prim::block_exclusive_scan_u32(nsel, &stot, s_ws);
            if (tid == 0) s_base = stot ? atomicAdd(cursor, (unsigned long long)stot) : 0;
            u32 o = ex;
#pragma unroll
            for (int r = 0; r < EX_IPT; ++r) {
                if (sel[r]) {
                    s_key[o] = key[r];
                    s_val[o] = val[r];
                    ++o;
                }
            }
            __syncthreads();
            const u64 base = s_base;
            for (u32 i = tid; i < stot; i += EX_THREADS) {
                keys[base + i] = s_key[i];
                vals[base + i] = s_val[i];
            }
        }
        __syncthreads();  // s_sym / s_cnt / staging are re-used by the next tile
    }
    if (MODE == 0) {
        __syncthreads();
        for (int i = tid; i < EX_HIST; i += EX_THREADS)
            if (s_hist[i]) atomicAdd(bucket_hist + i, (unsigned long long)s_hist[i]);
    }
}

// ---- tie detection + ordered compaction --------------------------------------------------------------------
// An element is "tied" when its key (and group, in refinement rounds) equals a neighbour's; the first element of a
// tied run is its head.  Tied elements are compacted in order into the next worklist.  A CTA covers 2048 consecutive
// elements, warp w the 256 elements [w*256, (w+1)*256), lane l the elements r*32 + l of them (coalesced loads);
// neighbours come from shuffles, positions from ballots.
constexpr int TC_THREADS = 256;
constexpr int TC_IPT = 8;
constexpr int TC_TILE = TC_THREADS * TC_IPT;

template <bool HasGid>
__device__ __forceinline__ void tie_ballots(const u64* __restrict__ keys, const u32* __restrict__ gid, u64 m, u64 wbase,
                                            int lane, u32 bt[TC_IPT], u32 bh[TC_IPT]) {
    u64 k[TC_IPT];
    u32 g[TC_IPT];
#pragma unroll
    for (int r = 0; r < TC_IPT; ++r) {
        const u64 i = wbase + r * 32 + lane;
        k[r] = i < m ? ld_stream_u64(keys + i) : 0;
        g[r] = (HasGid && i < m) ? __ldg(gid + i) : 0;
    }
    // the elements just outside the warp's range
    u64 kb = 0, ka = 0;
    u32 gb = 0, ga = 0;
    const bool has_b = wbase > 0 && wbase <= m, has_a = wbase + 32 * TC_IPT < m;
    if (lane == 0 && has_b) {
        kb = keys[wbase - 1];
        gb = HasGid ? gid[wbase - 1] : 0;
    }
    if (lane == 31 && has_a) {
        ka = keys[wbase + 32 * TC_IPT];
        ga = HasGid ? gid[wbase + 32 * TC_IPT] : 0;
    }
#pragma unroll
    for (int r = 0; r < TC_IPT; ++r) {
        const u64 i = wbase + r * 32 + lane;
        u64 kp = __shfl_up_sync(0xffffffffu, k[r], 1), kn = __shfl_down_sync(0xffffffffu, k[r], 1);
        u32 gp = __shfl_up_sync(0xffffffffu, g[r], 1), gn = __shfl_down_sync(0xffffffffu, g[r], 1);
        const u64 kp0 = r > 0 ? __shfl_sync(0xffffffffu, k[r > 0 ? r - 1 : 0], 31) : kb;
        const u32 gp0 = r > 0 ? __shfl_sync(0xffffffffu, g[r > 0 ? r - 1 : 0], 31) : gb;
        const u64 kn31 = r + 1 < TC_IPT ? __shfl_sync(0xffffffffu, k[r + 1 < TC_IPT ? r + 1 : r], 0) : ka;
        const u32 gn31 = r + 1 < TC_IPT ? __shfl_sync(0xffffffffu, g[r + 1 < TC_IPT ? r + 1 : r], 0) : ga;
        if (lane == 0) {
            kp = kp0;
            gp = gp0;
        }
        if (lane == 31) {
            kn = kn31;
            gn = gn31;
        }
        const bool valid = i < m;
        const bool prev_ok = i > 0 && (r > 0 || lane > 0 || has_b);
        const bool next_ok = i + 1 < m && (r + 1 < TC_IPT || lane < 31 || has_a);
        const bool eq_prev = valid && prev_ok && kp == k[r] && (!HasGid || gp == g[r]);
        const bool eq_next = valid && next_ok && kn == k[r] && (!HasGid || gn == g[r]);
        const bool tied = eq_prev || eq_next;
        bt[r] = __ballot_sync(0xffffffffu, tied);
        bh[r] = __ballot_sync(0xffffffffu, tied && !eq_prev);
    }
}

// packed count: low 32 bits = tied elements, high 32 bits = group heads among them
template <bool HasGid>
__global__ void __launch_bounds__(TC_THREADS) ties_count_kernel(const u64* __restrict__ keys, const u32* __restrict__ gid,
                                                                u64 m, u64* __restrict__ bsum) {
    __shared__ u64 ws[TC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 bt[TC_IPT], bh[TC_IPT];
    tie_ballots<HasGid>(keys, gid, m, (u64)blockIdx.x * TC_TILE + (u64)warp * 32 * TC_IPT, lane, bt, bh);
    u64 s = 0;
#pragma unroll
    for (int r = 0; r < TC_IPT; ++r) s += (u64)__popc(bt[r]) + ((u64)__popc(bh[r]) << 32);
    if (lane == 0) ws[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 t = 0;
        for (int w = 0; w < TC_THREADS / 32; ++w) t += ws[w];
        bsum[blockIdx.x] = t;
    }
}

template <bool HasGid, typename P>
__global__ void __launch_bounds__(TC_THREADS) ties_write_kernel(const u64* __restrict__ keys, const u32* __restrict__ gid,
                                                                const u32* __restrict__ widx_in, const P* __restrict__ pay_in,
                                                                u64 m, const u64* __restrict__ bsum,
                                                                u32* __restrict__ widx_out, P* __restrict__ pay_out,
                                                                u32* __restrict__ gid_out) {
    __shared__ u64 ws[TC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 wbase = (u64)blockIdx.x * TC_TILE + (u64)warp * 32 * TC_IPT;
    u32 bt[TC_IPT], bh[TC_IPT];
    tie_ballots<HasGid>(keys, gid, m, wbase, lane, bt, bh);
    u64 s = 0;
#pragma unroll
    for (int r = 0; r < TC_IPT; ++r) s += (u64)__popc(bt[r]) + ((u64)__popc(bh[r]) << 32);
    if (lane == 0) ws[warp] = s;
    __syncthreads();
    u64 ex = bsum[blockIdx.x];
    for (int w = 0; w < warp; ++w) ex += ws[w];
    u32 pos = (u32)ex, heads = (u32)(ex >> 32);
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < TC_IPT; ++r) {
        if ((bt[r] >> lane) & 1u) {
            const u64 i = wbase + r * 32 + lane;
            const u32 o = pos + __popc(bt[r] & lt);
            const u32 h = heads + __popc(bh[r] & (lt | (1u << lane)));  // heads up to and including this element
            widx_out[o] = HasGid ? widx_in[i] : (u32)i;
            pay_out[o] = pay_in[i];
            gid_out[o] = h - 1;
        }
        pos += __popc(bt[r]);
        heads += __popc(bh[r]);
    }
}

// ---- refinement keys ---------------------------------------------------------------------------------------
// depth = number of symbols already known equal inside the element's group.
template <typename P>
__global__ void nextkey_kernel(const P* __restrict__ pay, u64 m, const u8* __restrict__ text,
                               const i64* __restrict__ doc_off, SymTab tab, int b, int S, int bits1, u64 mask,
                               i64 depth, u64* __restrict__ keys, u32* __restrict__ iota) {
    __shared__ u16 s_tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = tab.sym[i];
    __syncthreads();
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    u64 p = (u64)pay[j];
    i64 d = (i64)(p & mask);
    i64 off = (i64)(p >> bits1);
    i64 ds = __ldg(doc_off + d), de = __ldg(doc_off + d + 1);
    i64 rem = de - ds - off;
    u64 k;
    if (rem < depth) {
        k = p;  // byte-identical group: final order is ascending packed value (note N2)
    } else {
        i64 g = ds + off + depth;
        i64 left = rem - depth;
        int lim = left < (i64)S ? (int)left : S;
        k = 0;
        for (int q = 0; q < S; ++q) k = (k << b) | (u64)(q < lim ? s_tab[__ldg(text + g + q)] : 0);
    }
    keys[j] = k;
    iota[j] = (u32)j;
}

__global__ void gather_gid_kernel(const u32* __restrict__ gid, const u32* __restrict__ perm, u64 m,
                                  u64* __restrict__ gkeys, u32* __restrict__ iota) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    gkeys[j] = gid[perm[j]];
    iota[j] = (u32)j;
}

// perm2[j] = position in key order (ks/ps) of the element that belongs at worklist slot j
template <typename P>
__global__ void apply_perm_kernel(const u32* __restrict__ perm2, const u64* __restrict__ ks, const u32* __restrict__ ps,
                                  const P* __restrict__ pay, const u32* __restrict__ widx, u64 m,
                                  u64* __restrict__ key2, P* __restrict__ pay2, P* __restrict__ chunk_vals) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    u32 q = perm2[j];
    P v = pay[ps[q]];
    key2[j] = ks[q];
    pay2[j] = v;
    chunk_vals[widx[j]] = v;
}

// ---- prefix doubling over the ties that survive the extension rounds -------------------------------------------------------
// Key extension adds S symbols per round: two byte-identical 64 KB documents would need ~8 000 rounds.  After
// kExtensionRounds of them a chunk hands its remaining ties over; once every chunk is in place the inverse suffix
// array rank[text position] is built (own rank for settled suffixes, rank of the group's first member for tied ones)
// and the groups are refined by rank doubling: suffixes that share their first h symbols are ordered by the rank of the
// suffix h symbols further on (Manber-Myers / Larsson-Sadakane), h doubling per level — O(log maxlen) levels.  The sorts,
// the scatter into the array and the tie detection are the extension rounds' own.
constexpr int kExtensionRounds = 4;

template <typename P>
struct PendingTies {
    DevBuf<u32> widx, gid;  // chunk-relative suffix-array slot and group number of every tied suffix, in array order
    DevBuf<P> pay;          // its packed element
    u64 wm = 0, ngroups = 0;
    P* vals = nullptr;      // the chunk's range of the suffix array
    i64 base = 0;           // rank of vals[0] in the whole array
};

template <typename P>
__device__ __forceinline__ i64 text_pos(P packed, u64 mask, int bits1, const i64* __restrict__ doc_off, i64* rem) {
    const u64 p = (u64)packed;
    const i64 d = (i64)(p & mask), off = (i64)(p >> bits1);
    const i64 ds = __ldg(doc_off + d), de = __ldg(doc_off + d + 1);
    *rem = de - ds - off;
    return ds + off;
}

template <typename P>
__global__ void isa_init_kernel(const P* __restrict__ sa, i64 n, u64 mask, int bits1, const i64* __restrict__ doc_off,
                                u64* __restrict__ rank) {
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) {
        i64 rem;
        rank[text_pos<P>(sa[j], mask, bits1, doc_off, &rem)] = (u64)j;
    }
}

// head[j] = 1 when worklist slot j starts a group (same group number AND same key as its predecessor = same group)
__global__ void group_heads_kernel(const u32* __restrict__ gid, const u64* __restrict__ keys, u64 m, u8* __restrict__ head) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    head[j] = (j == 0 || gid[j] != gid[j - 1] || (keys && keys[j] != keys[j - 1])) ? 1 : 0;
}
__global__ void group_first_kernel(const u8* __restrict__ head, const u64* __restrict__ ord, u64 m, u32* __restrict__ first) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m && head[j]) first[ord[j]] = (u32)j;
}
// rank of every worklist member = rank of its group's first member
template <typename P>
__global__ void group_rank_kernel(const u8* __restrict__ head, const u64* __restrict__ ord, const u32* __restrict__ first,
                                  const u32* __restrict__ widx, const P* __restrict__ pay, u64 m, i64 base, u64 mask, int bits1,
                                  const i64* __restrict__ doc_off, u64* __restrict__ rank) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const u64 g = ord[j] + head[j] - 1;
    i64 rem;
    rank[text_pos<P>(pay[j], mask, bits1, doc_off, &rem)] = (u64)base + widx[first[g]];
}

// doubling key of a tied suffix whose group shares `depth` symbols: ended before that -> its group is byte-identical and
// the canonical order is the packed value (note N2); ended exactly there -> first (end-of-document is the smallest
// symbol); otherwise 1 + rank of the suffix `depth` symbols further on (same document, so the rank exists)
template <typename P>
__global__ void rankkey_kernel(const P* __restrict__ pay, u64 m, const i64* __restrict__ doc_off, int bits1, u64 mask, i64 depth,
                               const u64* __restrict__ rank, u64* __restrict__ keys, u32* __restrict__ iota) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    i64 rem;
    const i64 pos = text_pos<P>(pay[j], mask, bits1, doc_off, &rem);
    keys[j] = rem < depth ? (u64)pay[j] : (rem == depth ? 0ull : 1ull + rank[pos + depth]);
    iota[j] = (u32)j;
}

// ---- host driver ------------------------------------------------------------------------------------------------
static int bits_for(u64 v) {  // number of bits needed to represent values 0..v
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

// the reference's width rule (src/index.cpp:183-194): all-ones masks grown until they cover nd / maxlen
static int ones_needed(u64 v) {
    u64 m = 1;
    int b = 1;
    while (m < v) {
        m = (m << 1) + 1;
        ++b;
        if (b == 64) break;
    }
    return b;
}

struct BuildTimers {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sort_events;
    void begin(cudaStream_t st) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
        sort_events.push_back({a, b});
    }
    void end(cudaStream_t st) { cudaEventRecord(sort_events.back().second, st); }
    double total_ms() {
        double t = 0;
        for (auto& e : sort_events) {
            float ms = 0;
            cudaEventSynchronize(e.second);
            cudaEventElapsedTime(&ms, e.first, e.second);
            t += ms;
            cudaEventDestroy(e.first);
            cudaEventDestroy(e.second);
        }
        sort_events.clear();
        return t;
    }
};

template <typename P>
struct ChunkSorter {
    Index& ix;
    const SymTab& tab;
    int b, S0, S;  // S0 symbols in the round-0 key, S in every refinement key
    cudaStream_t st;
    BuildTimers& timers;

    // Sorts the m (key, packed) pairs in (k[0], v[0]) completely; returns where the finished suffix-array slice is:
    // `dest` when given (the last radix pass scatters straight into it), else one of v[0] / v[1].
    P* run(u64* k[2], P* v[2], u64 m, P* dest = nullptr, std::vector<PendingTies<P>>* pending = nullptr, i64 base = 0) {
        const int keybits = b * S;
        timers.begin(st);
        bool used = false;
        int c = rs::radix_sort_pairs<P>(k[0], k[1], v[0], v[1], m, 0, b * S0, st, nullptr, dest, &used);
        timers.end(st);
        u64* keys = k[c];
        P* vals = used ? dest : v[c];
        if (dest && !used && m) {  // no pass ran (a single distinct key): the input order is the result
            CDB_CUDA(cudaMemcpyAsync(dest, v[c], m * sizeof(P), cudaMemcpyDeviceToDevice, st));
            vals = dest;
        }
        // round-0 ties -> worklist
        DevBuf<u32> widx, gid;
        DevBuf<P> pay;
        u64 wm = 0, ngroups = 0;
        compact<false>(keys, nullptr, nullptr, vals, m, widx, pay, gid, wm, ngroups);
        i64 depth = S0;
        const int tiebits = ix.bits1 + ix.bits2;
        const int sortbits = keybits > tiebits ? keybits : tiebits;
        int ext_rounds = 0;
        while (wm > 0) {
            if (pending && ext_rounds == kExtensionRounds) {  // long repeats: the rest is refined by rank doubling
                pending->emplace_back();
                PendingTies<P>& pt = pending->back();
                pt.widx = std::move(widx);
                pt.gid = std::move(gid);
                pt.pay = std::move(pay);
                pt.wm = wm;
                pt.ngroups = ngroups;
                pt.vals = vals;
                pt.base = base;
                break;
            }
            ++ext_rounds;
            ix.rounds++;
            const unsigned gb = (unsigned)ceil_div((i64)wm, 256);
            DevBuf<u64> nk(wm, st), nk2(wm, st);
            DevBuf<u32> p0(wm, st), p1(wm, st);
            nextkey_kernel<P><<<gb, 256, 0, st>>>(pay.p, wm, ix.d_text, ix.d_off, tab, b, S, ix.bits1, ix.mask, depth,
                                                  nk.p, p0.p);
            CDB_LAUNCH_CHECK();
            timers.begin(st);
            int c1 = rs::radix_sort_pairs<u32>(nk.p, nk2.p, p0.p, p1.p, wm, 0, sortbits, st);
            timers.end(st);
            u64* ks = c1 ? nk2.p : nk.p;
            u32* ps = c1 ? p1.p : p0.p;
            u64* kfree = c1 ? nk.p : nk2.p;  // scratch from here on
            u32* pfree = c1 ? p0.p : p1.p;
            // stable sort of the key order by group number
            DevBuf<u64> gk(wm, st), gk2(wm, st);
            DevBuf<u32> q1(wm, st);
            gather_gid_kernel<<<gb, 256, 0, st>>>(gid.p, ps, wm, gk.p, pfree);
            CDB_LAUNCH_CHECK();
            timers.begin(st);
            int c2 = rs::radix_sort_pairs<u32>(gk.p, gk2.p, pfree, q1.p, wm, 0, bits_for(ngroups ? ngroups - 1 : 0), st);
            timers.end(st);
            u32* perm2 = c2 ? q1.p : pfree;
            DevBuf<P> pay2(wm, st);
            apply_perm_kernel<P><<<gb, 256, 0, st>>>(perm2, ks, ps, pay.p, widx.p, wm, kfree, pay2.p, vals);
            CDB_LAUNCH_CHECK();
            // new ties inside the worklist
            DevBuf<u32> widx2, gid2;
            DevBuf<P> pay3;
            u64 wm2 = 0, ng2 = 0;
            compact<true>(kfree, gid.p, widx.p, pay2.p, wm, widx2, pay3, gid2, wm2, ng2);
            widx = std::move(widx2);
            gid = std::move(gid2);
            pay = std::move(pay3);
            wm = wm2;
            ngroups = ng2;
            depth += S;
        }
        return vals;
    }

    template <bool HasGid>
    void compact(const u64* keys, const u32* gid_in, const u32* widx_in, const P* pay_in, u64 m, DevBuf<u32>& widx,
                 DevBuf<P>& pay, DevBuf<u32>& gid, u64& wm, u64& ngroups) {
        wm = 0;
        ngroups = 0;
        if (m < 2) return;
        const u64 nb = (m + TC_TILE - 1) / TC_TILE;
        DevBuf<u64> bsum(nb + 1, st);
        ties_count_kernel<HasGid><<<(unsigned)nb, TC_THREADS, 0, st>>>(keys, gid_in, m, bsum.p);
        CDB_LAUNCH_CHECK();
        prim::scan_blocksums_kernel<<<1, 1024, 0, st>>>(bsum.p, nb);
        CDB_LAUNCH_CHECK();
        u64 tot = 0;
        CDB_CUDA(cudaMemcpyAsync(&tot, bsum.p + nb, 8, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        wm = tot & 0xffffffffull;
        ngroups = tot >> 32;
        if (wm == 0) return;
        widx.alloc(wm, st);
        pay.alloc(wm, st);
        gid.alloc(wm, st);
        ties_write_kernel<HasGid, P><<<(unsigned)nb, TC_THREADS, 0, st>>>(keys, gid_in, widx_in, pay_in, m, bsum.p,
                                                                        widx.p, pay.p, gid.p);
        CDB_LAUNCH_CHECK();
    }

    // ranks of all members of one worklist (slots in array order, groups = runs of equal (gid, key)) -> rank[]
    void write_group_ranks(const PendingTies<P>& pt, const u32* widx, const u32* gid, const u64* keys, const P* pay, u64 wm,
                           u64* rank) {
        const unsigned gb = (unsigned)ceil_div((i64)wm, 256);
        DevBuf<u8> head(wm, st);
        DevBuf<u64> ord(wm + 1, st);
        DevBuf<u32> first(wm, st);
        group_heads_kernel<<<gb, 256, 0, st>>>(gid, keys, wm, head.p);
        CDB_LAUNCH_CHECK();
        prim::exclusive_scan<u8>(head.p, ord.p, wm, st);
        group_first_kernel<<<gb, 256, 0, st>>>(head.p, ord.p, wm, first.p);
        CDB_LAUNCH_CHECK();
        group_rank_kernel<P><<<gb, 256, 0, st>>>(head.p, ord.p, first.p, widx, pay, wm, pt.base, ix.mask, ix.bits1, ix.d_off, rank);
        CDB_LAUNCH_CHECK();
    }

    // Rank doubling over the ties every chunk handed over (all of them share `depth` symbols inside their groups).
    // sa = the whole suffix array, complete up to the order inside those groups.
    void finish_by_doubling(std::vector<PendingTies<P>>& pending, const P* sa, i64 depth) {
        if (pending.empty()) return;
        const i64 n = ix.n;
        BigBuf<u64> rank((size_t)n + 1);
        isa_init_kernel<P><<<num_sms() * 16, 256, 0, st>>>(sa, n, ix.mask, ix.bits1, ix.d_off, rank.p);
        CDB_LAUNCH_CHECK();
        for (PendingTies<P>& pt : pending) write_group_ranks(pt, pt.widx.p, pt.gid.p, nullptr, pt.pay.p, pt.wm, rank.p);
        const int tiebits = ix.bits1 + ix.bits2;
        const int rankbits = bits_for((u64)n + 1);
        const int sortbits = rankbits > tiebits ? rankbits : tiebits;
        bool any = true;
        while (any) {
            any = false;
            ix.rounds++;
            for (PendingTies<P>& pt : pending) {
                const u64 wm = pt.wm;
                if (wm == 0) continue;
                const unsigned gb = (unsigned)ceil_div((i64)wm, 256);
                DevBuf<u64> nk(wm, st), nk2(wm, st);
                DevBuf<u32> p0(wm, st), p1(wm, st);
                rankkey_kernel<P><<<gb, 256, 0, st>>>(pt.pay.p, wm, ix.d_off, ix.bits1, ix.mask, depth, rank.p, nk.p, p0.p);
                CDB_LAUNCH_CHECK();
                timers.begin(st);
                const int c1 = rs::radix_sort_pairs<u32>(nk.p, nk2.p, p0.p, p1.p, wm, 0, sortbits, st);
                timers.end(st);
                u64* ks = c1 ? nk2.p : nk.p;
                u32* ps = c1 ? p1.p : p0.p;
                u64* kfree = c1 ? nk.p : nk2.p;
                u32* pfree = c1 ? p0.p : p1.p;
                DevBuf<u64> gk(wm, st), gk2(wm, st);
                DevBuf<u32> q1(wm, st);
                gather_gid_kernel<<<gb, 256, 0, st>>>(pt.gid.p, ps, wm, gk.p, pfree);
                CDB_LAUNCH_CHECK();
                timers.begin(st);
                const int c2 = rs::radix_sort_pairs<u32>(gk.p, gk2.p, pfree, q1.p, wm, 0, bits_for(pt.ngroups ? pt.ngroups - 1 : 0), st);
                timers.end(st);
                u32* perm2 = c2 ? q1.p : pfree;
                DevBuf<P> pay2(wm, st);
                apply_perm_kernel<P><<<gb, 256, 0, st>>>(perm2, ks, ps, pt.pay.p, pt.widx.p, wm, kfree, pay2.p, pt.vals);
                CDB_LAUNCH_CHECK();
                // every member's new rank (the groups just split), then the ties that are left
                write_group_ranks(pt, pt.widx.p, pt.gid.p, kfree, pay2.p, wm, rank.p);
                DevBuf<u32> widx2, gid2;
                DevBuf<P> pay3;
                u64 wm2 = 0, ng2 = 0;
                compact<true>(kfree, pt.gid.p, pt.widx.p, pay2.p, wm, widx2, pay3, gid2, wm2, ng2);
                pt.widx = std::move(widx2);
                pt.gid = std::move(gid2);
                pt.pay = std::move(pay3);
                pt.wm = wm2;
                pt.ngroups = ng2;
                any = any || wm2 > 0;
            }
            depth *= 2;
            if (depth > ((i64)1 << 62)) throw Error(CDB_ERR_STATE, "suffix-array build: rank doubling did not converge");
        }
        CDB_CUDA(cudaStreamSynchronize(st));
    }
};

template <typename P>
static void build_typed(Index& ix, const SymTab& tab, int b, int S, cudaStream_t st) {
    const i64 n = ix.n;
    BuildTimers timers;
    // Round-0 key length: enough symbols that random text is already (almost) tie-free, log2(n) + 10 bits, instead of
    // all S that fit 64 bits — every 8 key bits saved is one full radix pass over the chunk.  Texts with long repeats
    // just leave more suffixes to the refinement rounds (which always use S symbols).  CDB_SA_KEY_SLACK overrides.
    int S0 = S;
    {
        int slack = 10;
        if (const char* e = getenv("CDB_SA_KEY_SLACK")) slack = atoi(e);
        int need = slack;
        while (need - slack < 62 && ((i64)1 << (need - slack)) < n) ++need;
        const int want = (need + b - 1) / b;
        if (want < S0) S0 = want < 1 ? 1 : want;
    }
    const int keybits = b * S0;
    // workspace -> chunk capacity
    size_t free_b = 0, total_b = 0;
    CDB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // one chunk: ping-pong (key, packed) buffers, the value buffer that ends up sorted becomes the suffix array.
    // several chunks: ping-pong keys + ONE value buffer — the chunk's own (still unused) suffix-array range is the
    // other value buffer, and the last radix pass lands in it.
    const size_t per_item_single = 2 * (8 + sizeof(P)) + 1;  // + look-back status share
    const size_t per_item_chunk = 2 * 8 + sizeof(P) + 1;
    const size_t sa_bytes = (size_t)n * sizeof(P);
    i64 cap;
    {
        size_t ws = ix.opt.workspace_bytes > 0 ? (size_t)ix.opt.workspace_bytes : 0;
        if (const char* e = getenv("CDB_BUILD_WORKSPACE_MB")) ws = (size_t)atoll(e) << 20;  // profiling aid: forces chunking
        size_t per_item = per_item_chunk;
        if (ws == 0) {
            size_t single = (size_t)n * per_item_single;
            if (single < (size_t)(free_b * 0.85)) {
                ws = single + per_item_single * rs::TILE;
                per_item = per_item_single;
            } else {
                ws = free_b > sa_bytes ? (size_t)((free_b - sa_bytes) * 0.80) : 0;
            }
        } else if ((size_t)n * per_item_single <= ws) {
            per_item = per_item_single;
        }
        cap = (i64)(ws / per_item);
        const i64 hard = ((i64)1 << 32) - 2 * rs::TILE;
        if (cap > hard) cap = hard;
        if (n > cap && cap < 2 * rs::TILE)
            throw Error(CDB_ERR_NOMEM, "not enough device memory for the suffix-array build workspace");
    }
    const i64 ex_tiles = ceil_div(n, EX_TILE);
    const unsigned ex_grid = (unsigned)std::min<i64>(ex_tiles, (i64)num_sms() * 8);
    DevBuf<i64> tile_doc((size_t)ex_tiles + 1, st);
    tile_doc_kernel<<<(unsigned)ceil_div(ex_tiles + 1, 256), 256, 0, st>>>(ix.d_off, ix.nd, n, ex_tiles, tile_doc.p);
    CDB_LAUNCH_CHECK();
    if (n <= cap) {
        // ---- one chunk: the sorted value buffer becomes the suffix array itself
        ix.chunks = 1;
        BigBuf<u64> k0(n), k1(n);
        BigBuf<P> v0(n), v1(n);
        extract_kernel<P, 1><<<ex_grid, EX_THREADS, 0, st>>>(ix.d_text, ix.d_off, tile_doc.p, ix.nd, n, ex_tiles, tab, b, S0,
                                                             ix.bits1, 0, 0, 0, k0.p, v0.p, nullptr, nullptr);
        CDB_LAUNCH_CHECK();
        u64* k[2] = {k0.p, k1.p};
        P* v[2] = {v0.p, v1.p};
        ChunkSorter<P> cs{ix, tab, b, S0, S, st, timers};
        std::vector<PendingTies<P>> pending;
        P* fin = cs.run(k, v, (u64)n, nullptr, &pending, 0);
        if (!pending.empty()) {
            // the spare key / value buffers are not needed any more: their memory holds the inverse array
            k0.release();
            k1.release();
            (fin == v1.p ? v0 : v1).release();
            cs.finish_by_doubling(pending, fin, (i64)S0 + (i64)kExtensionRounds * S);
        }
        CDB_CUDA(cudaStreamSynchronize(st));
        ix.d_sa = fin == v1.p ? (void*)v1.detach() : (void*)v0.detach();
        ix.sort_ms = timers.total_ms();
        return;
    }
    // ---- several chunks: partition by the top cb bits of the round-0 key
    const int cb = keybits < 12 ? keybits : 12;
    const int cbshift = keybits - cb;
    const u32 nbuckets = 1u << cb;
    DevBuf<unsigned long long> d_hist(nbuckets + 1, st);
    CDB_CUDA(cudaMemsetAsync(d_hist.p, 0, d_hist.bytes(), st));
    extract_kernel<P, 0><<<ex_grid, EX_THREADS, 0, st>>>(ix.d_text, ix.d_off, tile_doc.p, ix.nd, n, ex_tiles, tab, b, S0,
                                                         ix.bits1, cbshift, 0, 0, nullptr, nullptr, nullptr, d_hist.p);
    CDB_LAUNCH_CHECK();
    std::vector<unsigned long long> hist(nbuckets);
    CDB_CUDA(cudaMemcpyAsync(hist.data(), d_hist.p, nbuckets * 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    const auto t_a0 = std::chrono::steady_clock::now();
    BigBuf<P> sa(n);
    BigBuf<u64> k0(cap), k1(cap);
    BigBuf<P> v0(cap);
    if (getenv("CDB_DEBUG_TIMING"))
        fprintf(stderr, "[cdb] build: cudaMalloc of SA (%.1f GB) + workspace (%.1f GB) took %.1f ms\n", n * sizeof(P) / 1e9,
                cap * (16.0 + sizeof(P)) / 1e9,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a0).count());
    unsigned long long* cursor = d_hist.p + nbuckets;
    std::vector<PendingTies<P>> pending;
    i64 sa_base = 0;
    u32 blo = 0;
    while (blo < nbuckets) {
        i64 cnt = 0;
        u32 bhi = blo;
        while (bhi < nbuckets && cnt + (i64)hist[bhi] <= cap) cnt += (i64)hist[bhi++];
        if (bhi == blo) throw Error(CDB_ERR_NOMEM, "a single 12-bit key bucket exceeds the suffix-array build workspace");
        if (cnt > 0) {
            ix.chunks++;
            CDB_CUDA(cudaMemsetAsync(cursor, 0, 8, st));
            // values ping-pong between v0 and the chunk's final suffix-array range; they start in the one that makes an
            // all-passes-run sort end in the range (a skipped pass costs one copy, see radix_sort_pairs)
            P* slice = sa.p + sa_base;
            const int planned = (keybits + rs::RADIX_BITS - 1) / rs::RADIX_BITS;
            P* v[2] = {planned % 2 ? v0.p : slice, planned % 2 ? slice : v0.p};
            extract_kernel<P, 2><<<ex_grid, EX_THREADS, 0, st>>>(ix.d_text, ix.d_off, tile_doc.p, ix.nd, n, ex_tiles, tab, b, S0,
                                                                 ix.bits1, cbshift, blo, bhi, k0.p, v[0], cursor, nullptr);
            CDB_LAUNCH_CHECK();
            u64* k[2] = {k0.p, k1.p};
            ChunkSorter<P> cs{ix, tab, b, S0, S, st, timers};
            cs.run(k, v, (u64)cnt, slice, &pending, sa_base);  // the chunk lands in its final suffix-array range
            sa_base += cnt;
        }
        blo = bhi;
    }
    if (!pending.empty()) {
        CDB_CUDA(cudaStreamSynchronize(st));
        k0.release();
        k1.release();
        v0.release();
        ChunkSorter<P> cs{ix, tab, b, S0, S, st, timers};
        cs.finish_by_doubling(pending, sa.p, (i64)S0 + (i64)kExtensionRounds * S);
    }
    CDB_CUDA(cudaStreamSynchronize(st));
    ix.d_sa = (void*)sa.detach();
    ix.sort_ms = timers.total_ms();
    if (getenv("CDB_DEBUG_TIMING")) {
        const auto t_f0 = std::chrono::steady_clock::now();
        k0.release();
        k1.release();
        v0.release();
        fprintf(stderr, "[cdb] build: cudaFree of the workspace took %.1f ms; %.1f ms since the allocations\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_f0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_a0).count());
    }
}

// ---- note N1: the reference's signed-radix / unsigned-leaf layout ---------------------------------------------------
// The reference splits a group of suffixes that share a d-byte prefix by the SIGNED value of byte d (end-of-document
// first, then 0x80..0xFF, then 0x00..0x7F; src/index.h:66-73, src/index.cpp:97-125) as long as the group is larger
// than chuck_size = max(4096, n/256), and sorts smaller groups by unsigned memcmp (src/index.cpp:86-94).  Relative to
// the plainly (unsigned) sorted array this is, inside every such large group, a rotation of the block of suffixes
// whose byte d is < 0x80 behind the block whose byte d is >= 0x80.  The groups are found level by level with
// binary searches on the sorted array (<= 256 groups per level can exceed n/256), then the rotations are applied
// parents first, each as three in-place reversals.
struct N1Node {
    i64 b, e, d;
};

template <typename P>
__global__ void n1_bounds_kernel(const P* __restrict__ sa, u64 mask, int bits1, const i64* __restrict__ doc_off,
                                 const u8* __restrict__ text, const N1Node* __restrict__ nodes, int nnodes,
                                 i64* __restrict__ bounds) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnodes * 257) return;
    const N1Node nd = nodes[t / 257];
    const int c = t % 257;  // bounds[c] = first rank in [b, e) whose suffix is longer than d bytes and has byte d >= c
    i64 L = nd.b, R = nd.e;
    if (c == 256) L = R;
    while (L < R) {
        const i64 M = L + (R - L) / 2;
        const u64 x = (u64)sa[M];
        const i64 doc = (i64)(x & mask);
        const i64 pos = __ldg(doc_off + doc) + (i64)(x >> bits1) + nd.d;
        const int key = pos >= __ldg(doc_off + doc + 1) ? -1 : (int)text[pos];
        if (key >= c)
            R = M;
        else
            L = M + 1;
    }
    bounds[t] = L;
}

template <typename P>
__global__ void reverse_kernel(P* __restrict__ a, u64 len) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len / 2; i += (u64)gridDim.x * blockDim.x) {
        const P x = a[i], y = a[len - 1 - i];
        a[i] = y;
        a[len - 1 - i] = x;
    }
}

template <typename P>
static void reverse_range(P* a, u64 len, cudaStream_t st) {
    if (len < 2) return;
    const int grid = (int)std::min<i64>(ceil_div((i64)(len / 2), 256), num_sms() * 16);
    reverse_kernel<P><<<grid, 256, 0, st>>>(a, len);
    CDB_LAUNCH_CHECK();
}

template <typename P>
static void apply_signed_radix_layout(Index& ix, cudaStream_t st) {
    P* sa = reinterpret_cast<P*>(ix.d_sa);
    struct Pending {
        N1Node node;
        i64 shift;  // where the group sits now = node.b + shift (rotations of its ancestors)
    };
    struct Rot {
        i64 at, lsize, hsize;
    };
    std::vector<Pending> level{{{0, ix.n, 0}, 0}}, next;
    std::vector<Rot> rots;
    std::vector<N1Node> hnodes;
    std::vector<i64> hb;
    while (!level.empty()) {
        const int nn = (int)level.size();
        hnodes.resize(nn);
        for (int j = 0; j < nn; ++j) hnodes[j] = level[j].node;
        DevBuf<N1Node> dn(nn, st);
        DevBuf<i64> db((size_t)nn * 257, st);
        CDB_CUDA(cudaMemcpyAsync(dn.p, hnodes.data(), sizeof(N1Node) * nn, cudaMemcpyHostToDevice, st));
        n1_bounds_kernel<P><<<(unsigned)ceil_div((i64)nn * 257, 256), 256, 0, st>>>(sa, ix.mask, ix.bits1, ix.d_off, ix.d_text,
                                                                                  dn.p, nn, db.p);
        CDB_LAUNCH_CHECK();
        hb.resize((size_t)nn * 257);
        CDB_CUDA(cudaMemcpyAsync(hb.data(), db.p, hb.size() * 8, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        next.clear();
        for (int j = 0; j < nn; ++j) {
            const i64* B = hb.data() + (size_t)j * 257;
            const i64 lsize = B[128] - B[0], hsize = B[256] - B[128];
            const bool rotated = lsize > 0 && hsize > 0;
            if (rotated) rots.push_back({B[0] + level[j].shift, lsize, hsize});
            for (int c = 0; c < 256; ++c) {
                if (B[c + 1] - B[c] <= ix.chuck_size) continue;
                const i64 sh = level[j].shift + (rotated ? (c < 128 ? hsize : -lsize) : 0);
                next.push_back({{B[c], B[c + 1], level[j].node.d + 1}, sh});
            }
        }
        level.swap(next);
    }
    for (const Rot& r : rots) {  // [L][H] -> [H][L]
        reverse_range<P>(sa + r.at, (u64)r.lsize, st);
        reverse_range<P>(sa + r.at + r.lsize, (u64)r.hsize, st);
        reverse_range<P>(sa + r.at, (u64)(r.lsize + r.hsize), st);
    }
    CDB_CUDA(cudaStreamSynchronize(st));
}

// 64-bit corpus hash (text, doc_off, ids): sum over all 8-byte words of mix(word, position) — order-independent, so
// it is one grid-stride pass with a warp-reduced atomic per warp.  Keys a saved suffix array to its corpus (persist.cu).
__host__ __device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void hash_words_kernel(const u64* __restrict__ w, u64 nwords, u64 salt, unsigned long long* __restrict__ acc) {
    u64 h = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (u64)gridDim.x * blockDim.x)
        h += mix64(ld_stream_u64(w + i) ^ mix64(i + salt));
#pragma unroll
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(acc, (unsigned long long)h);
}

__global__ void hash_tail_kernel(const u8* __restrict__ text, i64 from, i64 n, unsigned long long* __restrict__ acc) {
    u64 h = 0;
    for (i64 i = from + threadIdx.x; i < n; i += blockDim.x) h += mix64((u64)text[i] ^ mix64((u64)i + 0x7e57ull));
#pragma unroll
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0 && h) atomicAdd(acc, (unsigned long long)h);
}

u64 corpus_hash(const Index& ix, i64 n, cudaStream_t st) {
    DevBuf<unsigned long long> acc(1, st);
    CDB_CUDA(cudaMemsetAsync(acc.p, 0, 8, st));
    auto words = [&](const void* p, u64 nwords, u64 salt) {
        if (!nwords) return;
        const int g = (int)std::min<i64>(ceil_div((i64)nwords, 256), num_sms() * 8);
        hash_words_kernel<<<g, 256, 0, st>>>(reinterpret_cast<const u64*>(p), nwords, salt, acc.p);
        CDB_LAUNCH_CHECK();
    };
    words(ix.d_text, (u64)(n >> 3), 0x1000000000000000ull);  // text is 16-byte aligned
    if (n & 7) {
        hash_tail_kernel<<<1, 32, 0, st>>>(ix.d_text, n & ~(i64)7, n, acc.p);
        CDB_LAUNCH_CHECK();
    }
    words(ix.d_off, (u64)ix.nd + 1, 0x2000000000000000ull);
    if (ix.d_ids) words(ix.d_ids, (u64)ix.nd, 0x3000000000000000ull);
    unsigned long long h = 0;
    CDB_CUDA(cudaMemcpyAsync(&h, acc.p, 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    return (u64)h ^ mix64((u64)n * 1315423911ull + (u64)ix.nd);
}

void build_index(Index& ix, cudaStream_t st, const SavedArraySource* saved) {
    cudaEvent_t e0, e1;
    CDB_CUDA(cudaEventCreate(&e0));
    CDB_CUDA(cudaEventCreate(&e1));
    CDB_CUDA(cudaEventRecord(e0, st));
    ix.rounds = 0;
    ix.chunks = 0;
    ix.sort_ms = 0;
    // n and the width rule (src/index.cpp:183-208)
    i64 ends[2] = {0, 0};
    CDB_CUDA(cudaMemcpyAsync(&ends[0], ix.d_off, 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaMemcpyAsync(&ends[1], ix.d_off + ix.nd, 8, cudaMemcpyDeviceToHost, st));
    DevBuf<unsigned long long> d_max(1, st);
    DevBuf<u32> d_present(256, st);
    CDB_CUDA(cudaMemsetAsync(d_max.p, 0, 8, st));
    CDB_CUDA(cudaMemsetAsync(d_present.p, 0, 1024, st));
    if (ix.nd > 0) {
        int g = (int)std::min<i64>(ceil_div(ix.nd, 256), num_sms() * 8);
        doc_stats_kernel<<<g, 256, 0, st>>>(ix.d_off, ix.nd, d_max.p);
        CDB_LAUNCH_CHECK();
    }
    unsigned long long maxlen = 0;
    CDB_CUDA(cudaMemcpyAsync(&maxlen, d_max.p, 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    if (ends[0] != 0) throw Error(CDB_ERR_ARG, "doc_off[0] must be 0");
    ix.n = ends[1];
    ix.bits1 = ones_needed((u64)ix.nd);
    ix.bits2 = ones_needed((u64)maxlen);
    if (ix.bits1 + ix.bits2 > 64)
        throw Error(CDB_ERR_TOO_MUCH_DATA, "The amount of data exceeds the maximum range that CoffeeDB can handle");
    if (ix.bits1 > 32)
        throw Error(CDB_ERR_TOO_MANY_OBJECTS, "The number of objects exceeds the maximum range that CoffeeDB can handle");
    ix.mask = ix.bits1 >= 64 ? ~0ull : ((1ull << ix.bits1) - 1);
    ix.width = (ix.bits1 + ix.bits2 <= 32) ? 4 : 8;
    ix.chuck_size = std::max<i64>(4096, ix.n / 256);
    ix.mixed = false;
    if (ix.n > 0) {
        int g = (int)std::min<i64>(ceil_div(ix.n / 16 + 1, 256), num_sms() * 8);
        byte_presence_kernel<<<g, 256, 0, st>>>(ix.d_text, ix.n, d_present.p);
        CDB_LAUNCH_CHECK();
        u32 present[256];
        CDB_CUDA(cudaMemcpyAsync(present, d_present.p, 1024, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        SymTab tab;
        int sigma = 0;
        bool lo = false, hi = false;
        for (int c = 0; c < 256; ++c) {
            tab.sym[c] = 0;
            if (present[c]) {
                tab.sym[c] = (u16)(++sigma);
                (c < 0x80 ? lo : hi) = true;
            }
        }
        ix.mixed = lo && hi;
        ix.symtab = tab;
        ix.sigma = sigma;
        int b = bits_for((u64)sigma);  // symbols 0..sigma
        int S = 64 / b;
        if (S > EX_MAXS) S = EX_MAXS;
        ix.loaded_from_file = false;
        if (saved && saved->try_load(ix, st)) {
            // the saved array is the finished one (note-N1 layout included); only the prefix directory is rebuilt
            ix.loaded_from_file = true;
        } else {
            if (ix.width == 4)
                build_typed<u32>(ix, tab, b, S, st);
            else
                build_typed<u64>(ix, tab, b, S, st);
            if (ix.mixed && ix.opt.compat_signed && ix.n > ix.chuck_size) {
                if (ix.width == 4)
                    apply_signed_radix_layout<u32>(ix, st);
                else
                    apply_signed_radix_layout<u64>(ix, st);
            }
        }
        build_prefix_table(ix, st);
        // document listing in doc order (locate.cu): part of the index, so its cost is part of the build time
        get_listing(ix, 0, st);
    }
    CDB_CUDA(cudaEventRecord(e1, st));
    CDB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (getenv("CDB_DEBUG_TIMING")) fprintf(stderr, "[cdb] build: %.1f ms between the first and last event, sort %.1f ms\n", ms, ix.sort_ms);
    ix.build_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ix.built = true;
}

}  // namespace cdb
