// The document listing of the prefix directory's buckets (Listing, index.cuh) — no counterpart in the
// reference, whose query() gathers, sorts and run-length encodes its suffix-array interval on every call
// (src/index.cpp:288-322): for every directory bucket of at most 1024 suffixes the ids of the suffixes' documents are
// stored at the suffixes' ranks, already in the order query() (doc index) or filter() (id) reports them, so that a keyword
// of exactly pt_k symbols is answered by streaming its bucket.
//   build   listing_build_kernel (+ listing_translate_kernel when ids[] outgrows L2), once per index and order
//   query   listing_rowlen_kernel (row lengths come from the directory's tags), listing_emit_kernel (the rows)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>

#include "index.cuh"
#include "locate.cuh"
#include "primitives.cuh"
#include "radix_sort.cuh"
#include "warp_sort.cuh"

namespace cdb {

// Built once per index and order from the finished suffix array: one warp per directory bucket of <= kWarpCap suffixes
// reads the bucket, maps the elements to their sort keys (doc index; id rank for the id order), sorts them with the
// gather's own sorts and stores (id - base) of every suffix's document at the suffix's rank, in key order.  A keyword of
// exactly pt_k symbols is one bucket: listing_emit_kernel streams its row — 4 + hw bytes read per occurrence, 16 bytes
// written per (id, count) pair, no sort and no random lookup at query time.
__global__ void ids_minmax_kernel(const i64* __restrict__ ids, i64 nd, long long* __restrict__ mm) {
    long long lo = 0x7fffffffffffffffll, hi = -0x7fffffffffffffffll - 1;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += (i64)gridDim.x * blockDim.x) {
        const long long v = ids[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 0, lo);
        atomicMax(mm + 1, hi);
    }
}

template <typename SAT, int R>
__device__ __forceinline__ u32 listing_bucket(const SAT* __restrict__ sa, i64 l, int occ, u64 mask, u32 bucket_mul, u32* s_a, u32* s_b,
                                              int lane, const u32* __restrict__ remap, const i64* __restrict__ table, i64 base,
                                              int hw, u32* __restrict__ out_lo, void* __restrict__ out_hi, int* __restrict__ dup_flag,
                                              u16* __restrict__ seg_row, int nranges, int rshift) {
    u32 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        x[r] = i < occ ? (u32)((u64)ld_stream(sa + l + i) & mask) : 0xffffffffu;
    }
    if (remap) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r * 32 + lane < occ) x[r] = ld_stream_u32(remap + x[r]);
    }
    bool sorted = false;
    if constexpr (R >= kBucketMinR) {
        if (bucket_mul) sorted = warp_bucket_sort<R>(x, occ, bucket_mul, s_a, s_b, lane);
    }
    if (!sorted) warp_bitonic_regs<R>(x, lane, s_a);
    // blocked layout: lane holds ranks lane*R .. lane*R + R-1; a rank is a run head when its key differs from the previous one's
    const u32 prev_x = __shfl_up_sync(0xffffffffu, x[R - 1], 1);
    int nh = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int idx = lane * R + r;
        const u32 px = r ? x[r > 0 ? r - 1 : 0] : prev_x;
        nh += (idx < occ && (idx == 0 || x[r] != px)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) nh += __shfl_xor_sync(0xffffffffu, nh, o);
    // through shared memory into the striped layout (rank r*32 + lane): coalesced stores, and the keys leave the registers
    // before the values arrive
    __syncwarp();
    {
        u32* da = s_a + lane * R + ((lane * R) >> 5);
#pragma unroll
        for (int r = 0; r < R; ++r) da[r] = x[r];
    }
    __syncwarp();
    if (seg_row) {
        // two-phase build: the sorted KEYS go out now (coalesced), with the row's split points at the doc-range boundaries;
        // listing_translate_kernel turns them into ids range by range, so that the slice of ids[] in use stays in L2
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = r * 32 + lane;
            if (i < occ) out_lo[l + i] = s_a[pad_idx(i)];
        }
        for (int r = lane; r <= nranges; r += 32) {
            const u64 bound = (u64)r << rshift;
            int lo = 0, hi = occ;  // first rank whose key is >= bound
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((u64)s_a[pad_idx(mid)] < bound)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            seg_row[r * kTileWarps] = (u16)lo;
        }
        __syncwarp();
        return (u32)nh;
    }
    u64 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        x[r] = s_a[pad_idx(i)];
        // (no L1 allocation: a caching load of these random 8-byte reads moved ~110 bytes of DRAM per lookup)
        v[r] = i < occ ? (u64)((i64)ld_stream_u64(reinterpret_cast<const u64*>(table) + x[r]) - base) : 0ull;
    }
    bool dup = false;
    u32 carry_x = 0;
    u64 carry_v = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        u32 px = __shfl_up_sync(0xffffffffu, x[r], 1);
        u64 pv = __shfl_up_sync(0xffffffffu, v[r], 1);
        if (lane == 0) {
            px = carry_x;
            pv = carry_v;
        }
        dup |= i > 0 && i < occ && x[r] != px && v[r] == pv;  // two documents, one id: a listed row could not tell them apart
        carry_x = __shfl_sync(0xffffffffu, x[r], 31);
        carry_v = __shfl_sync(0xffffffffu, v[r], 31);
        if (i < occ) {
            out_lo[l + i] = (u32)v[r];
            const u32 h = (u32)(v[r] >> 32);
            if (hw == 1) reinterpret_cast<u8*>(out_hi)[l + i] = (u8)h;
            else if (hw == 2) reinterpret_cast<u16*>(out_hi)[l + i] = (u16)h;
            else if (hw == 4) reinterpret_cast<u32*>(out_hi)[l + i] = h;
        }
    }
    if (dup) *dup_flag = 1;
    __syncwarp();
    return (u32)nh;
}

template <typename SAT>
__global__ void __launch_bounds__(kTileWarps * 32, 2) listing_build_kernel(const SAT* __restrict__ sa, u64 mask, u32 bucket_mul,
                                                                          u64* ptab, u64 nentries,
                                                                          const u32* __restrict__ remap, const i64* __restrict__ table,
                                                                          i64 base, int hw, u32* __restrict__ out_lo,
                                                                          void* __restrict__ out_hi, int* __restrict__ dup_flag,
                                                                          u16* __restrict__ seg, int nranges, int rshift) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 c = (u64)blockIdx.x * kTileWarps + warp;
    if (c >= nentries) return;
    // seg layout as gather_kernel's: [c / 8][r = 0..nranges][c % 8]
    u16* seg_row = seg ? seg + (size_t)(c / kTileWarps) * (nranges + 1) * kTileWarps + (c % kTileWarps) : nullptr;
    // other warps tag their own entries while this one reads: 8-byte accesses, the rank bits never change
    const u64 e_lo = *reinterpret_cast<const volatile u64*>(ptab + c);
    const i64 l = (i64)(e_lo & kPtRank);
    const i64 occ64 = (i64)(*reinterpret_cast<const volatile u64*>(ptab + c + 1) & kPtRank) - l;
    if (occ64 <= 0 || occ64 > kWarpCap) {
        if (lane == 0 && (e_lo >> kPtCountShift)) ptab[c] = e_lo & kPtRank;
        if (seg_row)
            for (int r = lane; r <= nranges; r += 32) seg_row[r * kTileWarps] = 0;
        return;
    }
    u32* s_a = reinterpret_cast<u32*>(smem_raw + (size_t)warp * warp_smem_bytes<32>());
    u32* s_b = s_a + 32 * 32 + 32;
    const int occ = (int)occ64;
    u32 d;
#define CDB_LB(RR) listing_bucket<SAT, RR>(sa, l, occ, mask, bucket_mul, s_a, s_b, lane, remap, table, base, hw, out_lo, out_hi, dup_flag, seg_row, nranges, rshift)
    if (occ <= 32) d = CDB_LB(1);
    else if (occ <= 64) d = CDB_LB(2);
    else if (occ <= 128) d = CDB_LB(4);
    else if (occ <= 256) d = CDB_LB(8);
    else if (occ <= 512) d = CDB_LB(16);
    else d = CDB_LB(32);
#undef CDB_LB
    if (lane == 0) ptab[c] = (u64)l | (1ull << 63) | (d != (u32)occ ? (1ull << 62) : 0ull) | ((u64)d << kPtCountShift);
}

// Phase 2 of the two-phase listing build: lo[] holds the sorted keys (doc indices / id ranks) of every listed bucket; they
// become (id - base) in place.  Same order of work as translate_kernel — item = (doc range, block of 32 buckets), handed
// out through one ticket so that the whole grid looks up one <= 32 MB slice of the table at a time (L2 evict_last) —
// because looked up bucket by bucket every random 8-byte read cost a ~110-byte DRAM fetch (1.18 TB for 10^10 suffixes).
constexpr int kLtU = 4;
__global__ void __launch_bounds__(kTrWarps * 32) listing_translate_kernel(u32* __restrict__ lo, void* __restrict__ hi, int hw,
                                                                          const u64* __restrict__ ptab, const u16* __restrict__ seg,
                                                                          const i64* __restrict__ table, i64 base, u64 nentries,
                                                                          int nranges, unsigned long long* ticket) {
    __shared__ u32 s_excl[kTrWarps][32];
    __shared__ u64 s_pos[kTrWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 ntile = (i64)((nentries + 31) >> 5);
    const i64 nitems = ntile * nranges;
    const u64 pol_keep = l2_policy_evict_last();
    const u64 pol_stream = l2_policy_evict_first();
    const u64* tab = reinterpret_cast<const u64*>(table);
    for (;;) {
        i64 item = 0;
        if (lane == 0) item = (i64)atomicAdd(ticket, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const int r = (int)(item / ntile);
        const u64 c = (u64)(item - (i64)r * ntile) * 32 + lane;
        u32 len = 0;
        u64 pos0 = 0;
        if (c < nentries) {
            const u16* sg = seg + ((size_t)(c / kTileWarps) * (nranges + 1) + r) * kTileWarps + (c % kTileWarps);
            const u32 s = sg[0], e = sg[kTileWarps];
            len = e - s;
            if (len) pos0 = (__ldg(ptab + c) & kPtRank) + s;
        }
        u32 incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const u32 tot = __shfl_sync(0xffffffffu, incl, 31);
        if (tot == 0) continue;
        __syncwarp();
        s_excl[warp][lane] = incl - len;
        s_pos[warp][lane] = pos0 - (incl - len);
        __syncwarp();
        for (u32 i0 = 0; i0 < tot; i0 += 32 * kLtU) {
            u64 p[kLtU];
            u32 key[kLtU];
#pragma unroll
            for (int u = 0; u < kLtU; ++u) {
                const u32 idx = i0 + u * 32 + lane;
                if (idx < tot) {
                    int j = 0;  // largest j with excl[j] <= idx
#pragma unroll
                    for (int st = 16; st; st >>= 1)
                        if (s_excl[warp][j + st] <= idx) j += st;
                    p[u] = s_pos[warp][j] + idx;
                    key[u] = ld_hint_u32(lo + p[u], pol_stream);
                }
            }
            __syncwarp();
            u64 v[kLtU];
#pragma unroll
            for (int u = 0; u < kLtU; ++u) {
                const u32 idx = i0 + u * 32 + lane;
                if (idx < tot) v[u] = (u64)((i64)ld_hint_u64(tab + key[u], pol_keep) - base);
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kLtU; ++u) {
                const u32 idx = i0 + u * 32 + lane;
                if (idx < tot) {
                    lo[p[u]] = (u32)v[u];
                    const u32 h = (u32)(v[u] >> 32);
                    if (hw == 1) reinterpret_cast<u8*>(hi)[p[u]] = (u8)h;
                    else if (hw == 2) reinterpret_cast<u16*>(hi)[p[u]] = (u16)h;
                    else if (hw == 4) reinterpret_cast<u32*>(hi)[p[u]] = h;
                }
            }
        }
    }
}

__global__ void adjacent_equal_kernel(const u64* __restrict__ keys, i64 n, int* __restrict__ flag) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < n && keys[i] == keys[i + 1]) *flag = 1;
}
__global__ void ids_keys_kernel(const i64* __restrict__ ids, i64 nd, u64* __restrict__ keys) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nd) keys[i] = (u64)ids[i];
}

// do two documents share an id?  (a listed row could not tell them apart: the listing is refused then)
static bool ids_have_duplicates(const Index& ix, cudaStream_t st) {
    const i64 nd = ix.nd;
    if (nd < 2) return false;
    BigBuf<u64> k0((size_t)nd), k1((size_t)nd);
    const unsigned grid = (unsigned)ceil_div(nd, 256);
    ids_keys_kernel<<<grid, 256, 0, st>>>(ix.d_ids, nd, k0.p);
    CDB_LAUNCH_CHECK();
    const int cur = rs::radix_sort_pairs<rs::NoValue>(k0.p, k1.p, nullptr, nullptr, (u64)nd, 0, 64, st);
    DevBuf<int> flag(1, st);
    CDB_CUDA(cudaMemsetAsync(flag.p, 0, 4, st));
    adjacent_equal_kernel<<<grid, 256, 0, st>>>(cur ? k1.p : k0.p, nd, flag.p);
    CDB_LAUNCH_CHECK();
    int h = 0;
    CDB_CUDA(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    return h != 0;
}

template <int HW>
__device__ __forceinline__ u64 listing_value(const u32* __restrict__ lo, const void* __restrict__ hi, i64 i) {
    u64 v = (u64)ld_stream_u32(lo + i);
    if (HW == 1) v |= (u64)__ldg(reinterpret_cast<const u8*>(hi) + i) << 32;
    if (HW == 2) v |= (u64)__ldg(reinterpret_cast<const u16*>(hi) + i) << 32;
    if (HW == 4) v |= (u64)ld_stream_u32(reinterpret_cast<const u32*>(hi) + i) << 32;
    return v;
}

__global__ void listing_rowlen_kernel(const u64* __restrict__ pre, i64 npat, u64* __restrict__ rowlen, int write_zero) {
    const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npat) return;
    const u64 p = pre[q];
    if (p & kPreListed)
        rowlen[q] = p & kPreCount;
    else if (write_zero)
        rowlen[q] = 0;
}

// The listed rows of a batch: pairs[row_off[q] + j] = (base + listing[left[q] + j], 1) — a streaming copy — or, for the
// rows in which a document repeats, the run-length encoding of the listed values (equal values are adjacent, ids are
// distinct).  Persistent grid, one warp per row, the next row's descriptors in flight while the current one streams.
#ifndef CDB_EMIT_U
#define CDB_EMIT_U 4
#endif
#ifndef CDB_EMIT_MINB
#define CDB_EMIT_MINB 1
#endif
#ifndef CDB_EMIT_ST
#define CDB_EMIT_ST 0
#endif
constexpr int kEmU = CDB_EMIT_U;
__device__ __forceinline__ void emit_store(i64* p, longlong2 v, u64 pol) {
#if CDB_EMIT_ST == 0
    st_hint_v2(p, v, pol);
#elif CDB_EMIT_ST == 1
    *reinterpret_cast<longlong2*>(p) = v;
#else
    asm volatile("st.global.cs.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
#endif
}
template <int HW>
__global__ void __launch_bounds__(kTileWarps * 32, CDB_EMIT_MINB) listing_emit_kernel(const u32* __restrict__ lo, const void* __restrict__ hi, i64 base,
                                                                     const u64* __restrict__ pre, const i64* __restrict__ left,
                                                                     const i64* __restrict__ right, const u64* __restrict__ row_off,
                                                                     i64 npat, i64* __restrict__ pairs, int mode,
                                                                     const u8* __restrict__ need) {
    // mode 0: every listed row; 1: only the rows in which a document repeats (the others are read lazily by the caller);
    // 2: the rows without repeats that the caller needs in full after all (need[q] != 0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 stride = (i64)gridDim.x * kTileWarps;
    const u64 pol = l2_policy_evict_first();
    i64 q = (i64)blockIdx.x * kTileWarps + warp;
    u64 p = 0, out = 0;
    i64 l = 0, rg = 0;
    if (q < npat) {
        p = pre[q];
        l = left[q];
        rg = right[q];
        out = row_off[q];
    }
    while (q < npat) {
        const i64 qn = q + stride;
        u64 pn = 0, outn = 0;
        i64 ln = 0, rn = 0;
        if (qn < npat) {
            pn = pre[qn];
            ln = left[qn];
            rn = right[qn];
            outn = row_off[qn];
        }
        bool go = (p & kPreListed) != 0;
        if (mode == 1) go = go && (p & kPreRepeat);
        if (mode == 2) go = go && !(p & kPreRepeat) && need[q];
        if (go) {
            const int occ = (int)(rg - l);
            if (!(p & kPreRepeat)) {
                for (int j0 = 0; j0 < occ; j0 += 32 * kEmU) {
                    u64 v[kEmU];
#pragma unroll
                    for (int u = 0; u < kEmU; ++u) {
                        const int j = j0 + u * 32 + lane;
                        if (j < occ) v[u] = listing_value<HW>(lo, hi, l + j);
                    }
#pragma unroll
                    for (int u = 0; u < kEmU; ++u) {
                        const int j = j0 + u * 32 + lane;
                        if (j < occ) emit_store(pairs + 2 * (out + (u64)j), make_longlong2(base + (i64)v[u], 1), pol);
                    }
                }
            } else {
                int carry_head = 0;     // rank of the last run head before this chunk
                u32 heads_before = 0;   // run heads before this chunk
                u64 last_v = 0;
                for (int j0 = 0; j0 < occ; j0 += 32) {
                    const int j = j0 + lane;
                    const bool valid = j < occ;
                    const u64 v = valid ? listing_value<HW>(lo, hi, l + j) : 0ull;
                    u64 prev = __shfl_up_sync(0xffffffffu, v, 1);
                    if (lane == 0) prev = last_v;
                    u64 next = __shfl_down_sync(0xffffffffu, v, 1);
                    const bool nvalid = j + 1 < occ;
                    if (lane == 31 && nvalid) next = listing_value<HW>(lo, hi, l + j + 1);
                    const bool head = valid && (j == 0 || v != prev);
                    const bool tail = valid && (!nvalid || next != v);
                    const u32 hm = __ballot_sync(0xffffffffu, head);
                    const u32 le = hm & (0xffffffffu >> (31 - lane));  // run heads at lanes <= lane
                    const int headpos = le ? j0 + (31 - __clz(le)) : carry_head;
                    const u32 idx = heads_before + __popc(le) - 1;
                    if (tail) st_hint_v2(pairs + 2 * (out + (u64)idx), make_longlong2(base + (i64)v, (i64)(j - headpos + 1)), pol);
                    if (hm) carry_head = j0 + (31 - __clz(hm));
                    heads_before += __popc(hm);
                    last_v = __shfl_sync(0xffffffffu, v, 31);
                }
            }
        }
        q = qn;
        p = pn;
        l = ln;
        rg = rn;
        out = outn;
    }
}

static size_t device_memory_available(int device, size_t* total_out) {
    // memory held by the stream-ordered pool for query temporaries counts as available (it is re-used, not lost)
    size_t free_b = 0, total_b = 0;
    CDB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    cudaMemPool_t pool;
    unsigned long long reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
    }
    if (total_out) *total_out = total_b;
    return free_b + (size_t)(reserved > used ? reserved - used : 0);
}

static bool big_malloc(void** p, size_t bytes, int device) {
    if (cudaMalloc(p, bytes) == cudaSuccess) return true;
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    if (cudaMalloc(p, bytes) == cudaSuccess) return true;
    cudaGetLastError();
    *p = nullptr;
    return false;
}

struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    EventPair() {
        CDB_CUDA(cudaEventCreate(&a));
        CDB_CUDA(cudaEventCreate(&b));
    }
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};

// *deferred: the listing was not built because the other order's would have to go and this order has not been asked for
// often enough yet (see below), or because device memory is short right now (e.g. calls in flight still hold the listing that
// was dropped) — the caller takes the suffix-array path and the build is tried again later.  Any other empty result is final
// for this index (ids too wide, two documents with one id, CDB_LISTING_MAX_HW).
template <typename SAT>
static std::shared_ptr<Listing> build_listing_typed(const Index& ix, int order, cudaStream_t st, bool* deferred) {
    // order 1 is only asked for when the ids do not ascend with the doc index: rank_tab / ids_by_rank exist then
    const u32* remap = order ? ix.d_rank_tab : nullptr;
    const i64* table = order ? ix.d_ids_by_rank : ix.d_ids;
    const u64 nentries = 1ull << (ix.pt_b * ix.pt_k);
    EventPair evp;
    const cudaEvent_t e0 = evp.a, e1 = evp.b;
    CDB_CUDA(cudaEventRecord(e0, st));
    // width of (id - smallest id)
    long long h_mm[2] = {0x7fffffffffffffffll, -0x7fffffffffffffffll - 1};
    {
        DevBuf<long long> mm(2, st);
        CDB_CUDA(cudaMemcpyAsync(mm.p, h_mm, 16, cudaMemcpyHostToDevice, st));
        ids_minmax_kernel<<<num_sms() * 4, 256, 0, st>>>(ix.d_ids, ix.nd, mm.p);
        CDB_LAUNCH_CHECK();
        CDB_CUDA(cudaMemcpyAsync(h_mm, mm.p, 16, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
    }
    const u64 span = (u64)h_mm[1] - (u64)h_mm[0];  // exact in unsigned arithmetic
    if (h_mm[1] < h_mm[0] || span >> 63) return {};
    const int hw = span < (1ull << 32) ? 0 : span < (1ull << 40) ? 1 : span < (1ull << 48) ? 2 : 4;
    if (const char* e = getenv("CDB_LISTING_MAX_HW"))
        if (hw > atoi(e)) return {};
    const size_t need = (size_t)ix.n * (4 + hw) + 256;
    size_t total_b = 0;
    size_t avail = device_memory_available(ix.device, &total_b);
    const size_t keep_free = total_b / 8;  // query temporaries and results
    size_t budget = ~(size_t)0;            // CDB_LISTING_BUDGET_MB: cap on the listings of one index (tests: forces the swap)
    if (const char* e = getenv("CDB_LISTING_BUDGET_MB")) budget = (size_t)atoll(e) << 20;
    const int other = 1 - order;
    auto other_bytes = [&]() { return ix.listing[other] && ix.listing[other] != ix.listing[order] ? ix.listing[other]->bytes : (size_t)0; };
    if (avail < need + keep_free || need + other_bytes() > budget) {
        // make room: the listing of the other order goes (it is rebuilt when that order is asked for again; calls that are
        // using it keep it alive until they return) — but only once this order has been asked for CDB_LISTING_SWAP_AFTER
        // times (default 2) with no call of the other order in between: callers that alternate between query() and
        // filter() keep one listing and pay the suffix-array path for the other order instead of a rebuild per call
        if (other_bytes()) {
            int after = 2;
            if (const char* e = getenv("CDB_LISTING_SWAP_AFTER")) after = atoi(e);
            if (++ix.listing_miss[order] < after) {
                *deferred = true;
                return {};
            }
            ix.listing[other].reset();
            ix.listing_state[other] = 0;
            avail = device_memory_available(ix.device, &total_b);
        }
        if (need > budget) return {};
        if (avail < need + keep_free) {
            *deferred = true;
            ix.listing_skip[order] = 64;  // not before 64 more calls of this order: the check itself costs a scan of the ids
            return {};
        }
    }
    auto L = std::make_shared<Listing>();
    L->hw = hw;
    L->base = (i64)h_mm[0];
    L->bytes = need;
    if (!big_malloc((void**)&L->lo, (size_t)ix.n * 4, ix.device) || (hw && !big_malloc(&L->hi, (size_t)ix.n * hw, ix.device)) ||
        !big_malloc((void**)&L->d_flag, 4, ix.device)) {
        *deferred = true;
        ix.listing_skip[order] = 64;
        return {};
    }
    CDB_CUDA(cudaMemsetAsync(L->d_flag, 0, 4, st));
    const char* env_buckets = getenv("CDB_GATHER_BUCKETS");
    const bool use_buckets = !env_buckets || atoi(env_buckets) != 0;
    const u32 bucket_mul = use_buckets && ix.nd > 0 ? (u32)std::min<u64>(0xffffffffull, (1024ull << 32) / (u64)ix.nd) : 0u;
    const size_t smem = (size_t)kTileWarps * warp_smem_bytes<32>();
    CDB_CUDA(cudaFuncSetAttribute(listing_build_kernel<SAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Two phases when the table of ids is larger than L2 can hold (CDB_LISTING_TWO_PHASE forces 1 / 0): sorted keys + split
    // points first, ids range by range afterwards.  One phase otherwise: the lookups hit L2 anyway.
    int nranges, rshift;
    ids_ranges(ix.nd, &nranges, &rshift);
    bool two = (size_t)ix.nd * 8 > ((size_t)48 << 20);
    if (const char* e = getenv("CDB_LISTING_TWO_PHASE")) two = atoi(e) != 0;
    BigBuf<u16> seg;
    if (two) {
        if (ids_have_duplicates(ix, st)) return {};
        const size_t nseg = (size_t)ceil_div((i64)nentries, kTileWarps) * kTileWarps * (nranges + 1);
        u16* p = nullptr;
        if (big_malloc((void**)&p, nseg * 2, ix.device)) {
            seg.p = p;
            seg.n = nseg;
        } else {
            two = false;
        }
    }
    listing_build_kernel<SAT><<<(unsigned)ceil_div((i64)nentries, kTileWarps), kTileWarps * 32, smem, st>>>(
        reinterpret_cast<const SAT*>(ix.d_sa), ix.mask, bucket_mul, ix.d_ptab, nentries, remap, table, L->base, hw, L->lo, L->hi,
        L->d_flag, seg.p, nranges, rshift);
    CDB_LAUNCH_CHECK();
    EventPair evm;
    CDB_CUDA(cudaEventRecord(evm.a, st));
    if (two) {
        DevBuf<unsigned long long> ticket(1, st);
        CDB_CUDA(cudaMemsetAsync(ticket.p, 0, 8, st));
        const i64 nitems = ceil_div((i64)nentries, 32) * nranges;
        const int per_sm = resident_ctas((const void*)listing_translate_kernel, kTrWarps * 32);
        const int grid = (int)std::min<i64>(ceil_div(nitems, kTrWarps), (i64)num_sms() * per_sm);
        listing_translate_kernel<<<grid, kTrWarps * 32, 0, st>>>(L->lo, L->hi, hw, ix.d_ptab, seg.p, table, L->base, nentries, nranges,
                                                                  ticket.p);
        CDB_LAUNCH_CHECK();
        CDB_CUDA(cudaStreamSynchronize(st));  // seg and the ticket are released below
    }
    int h_flag = 0;
    CDB_CUDA(cudaMemcpyAsync(&h_flag, L->d_flag, 4, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaEventRecord(e1, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    L->build_ms = ms;
    if (getenv("CDB_DEBUG_TIMING")) {
        float ms1 = 0;
        cudaEventElapsedTime(&ms1, e0, evm.a);
        fprintf(stderr, "[cdb] document listing (order %d, %d + 4 bytes per suffix, %s): %.1f ms (%.1f ms up to the end of the sort phase)%s\n", order, hw,
                two ? "two phases" : "one phase", ms, ms1, h_flag ? " — dropped: two documents share an id" : "");
    }
    if (h_flag) return {};
    return L;
}

std::shared_ptr<Listing> get_listing(const Index& ix, int order, cudaStream_t st) {
    if (!ix.d_ptab || ix.pt_k <= 0 || ix.n <= 0 || ix.nd <= 0) return {};
    std::lock_guard<std::mutex> lk(ix.listing_mu);
    if (ix.listing_state[order] > 0) {
        ix.listing_miss[1 - order] = 0;
        return ix.listing[order];
    }
    if (ix.listing_state[order] < 0) return {};
    if (ix.listing_skip[order] > 0) {  // memory was short a moment ago
        --ix.listing_skip[order];
        return {};
    }
    const char* e = getenv("CDB_LISTING");
    std::shared_ptr<Listing> L;
    bool deferred = false;
    if (!e || atoi(e) != 0)
        L = ix.width == 4 ? build_listing_typed<u32>(ix, order, st, &deferred) : build_listing_typed<u64>(ix, order, st, &deferred);
    if (deferred) return {};
    ix.listing[order] = L;
    ix.listing_state[order] = L ? 1 : -1;
    ix.listing_miss[order] = 0;
    return L;
}

void launch_listing_emit(const Listing& L, const u64* pre, const i64* left, const i64* right, const u64* row_off, i64 npat,
                                i64* pairs, int mode, const u8* need, cudaStream_t st, int spare_ctas) {
    auto emit = [&](auto kernel) {
        int per_sm = resident_ctas((const void*)kernel, kTileWarps * 32);
        if (const char* e = getenv("CDB_EMIT_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // experiment knob
        // spare_ctas: room left for a collective that runs beside this persistent grid (a sharded caller's all_gather)
        const i64 wave = std::max<i64>((i64)num_sms(), (i64)num_sms() * per_sm - spare_ctas);
        const i64 grid = std::min<i64>(ceil_div(npat, kTileWarps), wave);
        kernel<<<(unsigned)grid, kTileWarps * 32, 0, st>>>(L.lo, L.hi, L.base, pre, left, right, row_off, npat, pairs, mode, need);
    };
    switch (L.hw) {
        case 0: emit(listing_emit_kernel<0>); break;
        case 1: emit(listing_emit_kernel<1>); break;
        case 2: emit(listing_emit_kernel<2>); break;
        default: emit(listing_emit_kernel<4>); break;
    }
    CDB_LAUNCH_CHECK();
}

void emit_listed_rows(const cdb_device_result& res, const LazyListed& lazy, const u8* d_need, cudaStream_t st) {
    if (!lazy.active || !lazy.lst || res.npat <= 0) return;
    launch_listing_emit(*lazy.lst, lazy.pre, res.left, res.right, reinterpret_cast<const u64*>(res.row_off), res.npat, res.pairs, 2,
                        d_need, st, 0);
}

void launch_listing_rowlen(const u64* pre, i64 npat, u64* rowlen, int write_zero, cudaStream_t st) {
    listing_rowlen_kernel<<<(unsigned)ceil_div(npat, 256), 256, 0, st>>>(pre, npat, rowlen, write_zero);
    CDB_LAUNCH_CHECK();
}

}  // namespace cdb
