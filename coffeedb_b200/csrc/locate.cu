// Batched substring locate — replaces string_index::query() (src/index.cpp:237-326; kernels K4-K7 of
// SURVEY.md §2a) for a whole batch of patterns at once.
//
//   K4  search_kernel      one thread per pattern.  Sorted array: the prefix directory (first k symbols -> SA
//                          interval) answers keywords of <= k symbols with two lookups and narrows longer ones to one
//                          bucket, refined by lower/upper bounds.  Note-N1 layout: the reference's two binary-search
//                          recurrences verbatim (same midpoints, same predicates), so the interval is identical on
//                          that not-quite-sorted array.  A probe = SA element -> doc_off pair -> 16 bytes of text,
//                          compared as big-endian integers (== unsigned memcmp).
//   K5-K7 gather_kernel    phase A, one launch per batch, one warp per pattern (occ <= kWarpCap): coalesced read of the SA
//                          interval, doc = element & mask, sort of the <= 1024 doc indices in registers and the warp's
//                          shared memory — a DISTRIBUTION SORT when the interval's documents are spread over the corpus
//                          (the usual case: counting + scatter with shared-memory atomics, then a few odd-even
//                          transposition phases), an all-ascending bitonic network when a bucket overflows (clustered
//                          or repeated documents) and for intervals <= 128 — then run-length encoding from the
//                          registers.  Warps never wait for each other: the compact row (u32 docs, + u16 counts when a
//                          document is hit more than once) goes to the scan of the occurrence counts (an upper bound),
//                          the exact CSR offsets come from a scan of the row lengths afterwards.  Variants for batches
//                          of short intervals (<= 128 / 256 / 512) use fewer registers; the two shortest run as a
//                          persistent grid.  In id order (for cdb_filter) the keys are id ranks instead of doc indices.
//         translate_kernel phase B: pairs = (ids[doc], count), ordered by doc range so that the slice of ids[] in use
//                          stays in L2 (a fused gather spent 99 GB of DRAM reads on 25.8 GB of algorithmic bytes).
//                          Bound by the per-SM L1 -> crossbar request port, not by HBM.
//   listed rows            keywords of exactly pt_k symbols whose directory bucket is listed (listing.cu): the search reads the
//                          row length from the directory entry's tag, the row is streamed from the document listing
//                          (listing_emit_kernel) — phases A and B are skipped for these rows, for the whole batch when
//                          every row is listed.
//   small batches          <= 256 keywords: one upload, search_kernel + ONE fused kernel (small_gather_kernel) writing
//                          into mapped pinned memory, one synchronisation (a query() per request: 34 us instead of 150).
//   K5-K7 large path       patterns with longer intervals are expanded into (entry << 32 | doc) keys, sorted by
//                          the device radix sort (the same engine as the build) and run-length encoded, in
//                          sub-batches of bounded size; their row lengths are known before gather_kernel runs and
//                          enter the same scan, the rows themselves are emitted straight into the CSR result.
//   K8  highlight spans    live in spans.cu (batched over (request, document) texts).
// All integer work (64*S + w*occ + 24*d algorithmic bytes per pattern).
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "index.cuh"
#include "locate.cuh"
#include "primitives.cuh"
#include "radix_sort.cuh"
#include "warp_sort.cuh"

namespace cdb {

// ---- K4 ------------------------------------------------------------------------------------------------------
struct SearchCtx {
    const void* sa;
    i64 n;
    int bits1;
    u64 mask;
    const i64* doc_off;
    const u8* text;
    const u64* ptab;  // prefix directory (Index::d_ptab) or nullptr
    int pt_b, pt_k;
    bool listed;  // a document listing of the directory's buckets is in use: the high bits of ptab[] say which rows it answers
};

static SearchCtx make_ctx(const Index& ix) {
    return SearchCtx{ix.d_sa, ix.n, ix.bits1, ix.mask, ix.d_off, ix.d_text, ix.d_ptab, ix.pt_b, ix.pt_k, false};
}

// (kPreListed / kPreRepeat / kPreCount — the per-pattern word of a row answered from the document listing — are in locate.cuh)
// three-way comparison of keyword vs the suffix stored at SA rank M:
//   -1: keyword <  suffix            0: keyword is a prefix of suffix (keyword <= suffix, starts_with)
//   +1: keyword >  suffix (including "suffix is a proper prefix of keyword")
// == std::string_view::compare / substr(0,m)== of src/index.cpp:268,280 (unsigned bytes, shorter first)
template <typename SAT>
__device__ __forceinline__ int compare_at(const SearchCtx& c, i64 M, const u8* __restrict__ kw, i64 m, u64 p8) {
    const u64 e = (u64) reinterpret_cast<const SAT*>(c.sa)[M];
    const i64 doc = (i64)(e & c.mask);
    const i64 off = (i64)(e >> c.bits1);
    const i64 ds = __ldg(c.doc_off + doc), de = __ldg(c.doc_off + doc + 1);
    const i64 s = ds + off;
    const i64 slen = de - s;
    u64 pp = p8;
    i64 o = 0;
    for (;;) {
        const i64 rm = m - o, rs_ = slen - o;
        if (rm <= 0) return 0;
        if (rs_ <= 0) return 1;
        i64 kk = rm < rs_ ? rm : rs_;
        if (kk > 8) kk = 8;
        const u64 tt = load_be64(c.text, s + o);
        const int sh = 64 - 8 * (int)kk;
        const u64 a = pp >> sh, b = tt >> sh;
        if (a != b) return a < b ? -1 : 1;
        o += kk;
        if (o < m && kk == 8) {  // next 8 keyword bytes (rare: keywords longer than 8 that match so far)
            pp = 0;
            const i64 r2 = m - o < 8 ? m - o : 8;
            for (i64 i = 0; i < r2; ++i) pp |= (u64)kw[o + i] << (56 - 8 * i);
        }
    }
}

template <typename SAT>
__device__ __forceinline__ i64 search_one(const SearchCtx& c, const u8* __restrict__ pat, const i64* __restrict__ pat_off,
                                          i64 q, i64* __restrict__ left_out, i64* __restrict__ right_out,
                                          int* __restrict__ err, const u16* s_tab, u64& pre) {
    pre = 0;
    const i64 ps = pat_off[q];
    const i64 m = pat_off[q + 1] - ps;
    if (m <= 0) {  // src/index.cpp:239-241
        *err = 1;
        left_out[q] = 0;
        right_out[q] = 0;
        return 0;
    }
    const u8* kw = pat + ps;
    u64 p8 = 0;
    {
        const int m8 = m < 8 ? (int)m : 8;
        for (int i = 0; i < m8; ++i) p8 |= (u64)kw[i] << (56 - 8 * i);
    }
    if (c.n == 0) {
        left_out[q] = 0;
        right_out[q] = 0;
        return 0;
    }
    if (c.ptab) {
        // Sorted array (one comparator): the directory gives the interval of the first min(m, k) symbols with two
        // lookups; longer keywords refine inside it with plain lower/upper bounds.  Same interval as the
        // reference's recurrences whenever the array is sorted, which is the only case the directory exists in.
        const int k = c.pt_k, b = c.pt_b;
        const int kk = m < (i64)k ? (int)m : k;
        u64 code = 0;
        bool absent = false;
        for (int i = 0; i < kk; ++i) {
            const u32 sym = s_tab[kw[i]];
            absent |= sym == 0;
            code = (code << b) | sym;
        }
        i64 lo = 0, hi = 0;
        if (!absent) {
            const int sh = b * (k - kk);
            const u64 e_lo = __ldg(c.ptab + (code << sh));
            lo = (i64)(e_lo & kPtRank);
            hi = (i64)(__ldg(c.ptab + ((code + 1) << sh)) & kPtRank);
            if (m > (i64)k && lo < hi) {
                i64 L = lo, R = hi;
                while (L < R) {
                    const i64 M = L + (R - L) / 2;
                    if (compare_at<SAT>(c, M, kw, m, p8) <= 0)
                        R = M;
                    else
                        L = M + 1;
                }
                lo = L;
                R = hi;
                while (L < R) {
                    const i64 M = L + (R - L) / 2;
                    if (compare_at<SAT>(c, M, kw, m, p8) == 0)
                        L = M + 1;
                    else
                        R = M;
                }
                hi = L;
            }
            // a keyword of exactly k symbols is one bucket of the directory: its row is listed (unless the bucket is too long)
            if (c.listed && m == (i64)k && lo < hi && (e_lo >> 63))
                pre = kPreListed | (((e_lo >> 62) & 1) ? kPreRepeat : 0ull) | ((e_lo >> kPtCountShift) & kPtCountMask);
        }
        left_out[q] = lo;
        right_out[q] = hi;
        return hi - lo;
    }
    // src/index.cpp:262-274
    i64 L = 0, R = c.n - 1;
    while (L < R) {
        const i64 M = L + (R - L) / 2;
        if (compare_at<SAT>(c, M, kw, m, p8) <= 0)
            R = M;
        else
            L = M + 1;
    }
    const i64 left = L;
    // src/index.cpp:275-287
    L = left - 1;
    R = c.n - 1;
    while (L < R) {
        const i64 M = L + (R - L + 1) / 2;
        if (compare_at<SAT>(c, M, kw, m, p8) == 0)
            L = M;
        else
            R = M - 1;
    }
    const i64 right = L + 1 > left ? L + 1 : left;
    left_out[q] = left;
    right_out[q] = right;
    return right - left;
}

template <typename SAT>
__global__ void __launch_bounds__(256) search_kernel(SearchCtx c, SymTab tab, const u8* __restrict__ pat,
                                                     const i64* __restrict__ pat_off, i64 npat,
                                                     i64* __restrict__ left_out, i64* __restrict__ right_out,
                                                     int* __restrict__ err, u32* __restrict__ large_list,
                                                     unsigned long long* __restrict__ counters, u64* __restrict__ wocc,
                                                     u64* __restrict__ pre_out, i64* __restrict__ gright) {
    __shared__ u16 s_tab[256];
    s_tab[threadIdx.x] = tab.sym[threadIdx.x];
    __syncthreads();
    const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 occ = 0;
    u64 pre = 0;
    if (q < npat) occ = search_one<SAT>(c, pat, pat_off, q, left_out, right_out, err, s_tab, pre);
    if (counters == nullptr) return;
    // classification: long intervals go to the large path, listed rows are streamed from the document listing; the rest
    // is summed (capacity of the result buffer)
    if (occ > kWarpCap) {
        large_list[atomicAdd(counters + 0, 1ull)] = (u32)q;
        occ = 0;
    }
    unsigned long long locc = 0, ld = 0;
    if (pre_out) {
        if (pre) {  // phase A sees an empty interval for this row
            locc = (unsigned long long)occ;
            ld = (unsigned long long)(pre & kPreCount);
            occ = 0;
        }
        if (q < npat) {
            pre_out[q] = pre;
            gright[q] = pre ? left_out[q] : right_out[q];
        }
    }
    if (q < npat) wocc[q] = (u64)occ;  // occurrences the warp path will read for this pattern (0: large path, listed or no hit)
    unsigned long long s = (unsigned long long)occ, mx = s;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = t > mx ? t : mx;
    }
    if ((threadIdx.x & 31) == 0 && s) {
        atomicAdd(counters + 1, s);
        atomicMax(counters + 5, mx);  // longest interval on the warp path: picks the gather_kernel variant
    }
    if (pre_out && __any_sync(0xffffffffu, pre != 0)) {
        unsigned long long nrow = pre ? 1ull : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            locc += __shfl_xor_sync(0xffffffffu, locc, o);
            ld += __shfl_xor_sync(0xffffffffu, ld, o);
            nrow += __shfl_xor_sync(0xffffffffu, nrow, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 3, ld);    // result pairs of the listed rows (exact)
            atomicAdd(counters + 6, nrow);  // listed rows
            atomicAdd(counters + 7, locc);  // their occurrences
        }
    }
}

// ---- prefix directory ------------------------------------------------------------------------------------------
// ptab[c] = first SA rank whose k-symbol code (end-of-document = 0 padding) is >= c; one thread per entry runs a
// lower bound over the finished array.  Neighbouring entries follow almost the same path, so the probes hit L1/L2.
template <typename SAT>
__global__ void __launch_bounds__(256) ptab_kernel(SearchCtx c, SymTab tab, int b, int k, u64 nentries, u64* __restrict__ ptab) {
    __shared__ u16 s_tab[256];
    s_tab[threadIdx.x] = tab.sym[threadIdx.x];
    __syncthreads();
    const u64 code = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (code > nentries) return;
    if (code == nentries) {
        ptab[code] = (u64)c.n;
        return;
    }
    i64 L = 0, R = c.n;
    while (L < R) {
        const i64 M = L + (R - L) / 2;
        const u64 e = (u64) reinterpret_cast<const SAT*>(c.sa)[M];
        const i64 doc = (i64)(e & c.mask);
        const i64 s = __ldg(c.doc_off + doc) + (i64)(e >> c.bits1);
        const i64 slen = __ldg(c.doc_off + doc + 1) - s;
        const u64 t = load_be64(c.text, s);
        u64 sc = 0;
        for (int j = 0; j < k; ++j) sc = (sc << b) | (u64)((i64)j < slen ? s_tab[(t >> (56 - 8 * j)) & 255] : 0);
        if (sc >= code)
            R = M;
        else
            L = M + 1;
    }
    ptab[code] = (u64)L;
}

void build_prefix_table(Index& ix, cudaStream_t st) {
    if (ix.d_ptab) cudaFree(ix.d_ptab);
    ix.d_ptab = nullptr;
    ix.pt_k = ix.pt_b = 0;
    // note N1: on corpora mixing bytes < 0x80 and >= 0x80 the compat layout is not sorted under the comparator the
    // search uses, and the reference's exact recurrences must run on it
    if (ix.mixed && ix.opt.compat_signed && ix.n > ix.chuck_size) return;
    if (ix.n < 64 || ix.sigma == 0) return;
    int budget = 0;
    while (budget < 26 && ((i64)8 << (budget + 1)) <= ix.n) ++budget;  // <= n/8 entries, <= 512 MB
    if (const char* e = getenv("CDB_PTAB_BITS")) budget = atoi(e);
    int b = 1;
    while ((1 << b) <= ix.sigma) ++b;  // symbols 0..sigma
    int k = budget / b;
    if (k > 8) k = 8;  // one 8-byte text window per probe
    if (k < 1) return;
    const u64 nentries = 1ull << (b * k);
    CDB_CUDA(cudaMalloc(&ix.d_ptab, (nentries + 1) * 8));
    SearchCtx c = make_ctx(ix);
    c.ptab = nullptr;
    const unsigned grid = (unsigned)ceil_div((i64)nentries + 1, 256);
    if (ix.width == 4)
        ptab_kernel<u32><<<grid, 256, 0, st>>>(c, ix.symtab, b, k, nentries, ix.d_ptab);
    else
        ptab_kernel<u64><<<grid, 256, 0, st>>>(c, ix.symtab, b, k, nentries, ix.d_ptab);
    CDB_LAUNCH_CHECK();
    ix.pt_b = b;
    ix.pt_k = k;
}

// ---- K5-K7 gather (phase A) + translate (phase B) ---------------------------------------------------------------


// Loads SA[l, l+occ) (coalesced, all R loads of a lane in flight at once), reduces to doc indices, sorts them in
// registers and run-length encodes them straight from the registers: the distinct docs go to s_doc[0 .. nheads) in
// ascending order, the rank of each run's first element to s_pos (s_pos[nheads] = occ), both indexed through
// pad_idx.  Returns nheads; all_distinct tells that every run has length 1 (s_pos is then not written).
template <typename SAT, int R>
__device__ __forceinline__ int load_sort_rle(const SAT* __restrict__ sa, i64 l, int occ, u64 mask, u32 bucket_mul,
                                             u32* s_doc, u32* s_pos, int lane, bool& all_distinct,
                                             const u32* __restrict__ remap = nullptr) {
    SAT v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        v[r] = i < occ ? ld_stream(sa + l + i) : (SAT)0;
    }
    u32 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        x[r] = i < occ ? (u32)((u64)v[r] & mask) : 0xffffffffu;  // doc index <= 2^32-2 (bits1 <= 32)
    }
    if (remap) {  // id order: sort by the rank of the document's id (a bijection, so runs of equal keys are the same runs)
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r * 32 + lane < occ) x[r] = __ldg(remap + x[r]);
    }
    // s_doc / s_pos double as the sorts' scratch (33*R words each) before they are filled
    bool sorted = false;
    if constexpr (R >= kBucketMinR) {
        if (bucket_mul) sorted = warp_bucket_sort<R>(x, occ, bucket_mul, s_doc, s_pos, lane);
    }
    if (!sorted) warp_bitonic_regs<R>(x, lane, s_doc);
    // lane holds ranks lane*R .. lane*R + R-1; a rank is a run head when its doc differs from the previous rank's
    const u32 prev_last = __shfl_up_sync(0xffffffffu, x[R - 1], 1);
    u32 hm = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int idx = lane * R + r;
        const u32 pk = r ? x[r > 0 ? r - 1 : 0] : prev_last;
        const bool head = idx < occ && (idx == 0 || x[r] != pk);
        hm |= head ? (1u << r) : 0u;
    }
    const int nh = __popc(hm);
    int incl = nh;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int nheads = __shfl_sync(0xffffffffu, incl, 31);
    all_distinct = nheads == occ;
    if (all_distinct) {
        // the common case (no document hit twice): the compact list is the sorted list itself, every count is 1 —
        // R stores at fixed offsets (rank lane*R + r sits at pad_idx = lane*R + r + (lane*R >> 5)), no s_pos
        u32* dst = s_doc + lane * R + ((lane * R) >> 5);
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r] = x[r];
        __syncwarp();
        return nheads;
    }
    // compact list index j lives at pad_idx(j): when there are no repeats lane l writes j = l*R + r, and without the
    // padding all lanes of one store would hit the same bank
    int o = incl - nh;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if ((hm >> r) & 1u) {
            s_doc[pad_idx(o)] = x[r];
            s_pos[pad_idx(o)] = (u32)(lane * R + r);
            ++o;
        }
    }
    if (lane == 0) s_pos[pad_idx(nheads)] = (u32)occ;
    __syncwarp();
    return nheads;
}

// Phase A.  One warp per pattern, no communication between warps: read the SA interval, reduce to doc indices, sort,
// run-length encode.  The compact row (u32 docs in cdocs; u16 counts in ccnt only when some document is hit more than
// once, flagged in bit 15 of the row's seg entries) goes to alloc_off[q] — the exclusive scan of the
// warp-path occurrence counts, known right after the search, an upper bound of the row lengths — and the exact row
// length to rowlen[q]; a scan of rowlen gives the CSR offsets afterwards (no look-back chain, no CTA barrier: with one
// the kernel spent 27 % of its warp time waiting).  seg holds the row's split points at the doc-range boundaries for
// phase B.  Patterns on the large path report the row length computed there (dlarge) and write nothing else.
//   seg layout: [q / 8][r = 0..nranges][q % 8] u16, seg(q, r) = number of row entries with doc < (r << rshift)
// MAXR = 32: intervals up to kWarpCap, 3 CTAs per SM.  MAXR = 4: batches whose longest warp-path interval is <= 128
// occurrences (short rows: sharded corpora, long keywords) — the same code with the long sorting networks compiled
// out, half the registers and a tenth of the shared memory, so twice as many warps per SM hide the latency.
template <typename SAT, int MAXR, bool REMAP>
__global__ void __launch_bounds__(kTileWarps * 32, MAXR <= 4 ? 6 : MAXR == 8 ? 5 : MAXR == 16 ? 4 : 3) gather_kernel(const SAT* __restrict__ sa, u64 mask,
                                                                                    u32 bucket_mul,
                                                                                    const i64* __restrict__ left,
                                                                                    const i64* __restrict__ right, i64 npat,
                                                                                    const u64* __restrict__ dlarge,
                                                                                    const u64* __restrict__ alloc_off,
                                                                                    u64* __restrict__ rowlen,
                                                                                    u32* __restrict__ cdocs,
                                                                                    u16* __restrict__ ccnt,
                                                                                    u16* __restrict__ seg, int nranges,
                                                                                    int rshift, const u32* __restrict__ remap_arg) {
    const u32* const remap = REMAP ? remap_arg : nullptr;  // compiled out of the doc-order kernel (its register budget is tight)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Short-row variants (MAXR <= 8: the shards of a split corpus, long keywords) run as a persistent grid: a warp walks
    // the patterns q, q + stride, ... and has the next pattern's interval and row offset in flight while it sorts the
    // current one — with one CTA per 8 patterns the kernel was bound by CTA turnover and three dependent round trips
    // per warp (0.94 ms per 10^6 rows of ~100 entries).  The long-row variants keep one pattern per warp.
    constexpr bool LOOP = MAXR <= 8;
    const i64 stride = LOOP ? (i64)gridDim.x * kTileWarps : npat;
    i64 q = (i64)blockIdx.x * kTileWarps + warp;
    if (q >= npat) return;
    u32* s_doc = reinterpret_cast<u32*>(smem_raw + (size_t)warp * warp_smem_bytes<MAXR>());
    u32* s_pos = s_doc + 32 * MAXR + 32;
    i64 l = left[q], rgt = right[q];
    u64 row = alloc_off[q];
    for (;;) {
    i64 l_next = 0, rgt_next = 0;
    u64 row_next = 0;
    const i64 q_next = q + stride;
    if (LOOP && q_next < npat) {
        l_next = left[q_next];
        rgt_next = right[q_next];
        row_next = alloc_off[q_next];
    }
    int nheads = 0;
    bool all_distinct = false;
    u64 d = 0;
    const i64 occ64 = rgt - l;
    if (occ64 > kWarpCap) {
        d = dlarge[q];
    } else if (occ64 > 0) {
        const int occ = (int)occ64;
        if (occ <= 32) nheads = load_sort_rle<SAT, 1>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct, remap);
        else if (occ <= 64) nheads = load_sort_rle<SAT, 2>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct, remap);
        else if (MAXR <= 4 || occ <= 128) nheads = load_sort_rle<SAT, 4>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct, remap);
        else if (occ <= 256) nheads = load_sort_rle<SAT, (MAXR >= 8 ? 8 : 4)>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct, remap);
        else if (occ <= 512) nheads = load_sort_rle<SAT, (MAXR >= 16 ? 16 : 4)>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct, remap);
        else nheads = load_sort_rle<SAT, (MAXR >= 32 ? 32 : 4)>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct, remap);
        d = (u64)nheads;
    }
    if (lane == 0) rowlen[q] = d;
    // split points of the row at the doc-range boundaries (nheads == 0 for empty / large-path rows -> all zero)
    {
        u16* g = seg + (size_t)(q / kTileWarps) * (nranges + 1) * kTileWarps + (q % kTileWarps);
        for (int r = lane; r <= nranges; r += 32) {
            const u64 bound = (u64)r << rshift;
            int lo = 0, hi = nheads;  // first j with doc_j >= bound
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((u64)s_doc[pad_idx(mid)] < bound)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            g[r * kTileWarps] = (u16)(lo | ((nheads && !all_distinct) ? 0x8000 : 0));
        }
    }
    // compact row: coalesced 4-byte doc stores (+ 2-byte counts for the rare rows with repeats)
    for (int r = lane; r < nheads; r += 32) st_stream_u32(cdocs + row + r, s_doc[pad_idx(r)]);
    if (!all_distinct)
        for (int r = lane; r < nheads; r += 32) ccnt[row + r] = (u16)(s_pos[pad_idx(r + 1)] - s_pos[pad_idx(r)]);
    if (!LOOP || q_next >= npat) break;
    q = q_next;
    l = l_next;
    rgt = rgt_next;
    row = row_next;
    __syncwarp();  // s_doc / s_pos are reused by the next pattern
    }
}

// Phase B.  pairs[i] = (ids[doc_i], count_i) for every compact entry i.  ids[] (8 bytes per document, 800 MB at the
// 10 GB configuration) is read at random, so the work is ordered by doc range: item = (range r, block of 32
// patterns), all warps of the grid walk the items in order and therefore look up the same <= 32 MB slice of ids[]
// at the same time, which stays L2-resident (evict_last), while the compact rows and the result stream through
// (evict_first).  One warp per item; the lanes are spread evenly over the concatenated row segments of the
// item's 32 patterns (binary search over the segment prefix sums), kTrU independent lookups in flight per lane.
constexpr int kTrU = 4;

// kTrU x 32 consecutive entries of the item, starting at flat index i0.  FULL: all of them exist.
template <bool FULL>
__device__ __forceinline__ void translate_rounds(const u32* __restrict__ cdocs, const u16* __restrict__ ccnt,
                                                 const u64* __restrict__ ids,
                                                 i64* __restrict__ pairs, const u32* s_excl, const u64* s_in,
                                                 const u64* s_out, u32 i0, u32 tot, int lane, u64 pol_keep, u64 pol_stream) {
    u64 p[kTrU];   // where the (id, count) pair goes
    u32 doc[kTrU];
    u32 cnt[kTrU];
#pragma unroll
    for (int u = 0; u < kTrU; ++u) {
        const u32 idx = i0 + u * 32 + lane;
        if (FULL || idx < tot) {
            int j = 0;  // largest j with excl[j] <= idx
#pragma unroll
            for (int st = 16; st; st >>= 1)
                if (s_excl[j + st] <= idx) j += st;
            p[u] = s_out[j] + idx;
            const u64 in = s_in[j];  // bit 63: the row has repeated documents, its counts are in ccnt
            const u64 src = (in & ~(1ull << 63)) + idx;
            doc[u] = ld_hint_u32(cdocs + src, pol_stream);
            cnt[u] = (in >> 63) ? (u32)__ldg(ccnt + src) : 1u;
        }
    }
    // ptxas otherwise sinks every load next to its use and runs the kTrU chains one after another (one request in
    // flight per lane); a warp barrier between the phases keeps the kTrU loads of a phase together
    __syncwarp();
    longlong2 v[kTrU];
#pragma unroll
    for (int u = 0; u < kTrU; ++u) {
        const u32 idx = i0 + u * 32 + lane;
        if (FULL || idx < tot) {
            v[u].x = (i64)ld_hint_u64(ids + doc[u], pol_keep);
            v[u].y = (i64)cnt[u];
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < kTrU; ++u) {
        const u32 idx = i0 + u * 32 + lane;
        if (FULL || idx < tot) st_hint_v2(pairs + 2 * p[u], v[u], pol_stream);
    }
}

__global__ void __launch_bounds__(kTrWarps * 32) translate_kernel(const u32* __restrict__ cdocs,
                                                                  const u16* __restrict__ ccnt,
                                                                  const u64* __restrict__ alloc_off,
                                                                  const u64* __restrict__ row_off,
                                                                  const u16* __restrict__ seg,
                                                                  const i64* __restrict__ ids, i64* __restrict__ pairs,
                                                                  i64 npat, int nranges, unsigned long long* ticket,
                                                                  int keep_mode) {
    __shared__ u32 s_excl[kTrWarps][32];
    __shared__ u64 s_in[kTrWarps][32];   // compact rows sit at alloc_off (upper-bound offsets) ...
    __shared__ u64 s_out[kTrWarps][32];  // ... the result rows at row_off (exact CSR offsets)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 ntile = (npat + 31) >> 5;
    const i64 nitems = ntile * nranges;
    // keep_mode (CDB_TRANSLATE_KEEP): priority of the ids[] slice in L2 — 0 evict_last, 1 evict_normal, 2 evict_unchanged
    const u64 pol_keep = keep_mode == 1 ? l2_policy_evict_normal() : keep_mode == 2 ? l2_policy_evict_unchanged() : l2_policy_evict_last();
    const u64 pol_stream = l2_policy_evict_first();
    // Items are handed out through one global ticket, in order: at any moment all warps of the grid are within a few
    // thousand items of each other, i.e. inside one (at a range change: two) slices of ids[].  A static item
    // stride lets warps drift several ranges apart and the slices in use no longer fit L2 together (measured:
    // 45.6 GB of DRAM reads per 10^6 patterns at cfg3 against 15.6 GB with the ticket).
    for (;;) {
        i64 item = 0;
        if (lane == 0) item = (i64)atomicAdd(ticket, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) break;
        const int r = (int)(item / ntile);
        const i64 q = (item - (i64)r * ntile) * 32 + lane;
        u32 len = 0;
        u64 base_in = 0, base_out = 0;
        if (q < npat) {
            const u16* sg = seg + ((size_t)(q / kTileWarps) * (nranges + 1) + r) * kTileWarps + (q % kTileWarps);
            const u32 s = sg[0] & 0x7fffu, e = sg[kTileWarps] & 0x7fffu;
            len = e - s;
            if (len) {
                base_in = (alloc_off[q] + s) | ((u64)(sg[0] >> 15) << 63);
                base_out = row_off[q] + s;
            }
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
        s_in[warp][lane] = base_in - (incl - len);  // the flag in bit 63 survives: offsets stay far below 2^63
        s_out[warp][lane] = base_out - (incl - len);
        __syncwarp();
        const u64* idp = reinterpret_cast<const u64*>(ids);
        u32 i0 = 0;
        for (; i0 + 32 * kTrU <= tot; i0 += 32 * kTrU)
            translate_rounds<true>(cdocs, ccnt, idp, pairs, s_excl[warp], s_in[warp], s_out[warp], i0, tot, lane, pol_keep, pol_stream);
        if (i0 < tot)
            translate_rounds<false>(cdocs, ccnt, idp, pairs, s_excl[warp], s_in[warp], s_out[warp], i0, tot, lane, pol_keep, pol_stream);
    }
}

// ---- K5-K7 large path -------------------------------------------------------------------------------------------
__global__ void large_occ_kernel(const u32* __restrict__ list, u64 nl, const i64* __restrict__ left,
                                 const i64* __restrict__ right, u64* __restrict__ occ) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nl) occ[j] = (u64)(right[list[j]] - left[list[j]]);
}

template <typename SAT>
__global__ void large_expand_kernel(const SAT* __restrict__ sa, u64 mask, const u32* __restrict__ list, u64 nl,
                                    const i64* __restrict__ left, const u64* __restrict__ ooff, u64 total,
                                    u64* __restrict__ keys, const u32* __restrict__ remap) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (u64)gridDim.x * blockDim.x) {
        u64 lo = 0, hi = nl - 1;  // largest j with ooff[j] <= e
        while (lo < hi) {
            u64 mid = lo + (hi - lo + 1) / 2;
            if (__ldg(ooff + mid) <= e)
                lo = mid;
            else
                hi = mid - 1;
        }
        const u64 i = (u64)left[list[lo]] + (e - __ldg(ooff + lo));
        const u64 doc = (u64)sa[i] & mask;
        keys[e] = (lo << 32) | (remap ? (u64)__ldg(remap + doc) : doc);
    }
}

__global__ void large_flag_kernel(const u64* __restrict__ keys, u64 total, u8* __restrict__ flags) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (u64)gridDim.x * blockDim.x)
        flags[e] = (e == 0 || keys[e] != keys[e - 1]) ? 1 : 0;
}

__global__ void large_unique_kernel(const u64* __restrict__ keys, const u8* __restrict__ flags,
                                    const u64* __restrict__ pos, u64 total, u64* __restrict__ ukey,
                                    u64* __restrict__ ustart, u64* __restrict__ entry_first) {
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (u64)gridDim.x * blockDim.x) {
        if (!flags[e]) continue;
        const u64 u = pos[e], k = keys[e];
        ukey[u] = k;
        ustart[u] = e;
        if (e == 0 || (keys[e - 1] >> 32) != (k >> 32)) entry_first[k >> 32] = u;
    }
}

__global__ void large_dcount_kernel(const u32* __restrict__ list, u64 nl, const u64* __restrict__ entry_first,
                                    u64* __restrict__ dcount) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nl) dcount[list[j]] = entry_first[j + 1] - entry_first[j];
}

__global__ void large_emit_kernel(const u64* __restrict__ ukey, const u64* __restrict__ ustart, u64 nu, u64 total,
                                  const u32* __restrict__ list, const u64* __restrict__ entry_first,
                                  const u64* __restrict__ row_off, const i64* __restrict__ ids, i64* __restrict__ pairs) {
    for (u64 u = (u64)blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += (u64)gridDim.x * blockDim.x) {
        const u64 k = ukey[u];
        const u64 j = k >> 32;
        const u64 cnt = (u + 1 < nu ? ustart[u + 1] : total) - ustart[u];
        const u64 dst = row_off[list[j]] + (u - entry_first[j]);
        longlong2 v;
        v.x = __ldg(ids + (k & 0xffffffffull));
        v.y = (i64)cnt;
        *reinterpret_cast<longlong2*>(pairs + 2 * dst) = v;
    }
}

static int bits_for_u64(u64 v) {
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

// a[0..n) -> its exclusive scan, a[n] = total.  Small batches (a query() per request) take one single-block launch.
static void scan_in_place(u64* a, u64 n, cudaStream_t st) {
    if (n <= 16384) {
        prim::scan_blocksums_kernel<<<1, 1024, 0, st>>>(a, n);
        CDB_LAUNCH_CHECK();
    } else {
        prim::exclusive_scan<u64>(a, a, n, st);
    }
}

// ---- id order (filter) --------------------------------------------------------------------------------------------
__global__ void ids_descent_kernel(const i64* __restrict__ ids, i64 nd, int* __restrict__ flag) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 < nd && ids[i] >= ids[i + 1]) *flag = 1;
}
__global__ void rank_keys_kernel(const i64* __restrict__ ids, i64 nd, u64* __restrict__ keys, u32* __restrict__ docs) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nd) {
        keys[i] = (u64)ids[i] ^ (1ull << 63);  // signed order as unsigned
        docs[i] = (u32)i;
    }
}
__global__ void rank_scatter_kernel(const u64* __restrict__ keys, const u32* __restrict__ docs, i64 nd,
                                    u32* __restrict__ rank_tab, i64* __restrict__ ids_by_rank) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nd) {
        rank_tab[docs[r]] = (u32)r;
        ids_by_rank[r] = (i64)(keys[r] ^ (1ull << 63));
    }
}

template <typename SAT>
__global__ void __launch_bounds__(256) sa_rank_kernel(const SAT* __restrict__ sa, u64 mask, const u32* __restrict__ rank_tab, i64 n,
                                                      u32* __restrict__ out) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        u32 d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) d[u] = i + u * stride < n ? (u32)((u64)ld_stream(sa + i + u * stride) & mask) : 0u;
        u32 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(rank_tab + d[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * stride < n) out[i + u * stride] = v[u];
    }
}

// Examines the ids once per built index.  Returns true when the rows have to be re-keyed (ids do not ascend with the
// doc index): rank_tab / ids_by_rank are then filled.  Ids are assumed distinct, as the reference's are (object ids).
static bool id_order_tables(const Index& ix, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(ix.order_mu);
    if (ix.ids_order < 0) {
        const i64 nd = ix.nd;
        const unsigned grid = (unsigned)ceil_div(nd > 0 ? nd : 1, 256);
        DevBuf<int> flag(1, st);
        CDB_CUDA(cudaMemsetAsync(flag.p, 0, 4, st));
        ids_descent_kernel<<<grid, 256, 0, st>>>(ix.d_ids, nd, flag.p);
        CDB_LAUNCH_CHECK();
        int h = 0;
        CDB_CUDA(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        if (h) {
            BigBuf<u64> k0((size_t)nd), k1((size_t)nd);
            BigBuf<u32> v0((size_t)nd), v1((size_t)nd);
            rank_keys_kernel<<<grid, 256, 0, st>>>(ix.d_ids, nd, k0.p, v0.p);
            CDB_LAUNCH_CHECK();
            const int cur = rs::radix_sort_pairs<u32>(k0.p, k1.p, v0.p, v1.p, (u64)nd, 0, 64, st);
            CDB_CUDA(cudaMalloc((void**)&ix.d_rank_tab, (size_t)nd * 4));
            CDB_CUDA(cudaMalloc((void**)&ix.d_ids_by_rank, (size_t)nd * 8));
            rank_scatter_kernel<<<grid, 256, 0, st>>>(cur ? k1.p : k0.p, cur ? v1.p : v0.p, nd, ix.d_rank_tab, ix.d_ids_by_rank);
            CDB_LAUNCH_CHECK();
            CDB_CUDA(cudaStreamSynchronize(st));
            // Rank companion of the suffix array: sa_rank[i] = rank_tab[sa[i] & mask], 4 bytes per suffix.  With it an
            // id-ordered gather reads a plain contiguous interval (half the bytes of the packed array's) instead of one
            // random rank_tab lookup per occurrence (measured at the 10 GB configuration: 28 ms against 3.7 ms per 10^6
            // keywords).  Only taken when it leaves a quarter of the device memory free for the query temporaries.
            const char* ec = getenv("CDB_SA_RANK_COMPANION");
            if (ix.n > 0 && (!ec || atoi(ec) != 0)) {
                // memory held by the stream-ordered pool for query temporaries counts as available (it is re-used, not
                // lost), but is only handed back to the driver when the allocation does not fit otherwise
                const size_t need = (size_t)ix.n * 4;
                size_t free_b = 0, total_b = 0;
                CDB_CUDA(cudaMemGetInfo(&free_b, &total_b));
                cudaMemPool_t pool;
                const bool have_pool = cudaDeviceGetDefaultMemPool(&pool, ix.device) == cudaSuccess;
                unsigned long long reserved = 0, used = 0;
                if (have_pool) {
                    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
                    cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
                }
                const size_t avail = free_b + (size_t)(reserved > used ? reserved - used : 0);
                bool ok = avail > need && avail - need >= total_b / 4;
                if (ok && cudaMalloc((void**)&ix.d_sa_rank, need) != cudaSuccess) {
                    cudaGetLastError();
                    if (have_pool) cudaMemPoolTrimTo(pool, 0);
                    ok = cudaMalloc((void**)&ix.d_sa_rank, need) == cudaSuccess;
                }
                if (ok) {
                    const int g = num_sms() * 16;
                    if (ix.width == 4)
                        sa_rank_kernel<u32><<<g, 256, 0, st>>>(reinterpret_cast<const u32*>(ix.d_sa), ix.mask, ix.d_rank_tab, ix.n, ix.d_sa_rank);
                    else
                        sa_rank_kernel<u64><<<g, 256, 0, st>>>(reinterpret_cast<const u64*>(ix.d_sa), ix.mask, ix.d_rank_tab, ix.n, ix.d_sa_rank);
                    CDB_LAUNCH_CHECK();
                    CDB_CUDA(cudaStreamSynchronize(st));
                } else {
                    cudaGetLastError();
                    ix.d_sa_rank = nullptr;
                }
            }
        }
        ix.ids_order = h ? 0 : 1;
    }
    return ix.ids_order == 0;
}

// per-pattern (row length, occurrences) as 32-bit integers, [2][npat]: what a sharded index exchanges per batch
// + row_flags: bit 0 = the row's counts are not all 1 (gather_kernel left that in bit 15 of the row's seg entries;
// rows of the large path: not known, flagged)
__global__ void stats32_kernel(const u64* __restrict__ row_off, const i64* __restrict__ left, const i64* __restrict__ right,
                               i64 npat, const u16* __restrict__ seg, int nranges, const u64* __restrict__ pre,
                               int32_t* __restrict__ stats, u8* __restrict__ row_flags) {
    const i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npat) return;
    const u64 rl = row_off[q + 1] - row_off[q];
    const i64 oc = right[q] - left[q];
    const u64 p = pre ? pre[q] : 0;
    bool rep;
    if (p & kPreListed)
        rep = (p & kPreRepeat) != 0;
    else if (oc > kWarpCap)
        rep = true;
    else
        rep = seg ? (seg[(size_t)(q / kTileWarps) * (nranges + 1) * kTileWarps + (q % kTileWarps)] >> 15) != 0 : false;
    row_flags[q] = rep ? 1 : 0;
    stats[q] = rl > 0x7fffffffull ? 0x7fffffff : (int32_t)rl;
    stats[npat + q] = oc > 0x7fffffffll ? 0x7fffffff : (int32_t)oc;
}

// ---- host driver ------------------------------------------------------------------------------------------------
template <typename SAT>
static void locate_typed(const Index& ix, const u8* d_pat, const i64* d_pat_off, i64 npat, cudaStream_t st,
                         cdb_device_result* out, bool id_order, cdb_rows_ready_fn rows_ready, void* rows_ready_user,
                         LazyListed* lazy) {
    const SAT* sa = reinterpret_cast<const SAT*>(ix.d_sa);
    // id order: keys are id ranks (rank_tab) and the table translate reads is ids_by_rank; when the ids already ascend with
    // the doc index both orders coincide and nothing changes
    const u32* remap = nullptr;
    const u32* sa_rank = nullptr;
    const i64* ids_tab = ix.d_ids;
    int order = 0;
    if (id_order && id_order_tables(ix, st)) {
        order = 1;
        ids_tab = ix.d_ids_by_rank;
        if (ix.d_sa_rank)
            sa_rank = ix.d_sa_rank;  // sa_rank[i] = rank of the id of the document suffix i belongs to
        else
            remap = ix.d_rank_tab;   // no room for the companion: the ranks are looked up per occurrence (slower)
    }
    // document listing of this order (held for the duration of the call): keywords of exactly pt_k symbols are streamed from it
    // (CDB_LISTING_USE=0, read per call: this call takes the suffix-array path whatever the index has — A/B measurements)
    const char* env_use = getenv("CDB_LISTING_USE");
    const std::shared_ptr<Listing> lst = (env_use && atoi(env_use) == 0) ? std::shared_ptr<Listing>() : get_listing(ix, order, st);
    const i64 ntiles = ceil_div(npat, kTileWarps);
    DevBuf<i64> left(npat, st), right(npat, st);
    DevBuf<u64> row_off(npat + 1, st);
    DevBuf<u64> dlarge;                          // row counts of large-path patterns (only allocated when needed)
    DevBuf<unsigned long long> counters(8, st);  // [0] large patterns, [1] occurrences on the warp path, [2] err, [3] pairs of the listed rows, [4] translate ticket, [5] longest warp-path interval, [6] listed rows, [7] their occurrences
    DevBuf<u32> large_list(npat, st);
    CDB_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), st));
    cudaEvent_t* ev = thread_ctx(ix.device).ev;  // 7 of the thread's cached timing events
    CDB_CUDA(cudaEventRecord(ev[0], st));
    SearchCtx c = make_ctx(ix);
    DevBuf<u64> alloc_off(npat + 1, st);
    DevBuf<u64> pre;      // per pattern: answered from the listing (and its exact row length) or not
    DevBuf<i64> gright;   // right end of the interval as phase A sees it (== left for listed rows)
    if (lst) {
        c.listed = true;
        pre.alloc(npat, st);
        gright.alloc(npat, st);
    }
    int* err = reinterpret_cast<int*>(counters.p + 2);
    search_kernel<SAT><<<(unsigned)ceil_div(npat, 256), 256, 0, st>>>(c, ix.symtab, d_pat, d_pat_off, npat, left.p, right.p, err,
                                                                      large_list.p, counters.p, alloc_off.p, pre.p, gright.p);
    CDB_LAUNCH_CHECK();
    const i64* right_a = lst ? gright.p : right.p;
    // where every pattern's compact row goes: exclusive scan (in place) of the warp-path occurrence counts, which are
    // upper bounds of the row lengths
    scan_in_place(alloc_off.p, (u64)npat, st);
    CDB_CUDA(cudaEventRecord(ev[1], st));
    unsigned long long hc[8];
    CDB_CUDA(cudaMemcpyAsync(hc, counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    if ((int)(hc[2] & 0xffffffffu)) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");
    const u64 nl = hc[0];
    const u64 nlisted = hc[6], listed_pairs = hc[3];
    u64 total_occ = hc[1] + hc[7];
    // a batch answered from the listing alone (no warp-path occurrence, no long interval) skips phases A and B
    const bool general = hc[1] > 0 || nl > 0 || nlisted == 0;
    CDB_CUDA(cudaEventRecord(ev[2], st));
    // large path, phase 1: exact row counts of the long intervals.  The occurrences of the large patterns are
    // expanded, sorted and run-length encoded in sub-batches of bounded size (25 bytes of temporaries per occurrence);
    // what is kept until the rows can be emitted is 16 bytes per (pattern, document) pair.
    struct LargeChunk {
        u64 j0 = 0, nlc = 0, nu = 0, ltotal = 0;
        DevBuf<u64> ukey, ustart, entry_first;
    };
    std::vector<LargeChunk> lchunks;
    u64 nu = 0;
    if (nl > 0) {
        dlarge.alloc(npat, st);
        DevBuf<u64> ooff(nl + 1, st);
        large_occ_kernel<<<(unsigned)ceil_div((i64)nl, 256), 256, 0, st>>>(large_list.p, nl, left.p, right.p, ooff.p);
        CDB_LAUNCH_CHECK();
        prim::exclusive_scan<u64>(ooff.p, ooff.p, nl, st);
        std::vector<u64> h_ooff(nl + 1);
        CDB_CUDA(cudaMemcpyAsync(h_ooff.data(), ooff.p, (nl + 1) * 8, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        total_occ += h_ooff[nl];
        u64 limit = 1ull << 28;
        if (const char* e = getenv("CDB_LARGE_LIMIT")) limit = (u64)atoll(e);
        for (u64 j0 = 0; j0 < nl;) {
            u64 j1 = j0 + 1;  // at least one pattern, then as many as fit the limit
            while (j1 < nl && h_ooff[j1 + 1] - h_ooff[j0] <= limit) ++j1;
            lchunks.emplace_back();
            LargeChunk& lc = lchunks.back();
            lc.j0 = j0;
            lc.nlc = j1 - j0;
            lc.ltotal = h_ooff[j1] - h_ooff[j0];
            std::vector<u64> h_loff(lc.nlc + 1);
            for (u64 j = 0; j <= lc.nlc; ++j) h_loff[j] = h_ooff[j0 + j] - h_ooff[j0];
            DevBuf<u64> loff(lc.nlc + 1, st);
            CDB_CUDA(cudaMemcpyAsync(loff.p, h_loff.data(), (lc.nlc + 1) * 8, cudaMemcpyHostToDevice, st));
            DevBuf<u64> k0(lc.ltotal, st), k1(lc.ltotal, st);
            const int grid = (int)std::min<i64>(ceil_div((i64)lc.ltotal, 256), num_sms() * 16);
            if (sa_rank)
                large_expand_kernel<u32><<<grid, 256, 0, st>>>(sa_rank, 0xffffffffull, large_list.p + j0, lc.nlc, left.p, loff.p, lc.ltotal,
                                                               k0.p, nullptr);
            else
                large_expand_kernel<SAT><<<grid, 256, 0, st>>>(sa, ix.mask, large_list.p + j0, lc.nlc, left.p, loff.p, lc.ltotal, k0.p, remap);
            CDB_LAUNCH_CHECK();
            int cbuf = rs::radix_sort_pairs<rs::NoValue>(k0.p, k1.p, nullptr, nullptr, lc.ltotal, 0,
                                                         32 + bits_for_u64(lc.nlc - 1), st);
            u64* sorted = cbuf ? k1.p : k0.p;
            DevBuf<u8> flags(lc.ltotal, st);
            large_flag_kernel<<<grid, 256, 0, st>>>(sorted, lc.ltotal, flags.p);
            CDB_LAUNCH_CHECK();
            DevBuf<u64> pos(lc.ltotal + 1, st);
            prim::exclusive_scan<u8>(flags.p, pos.p, lc.ltotal, st);
            CDB_CUDA(cudaMemcpyAsync(&lc.nu, pos.p + lc.ltotal, 8, cudaMemcpyDeviceToHost, st));
            CDB_CUDA(cudaStreamSynchronize(st));  // also: h_loff must stay alive until its upload has happened
            lc.ukey.alloc(lc.nu, st);
            lc.ustart.alloc(lc.nu, st);
            lc.entry_first.alloc(lc.nlc + 1, st);
            large_unique_kernel<<<grid, 256, 0, st>>>(sorted, flags.p, pos.p, lc.ltotal, lc.ukey.p, lc.ustart.p, lc.entry_first.p);
            CDB_LAUNCH_CHECK();
            CDB_CUDA(cudaMemcpyAsync(lc.entry_first.p + lc.nlc, pos.p + lc.ltotal, 8, cudaMemcpyDeviceToDevice, st));
            large_dcount_kernel<<<(unsigned)ceil_div((i64)lc.nlc, 256), 256, 0, st>>>(large_list.p + j0, lc.nlc, lc.entry_first.p,
                                                                                      dlarge.p);
            CDB_LAUNCH_CHECK();
            nu += lc.nu;
            j0 = j1;
        }
    }
    CDB_CUDA(cudaEventRecord(ev[3], st));
    // phase A: rows <= occurrences on the warp path + exact rows of the large path
    const u64 cap_warp = hc[1] + nu;           // compact rows of phase A
    const u64 cap_pairs = cap_warp + listed_pairs;
    int nranges, rshift;
    ids_ranges(ix.nd, &nranges, &rshift);
    // Few lookups compared with the size of ids[] (short rows, e.g. long keywords): one pass over all rows costs less
    // than nranges sparse ones, and there is nothing for L2 to keep.
    if (nranges > 1 && cap_warp * 8 < (u64)ix.nd && !getenv("CDB_RANGE_BITS")) {
        nranges = 1;
        rshift = 40;
    }
    DevBuf<u32> cdocs;
    DevBuf<u16> ccnt;  // only touched for rows with repeated documents
    DevBuf<u16> seg;
    if (general) {
        cdocs.alloc((size_t)cap_warp, st);
        ccnt.alloc((size_t)cap_warp, st);
        seg.alloc((size_t)ntiles * (nranges + 1) * kTileWarps, st);
    }
    DevBuf<i64> pairs((size_t)cap_pairs * 2, st);
    DevBuf<u8> rowflag((size_t)npat, st);
    // distribution-sort scale: bucket = doc * 1024 / nd (0 switches the sorting network on for every interval)
    const char* env_buckets = getenv("CDB_GATHER_BUCKETS");
    const bool use_buckets = !env_buckets || atoi(env_buckets) != 0;
    const u32 bucket_mul =
        use_buckets && ix.nd > 0 ? (u32)std::min<u64>(0xffffffffull, (1024ull << 32) / (u64)ix.nd) : 0u;
    // row_off first receives the exact row lengths, then becomes their exclusive scan = the CSR offsets
    // kernel variant by the batch's longest warp-path interval: shorter sorting arrays need fewer registers and less
    // shared memory, so more warps per SM hide the latency.  The 256- and 512-key variants serve the shards of a split
    // corpus (CDB_GATHER_VARIANTS=0 leaves only the 128- and 1024-key kernels).
    const char* env_var = getenv("CDB_GATHER_VARIANTS");
    const bool mid_variants = !env_var || atoi(env_var) != 0;
    auto launch_gather = [&](auto maxr_tag) {
        constexpr int MAXR = decltype(maxr_tag)::value;
        const size_t smem = (size_t)kTileWarps * warp_smem_bytes<MAXR>();
        if (sa_rank) {  // id order from the rank companion of the suffix array: a plain u32 array, nothing to look up
            auto kernel = gather_kernel<u32, MAXR, false>;
            if (smem > 48 * 1024) CDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            i64 grid = ntiles;
            if (MAXR <= 8) grid = std::min<i64>(ntiles, (i64)num_sms() * resident_ctas((const void*)kernel, kTileWarps * 32, smem));
            kernel<<<(unsigned)grid, kTileWarps * 32, smem, st>>>(sa_rank, 0xffffffffull, bucket_mul, left.p, right_a, npat, dlarge.p,
                                                                 alloc_off.p, row_off.p, cdocs.p, ccnt.p, seg.p, nranges, rshift, nullptr);
            return;
        }
        auto go = [&](auto kernel) {
            if (smem > 48 * 1024) CDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            i64 grid = ntiles;
            if (MAXR <= 8) {  // persistent grid: one wave of resident CTAs
                const int per_sm = resident_ctas((const void*)kernel, kTileWarps * 32, smem);
                grid = std::min<i64>(ntiles, (i64)num_sms() * per_sm);
            }
            kernel<<<(unsigned)grid, kTileWarps * 32, smem, st>>>(sa, ix.mask, bucket_mul, left.p, right_a, npat, dlarge.p, alloc_off.p,
                                                                 row_off.p, cdocs.p, ccnt.p, seg.p, nranges, rshift, remap);
        };
        if (remap)
            go(gather_kernel<SAT, MAXR, true>);
        else
            go(gather_kernel<SAT, MAXR, false>);
    };
    if (general) {
        if (hc[5] <= 128)
            launch_gather(std::integral_constant<int, 4>{});
        else if (mid_variants && hc[5] <= 256)
            launch_gather(std::integral_constant<int, 8>{});
        else if (mid_variants && hc[5] <= 512)
            launch_gather(std::integral_constant<int, 16>{});
        else
            launch_gather(std::integral_constant<int, 32>{});
        CDB_LAUNCH_CHECK();
    }
    if (nlisted) {  // the listed rows' lengths are known from the directory (phase A left 0 for them)
        launch_listing_rowlen(pre.p, npat, row_off.p, general ? 0 : 1, st);
    }
    scan_in_place(row_off.p, (u64)npat, st);
    // per-pattern (row length, occurrences) are final here: a sharded caller starts its exchange now, under translate
    DevBuf<int32_t> stats((size_t)npat * 2, st);
    stats32_kernel<<<(unsigned)ceil_div(npat, 256), 256, 0, st>>>(row_off.p, left.p, right.p, npat, seg.p, nranges, pre.p, stats.p, rowflag.p);
    CDB_LAUNCH_CHECK();
    // A batch answered from the listing alone has no translate to hide the exchange under, and the hook's host work (tens of
    // microseconds in a Python caller) would keep the device waiting for the emit kernel.  The emit kernel is therefore
    // enqueued FIRST and the hook runs afterwards on a side stream that only waits for the statistics (recorded here): the
    // caller's collective is in the queue at once and runs as soon as SMs are free.  (Leaving 64 of the emit grid's 592 CTAs
    // out to make room for it cost the emit kernel 10 % and bought nothing measurable on 2 GPUs: launch_listing_emit's
    // spare_ctas stays 0.)
    // CDB_HOOK_AFTER_EMIT=0 keeps the hook here, on the launching stream.
    const char* env_hook = getenv("CDB_HOOK_AFTER_EMIT");
    const bool hook_late = rows_ready && !general && nlisted && (!env_hook || atoi(env_hook) != 0);
    if (rows_ready && !hook_late) rows_ready(rows_ready_user, stats.p, npat, (void*)st);
    if (hook_late) CDB_CUDA(cudaEventRecord(ev[8], st));
    CDB_CUDA(cudaEventRecord(ev[4], st));
    // phase B: doc index -> id, ordered by doc range so that the ids[] slice in use is L2-resident
    if (general) {
        const i64 nitems = ceil_div(npat, 32) * nranges;
        const int per_sm = resident_ctas((const void*)translate_kernel, kTrWarps * 32);
        const int grid = (int)std::min<i64>(ceil_div(nitems, kTrWarps), (i64)num_sms() * per_sm);
        const char* ek = getenv("CDB_TRANSLATE_KEEP");
        translate_kernel<<<grid, kTrWarps * 32, 0, st>>>(cdocs.p, ccnt.p, alloc_off.p, row_off.p, seg.p, ids_tab, pairs.p, npat, nranges,
                                                         counters.p + 4, ek ? atoi(ek) : 0);
        CDB_LAUNCH_CHECK();
    }
    CDB_CUDA(cudaEventRecord(ev[5], st));
    // the listed rows: streamed from the document listing into their CSR rows
    if (nlisted) {
        // a lazy caller reads the rows without repeats from the listing itself
        launch_listing_emit(*lst, pre.p, left.p, right.p, row_off.p, npat, pairs.p, lazy ? 1 : 0, nullptr, st, 0);
        if (lazy) {
            lazy->lst = lst;
            lazy->stream = st;
            lazy->active = true;
        }
    }
    CDB_CUDA(cudaEventRecord(ev[7], st));
    if (hook_late) {
        ThreadCtx& tc = thread_ctx(ix.device);
        if (!tc.copy_stream) CDB_CUDA(cudaStreamCreateWithFlags(&tc.copy_stream, cudaStreamNonBlocking));
        CDB_CUDA(cudaStreamWaitEvent(tc.copy_stream, ev[8], 0));
        rows_ready(rows_ready_user, stats.p, npat, (void*)tc.copy_stream);  // the statistics are ready on THIS stream
    }
    for (LargeChunk& lc : lchunks) {
        if (lc.nu == 0) continue;
        const int grid = (int)std::min<i64>(ceil_div((i64)lc.nu, 256), num_sms() * 16);
        large_emit_kernel<<<grid, 256, 0, st>>>(lc.ukey.p, lc.ustart.p, lc.nu, lc.ltotal, large_list.p + lc.j0, lc.entry_first.p,
                                                row_off.p, ids_tab, pairs.p);
        CDB_LAUNCH_CHECK();
    }
    u64 total_pairs = 0;
    CDB_CUDA(cudaMemcpyAsync(&total_pairs, row_off.p + npat, 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaEventRecord(ev[6], st));
    CDB_CUDA(cudaStreamSynchronize(st));
    {
        LocateStats& ls = g_locate_stats;
        cudaEventElapsedTime(&ls.search_ms, ev[0], ev[1]);
        cudaEventElapsedTime(&ls.large_ms, ev[2], ev[3]);
        cudaEventElapsedTime(&ls.gather_ms, ev[3], ev[4]);  // gather_kernel (phase A)
        cudaEventElapsedTime(&ls.translate_ms, ev[4], ev[5]);   // translate_kernel (phase B)
        cudaEventElapsedTime(&ls.listing_ms, ev[5], ev[7]);  // listing_emit_kernel (rows answered from the document listing)
        cudaEventElapsedTime(&ls.tail_ms, ev[7], ev[6]);   // large-path emit + final read-back
        cudaEventElapsedTime(&ls.total_ms, ev[0], ev[6]);
        ls.npat = npat;
        ls.total_pairs = (long long)total_pairs;
        ls.total_occ = (long long)total_occ;
        ls.nlarge = (long long)nl;
        ls.nlisted = (long long)nlisted;
        ls.listed_pairs = (long long)listed_pairs;
    }
    out->npat = npat;
    out->total_pairs = (i64)total_pairs;
    out->total_occurrences = (i64)total_occ;
    out->row_off = reinterpret_cast<i64*>(row_off.detach());
    out->pairs = pairs.detach();
    out->left = left.detach();
    out->right = right.detach();
    out->stats32 = stats.detach();
    out->row_flags = rowflag.detach();
    out->_owner = (void*)st;
    if (lazy && lazy->active) lazy->pre = pre.detach();
}


// ---- small batches (CDB_SMALL_BATCH = largest batch that takes this path, default 256; 0 = general path only) ------
// One keyword through the general path costs ~130 us of launches, synchronisations and stream-ordered allocations
// (a dozen kernels, three cudaStreamSynchronize, four copies), whatever the work.  A batch of up to kSmallMaxPat keywords
// takes ONE upload, TWO launches and ONE synchronisation instead: the packed request (zeroed counters + offsets +
// keyword bytes) goes up in a single copy, search_kernel finds the intervals, and small_gather_kernel does the rest per
// warp — sort, run-length, ids[] lookup (random reads do not matter at this size) — writing (id, count) pairs straight
// into mapped pinned memory of the calling thread at the scan of the occurrence counts, which every warp computes for
// itself.  Batches with a long interval (> kWarpCap) or more than kSmallCapPairs occurrences report that in the header
// and the caller falls back to the general path.  All buffers are cached per host thread and device.
constexpr int kSmallMaxPat = 256;
constexpr u64 kSmallCapPairs = 1u << 16;  // 1 MB of pairs
constexpr size_t kSmallPatBytes = 16384;
constexpr size_t kSmallHdrWords = 4;      // out: [0] long intervals [1] occurrences [2] empty-keyword flag [3] unused

template <typename SAT>
__global__ void __launch_bounds__(kTileWarps * 32) small_gather_kernel(const SAT* __restrict__ sa, u64 mask, u32 bucket_mul,
                                                                         const i64* __restrict__ left,
                                                                         const i64* __restrict__ right, int npat,
                                                                         const u64* __restrict__ wocc,
                                                                         const unsigned long long* __restrict__ counters,
                                                                         const i64* __restrict__ ids, u64* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kTileWarps + warp;
    u64* rowlen = out + kSmallHdrWords;          // exact row length of every pattern
    u64* rowocc = rowlen + kSmallMaxPat;         // its occurrence count (rows sit at the scan of these)
    longlong2* pairs = reinterpret_cast<longlong2*>(rowocc + kSmallMaxPat);
    const u64 nlarge = counters[0], total = counters[1], err = counters[2] & 0xffffffffull;
    if (q == 0 && lane == 0) {
        out[0] = nlarge;
        out[1] = total;
        out[2] = err;
    }
    if (q >= npat || nlarge != 0 || err != 0 || total > kSmallCapPairs) return;  // the host falls back (or throws)
    u64 base = 0;
    for (int i = lane; i < q; i += 32) base += wocc[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    u32* s_doc = reinterpret_cast<u32*>(smem_raw + (size_t)warp * warp_smem_bytes<32>());
    u32* s_pos = s_doc + 32 * 32 + 32;
    const i64 l = left[q];
    const int occ = (int)(right[q] - l);  // <= kWarpCap: no long interval in this batch
    int nheads = 0;
    bool all_distinct = false;
    if (occ > 0) {
        if (occ <= 32) nheads = load_sort_rle<SAT, 1>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct);
        else if (occ <= 64) nheads = load_sort_rle<SAT, 2>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct);
        else if (occ <= 128) nheads = load_sort_rle<SAT, 4>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct);
        else if (occ <= 256) nheads = load_sort_rle<SAT, 8>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct);
        else if (occ <= 512) nheads = load_sort_rle<SAT, 16>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct);
        else nheads = load_sort_rle<SAT, 32>(sa, l, occ, mask, bucket_mul, s_doc, s_pos, lane, all_distinct);
    }
    if (lane == 0) {
        rowlen[q] = (u64)nheads;
        rowocc[q] = (u64)occ;
    }
    for (int r = lane; r < nheads; r += 32) {
        const u32 doc = s_doc[pad_idx(r)];
        const i64 cnt = all_distinct ? 1 : (i64)(s_pos[pad_idx(r + 1)] - s_pos[pad_idx(r)]);
        pairs[base + r] = make_longlong2(__ldg(ids + doc), cnt);
    }
}

struct SmallCtx {
    u8* h_in = nullptr;   // pinned staging of the packed request
    u8* d_in = nullptr;   // [counters: 8 words, zero][pat_off: npat + 1][keyword bytes]
    u8* d_tmp = nullptr;  // left, right, wocc, large_list
    u64* h_out = nullptr; // mapped pinned: header, rowlen, rowocc, pairs
    u64* d_out = nullptr;
};
static constexpr size_t kSmallInBytes = 64 + 8 * (kSmallMaxPat + 1) + kSmallPatBytes;
static constexpr size_t kSmallOutBytes = (kSmallHdrWords + 2 * kSmallMaxPat) * 8 + kSmallCapPairs * 16;

// The buffer sets are pooled per device, not kept per host thread: behind cdb_query any of the server's worker threads may
// lead a batch, and a set per thread cost every new leader ~1 ms of page-locked allocations.  A set is checked out for
// one call and goes back when the caller has copied its rows (SmallResult's destructor).
static std::mutex g_small_mu;
static std::map<int, std::vector<SmallCtx*>> g_small_free;

static SmallCtx* small_ctx_checkout(int device) {
    {
        std::lock_guard<std::mutex> lk(g_small_mu);
        auto& v = g_small_free[device];
        if (!v.empty()) {
            SmallCtx* c = v.back();
            v.pop_back();
            return c;
        }
    }
    SmallCtx* c = new SmallCtx();
    try {
        CDB_CUDA(cudaHostAlloc((void**)&c->h_in, kSmallInBytes, cudaHostAllocDefault));
        CDB_CUDA(cudaMalloc((void**)&c->d_in, kSmallInBytes));
        CDB_CUDA(cudaMalloc((void**)&c->d_tmp, (size_t)kSmallMaxPat * 32 + 64));
        CDB_CUDA(cudaHostAlloc((void**)&c->h_out, kSmallOutBytes, cudaHostAllocMapped));
        CDB_CUDA(cudaHostGetDevicePointer((void**)&c->d_out, c->h_out, 0));
    } catch (...) {
        if (c->h_in) cudaFreeHost(c->h_in);
        if (c->d_in) cudaFree(c->d_in);
        if (c->d_tmp) cudaFree(c->d_tmp);
        if (c->h_out) cudaFreeHost(c->h_out);
        delete c;
        throw;
    }
    return c;
}

static void small_ctx_return(int device, SmallCtx* c) {
    std::lock_guard<std::mutex> lk(g_small_mu);
    g_small_free[device].push_back(c);
}

SmallResult::~SmallResult() {
    if (_ctx) small_ctx_return(_device, static_cast<SmallCtx*>(_ctx));
}

template <typename SAT>
static bool locate_small_typed(const Index& ix, const u8* pat, const i64* pat_off, int npat, cudaStream_t st, SmallResult* res) {
    struct Lease {  // back to the pool on every way out, unless the result takes it over
        int device;
        SmallCtx* c;
        ~Lease() {
            if (c) small_ctx_return(device, c);
        }
    } lease{ix.device, small_ctx_checkout(ix.device)};
    SmallCtx& sc = *lease.c;
    const i64 p0 = pat_off[0], pbytes = pat_off[npat] - p0;
    std::memset(sc.h_in, 0, 64);
    i64* h_off = reinterpret_cast<i64*>(sc.h_in + 64);
    for (int q = 0; q <= npat; ++q) h_off[q] = pat_off[q] - p0;
    u8* h_bytes = sc.h_in + 64 + 8 * (size_t)(npat + 1);
    std::memcpy(h_bytes, pat + p0, (size_t)pbytes);
    const size_t in_bytes = 64 + 8 * (size_t)(npat + 1) + (size_t)pbytes;
    CDB_CUDA(cudaMemcpyAsync(sc.d_in, sc.h_in, in_bytes, cudaMemcpyHostToDevice, st));
    unsigned long long* counters = reinterpret_cast<unsigned long long*>(sc.d_in);
    const i64* d_off = reinterpret_cast<const i64*>(sc.d_in + 64);
    const u8* d_pat = sc.d_in + 64 + 8 * (size_t)(npat + 1);
    i64* left = reinterpret_cast<i64*>(sc.d_tmp);
    i64* right = left + kSmallMaxPat;
    u64* wocc = reinterpret_cast<u64*>(right + kSmallMaxPat);
    u32* large_list = reinterpret_cast<u32*>(wocc + kSmallMaxPat);
    SearchCtx c = make_ctx(ix);
    search_kernel<SAT><<<1, 256, 0, st>>>(c, ix.symtab, d_pat, d_off, (i64)npat, left, right, reinterpret_cast<int*>(counters + 2),
                                         large_list, counters, wocc, nullptr, nullptr);
    CDB_LAUNCH_CHECK();
    const u32 bucket_mul = ix.nd > 0 ? (u32)std::min<u64>(0xffffffffull, (1024ull << 32) / (u64)ix.nd) : 0u;
    const size_t smem = (size_t)kTileWarps * warp_smem_bytes<32>();
    CDB_CUDA(cudaFuncSetAttribute(small_gather_kernel<SAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    small_gather_kernel<SAT><<<(unsigned)ceil_div((i64)npat, kTileWarps), kTileWarps * 32, smem, st>>>(
        reinterpret_cast<const SAT*>(ix.d_sa), ix.mask, bucket_mul, left, right, npat, wocc, counters, ix.d_ids, sc.d_out);
    CDB_LAUNCH_CHECK();
    CDB_CUDA(cudaStreamSynchronize(st));
    const u64* h = sc.h_out;
    if (h[2]) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");
    if (h[0] != 0 || h[1] > kSmallCapPairs) return false;  // long interval or too many occurrences: general path
    res->total_occ = (i64)h[1];
    res->rowlen = h + kSmallHdrWords;
    res->rowocc = res->rowlen + kSmallMaxPat;
    res->pairs = reinterpret_cast<const i64*>(res->rowocc + kSmallMaxPat);
    res->_ctx = lease.c;  // the rows stay valid until the result goes
    res->_device = ix.device;
    lease.c = nullptr;
    return true;
}

int small_batch_limit() {
    const char* e = getenv("CDB_SMALL_BATCH");  // read per call: the tests switch it (0 = general path only)
    if (!e) return kSmallMaxPat;
    const int v = atoi(e);
    return v < 0 ? 0 : (v > kSmallMaxPat ? kSmallMaxPat : v);
}

bool locate_small(const Index& ix, const u8* pat, const i64* pat_off, i64 npat, cudaStream_t st, SmallResult* res) {
    if (npat <= 0 || npat > kSmallMaxPat || pat_off[npat] - pat_off[0] > (i64)kSmallPatBytes) return false;
    return ix.width == 4 ? locate_small_typed<u32>(ix, pat, pat_off, (int)npat, st, res)
                         : locate_small_typed<u64>(ix, pat, pat_off, (int)npat, st, res);
}

void locate_device(const Index& ix, const u8* d_pat, const i64* d_pat_off, i64 npat, cudaStream_t st,
                   cdb_device_result* out, bool id_order, cdb_rows_ready_fn rows_ready, void* rows_ready_user, LazyListed* lazy) {
    if (ix.width == 4)
        locate_typed<u32>(ix, d_pat, d_pat_off, npat, st, out, id_order, rows_ready, rows_ready_user, lazy);
    else
        locate_typed<u64>(ix, d_pat, d_pat_off, npat, st, out, id_order, rows_ready, rows_ready_user, lazy);
}

}  // namespace cdb
