// Independent device-side check of a finished suffix array (cdb_verify_sa).  It shares no code with the build: it only
// reads text, doc_off and the packed array and applies the reference's comparator to every adjacent pair.
//
//   order        suffix(sa[i]) <= suffix(sa[i+1]) under the comparator of the reference's leaf sort and binary search
//                (std::string_view <: unsigned bytes, a proper prefix first; src/index.cpp:92-93, 268).  In the note-N1
//                layout (corpus mixes bytes < 0x80 and >= 0x80, n > chuck_size, compat_signed) a pair whose first
//                differing bytes lie on different sides of 0x80 is ordered by SIGNED byte iff the group of suffixes
//                sharing the common prefix is larger than chuck_size = max(4096, n/256) — the radix levels of
//                src/index.cpp:97-125 with character() of src/index.h:66-73 — and by unsigned byte otherwise.  The
//                group is the contiguous run around i (galloping + binary search on "shares the d-byte prefix").
//   permutation  every element names a valid (doc, offset) and no text position is named twice (bit map of n bits);
//                together with the length n that makes the array a permutation of all suffixes.
//   ties         adjacent byte-identical suffixes are counted, and how many of them are not in ascending packed order
//                (the canonical order this build promises, SURVEY.md note N2; the reference's is unspecified).
// n adjacent compares with ~3 random sector reads each: ~1 s at n = 10^10.
#include <algorithm>

#include "index.cuh"
#include "verify.cuh"

namespace cdb {

struct VerifyCtx {
    const void* sa;
    i64 n, nd;
    int bits1;
    u64 mask;
    const i64* doc_off;
    const u8* text;
};

struct SufRef {
    i64 start;  // text position of the suffix
    i64 len;    // bytes to the end of its document (<= 0: invalid element)
};

template <typename SAT>
__device__ __forceinline__ SufRef suffix_at(const VerifyCtx& c, i64 i, u64* packed) {
    const u64 e = (u64) reinterpret_cast<const SAT*>(c.sa)[i];
    *packed = e;
    const u64 doc = e & c.mask;
    const u64 off = e >> c.bits1;
    SufRef r{0, 0};
    if (doc >= (u64)c.nd) return r;
    const i64 ds = __ldg(c.doc_off + doc), de = __ldg(c.doc_off + doc + 1);
    if (off >= (u64)(de - ds)) return r;
    r.start = ds + (i64)off;
    r.len = de - r.start;
    return r;
}

// first position where the two suffixes differ (or one ends): returns the depth d; ca/cb = byte at d, -1 = ended
__device__ __forceinline__ i64 first_difference(const u8* __restrict__ text, SufRef a, SufRef b, int* ca, int* cb) {
    i64 o = 0;
    for (;;) {
        const i64 ra = a.len - o, rb = b.len - o;
        if (ra <= 0 || rb <= 0) {
            *ca = ra <= 0 ? -1 : (int)text[a.start + o];
            *cb = rb <= 0 ? -1 : (int)text[b.start + o];
            return o;
        }
        i64 k = ra < rb ? ra : rb;
        if (k > 8) k = 8;
        const int sh = 64 - 8 * (int)k;
        const u64 ta = load_be64(text, a.start + o) >> sh, tb = load_be64(text, b.start + o) >> sh;
        if (ta != tb) {
            const int lead = __clzll((long long)((ta ^ tb) << sh)) >> 3;  // equal leading bytes of the window
            *ca = (int)((ta >> (8 * ((int)k - 1 - lead))) & 255);
            *cb = (int)((tb >> (8 * ((int)k - 1 - lead))) & 255);
            return o + lead;
        }
        o += k;
    }
}

// counters: [0] inversions [1] invalid [2] duplicates [3] ties [4] ties out of packed order [5] pairs queued for the
// signed-rule check [6] queue overflow (pairs left unchecked)
template <typename SAT>
__global__ void __launch_bounds__(256) verify_pairs_kernel(VerifyCtx c, bool n1, u32* __restrict__ bitmap,
                                                           unsigned long long* __restrict__ cnt, i64* __restrict__ queue,
                                                           u64 queue_cap, i64 i_begin, i64 i_end,
                                                           unsigned long long* __restrict__ qcount) {
    // one launch covers the ranks [i_begin, i_end); a warp whose last lane is i_end - 1 still compares it with i_end
    const i64 i = i_begin + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    // (slices start at multiples of 32 and lanes beyond i_end stay for the shuffles below)
    const int lane = threadIdx.x & 31;
    u64 pa = 0;
    SufRef a{0, 0};
    if (i < c.n && i < i_end) {
        a = suffix_at<SAT>(c, i, &pa);
        if (a.len <= 0) {
            atomicAdd(cnt + 1, 1ull);
        } else {
            const u32 bit = 1u << (a.start & 31);
            if (atomicOr(bitmap + (a.start >> 5), bit) & bit) atomicAdd(cnt + 2, 1ull);
        }
    }
    // the right neighbour comes from the next lane; the last lane of a warp loads it itself
    SufRef b;
    b.start = __shfl_down_sync(0xffffffffu, a.start, 1);
    b.len = __shfl_down_sync(0xffffffffu, a.len, 1);
    u64 pb = __shfl_down_sync(0xffffffffu, pa, 1);
    if (i + 1 >= c.n || i >= i_end) return;
    if (lane == 31 || i + 1 >= i_end) b = suffix_at<SAT>(c, i + 1, &pb);
    if (a.len <= 0 || b.len <= 0) return;  // counted as invalid above (or by the neighbour's thread)
    int ca, cb;
    const i64 d = first_difference(c.text, a, b, &ca, &cb);
    if (ca < 0 && cb < 0) {
        atomicAdd(cnt + 3, 1ull);
        if (pa > pb) atomicAdd(cnt + 4, 1ull);
        return;
    }
    if (ca < 0) return;  // a is a proper prefix of b: in order under both rules (end-of-document first)
    if (cb < 0) {
        atomicAdd(cnt + 0, 1ull);
        return;
    }
    if (!n1 || ((ca < 0x80) == (cb < 0x80))) {
        if (ca > cb) atomicAdd(cnt + 0, 1ull);
        return;
    }
    atomicAdd(cnt + 5, 1ull);
    const u64 slot = atomicAdd(qcount, 1ull);
    if (slot < queue_cap) {
        queue[2 * slot] = i;
        queue[2 * slot + 1] = d;
    } else {
        atomicAdd(cnt + 6, 1ull);
    }
}

// does the suffix at rank j share the first d bytes of suffix a (and is it at least d long)?
template <typename SAT>
__device__ __forceinline__ bool shares_prefix(const VerifyCtx& c, SufRef a, i64 j, i64 d) {
    u64 p;
    const SufRef s = suffix_at<SAT>(c, j, &p);
    if (s.len < d) return false;
    for (i64 o = 0; o < d; o += 8) {
        const i64 k = d - o < 8 ? d - o : 8;
        const int sh = 64 - 8 * (int)k;
        if ((load_be64(c.text, a.start + o) >> sh) != (load_be64(c.text, s.start + o) >> sh)) return false;
    }
    return true;
}

template <typename SAT>
__global__ void __launch_bounds__(128) verify_signed_kernel(VerifyCtx c, i64 chuck, const i64* __restrict__ queue, u64 nq,
                                                            unsigned long long* __restrict__ cnt) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const i64 i = queue[2 * t], d = queue[2 * t + 1];
    u64 pa, pb;
    const SufRef a = suffix_at<SAT>(c, i, &pa), b = suffix_at<SAT>(c, i + 1, &pb);
    const int ca = (int)c.text[a.start + d], cb = (int)c.text[b.start + d];
    // extent of the group around i: gallop until the prefix is no longer shared (or chuck + 1 members are known)
    i64 lo = i, hi = i + 2;  // [lo, hi) shares the prefix so far (i and i + 1 do by construction)
    {
        i64 step = 1, bad = -1;  // largest index known NOT to share, to the left
        while (lo > 0 && hi - lo <= chuck) {
            const i64 j = lo - step < 0 ? 0 : lo - step;
            if (shares_prefix<SAT>(c, a, j, d)) {
                lo = j;
                step <<= 1;
            } else {
                bad = j;
                break;
            }
        }
        while (bad >= 0 && bad + 1 < lo) {  // binary search in (bad, lo)
            const i64 m = bad + (lo - bad) / 2;
            if (shares_prefix<SAT>(c, a, m, d))
                lo = m;
            else
                bad = m;
        }
    }
    {
        i64 step = 1, bad = -1;  // smallest index known NOT to share, to the right
        while (hi < c.n && hi - lo <= chuck) {
            const i64 j = hi - 1 + step >= c.n ? c.n - 1 : hi - 1 + step;
            if (shares_prefix<SAT>(c, a, j, d)) {
                hi = j + 1;
                step <<= 1;
            } else {
                bad = j;
                break;
            }
        }
        while (bad >= 0 && hi < bad) {  // binary search in [hi, bad)
            const i64 m = hi + (bad - hi) / 2;
            if (shares_prefix<SAT>(c, a, m, d))
                hi = m + 1;
            else
                bad = m;
        }
    }
    const bool signed_rule = hi - lo > chuck;
    // signed: bytes >= 0x80 are negative and come first; unsigned: plain byte order
    const bool ok = signed_rule ? (ca >= 0x80 && cb < 0x80) : (ca < cb);
    if (!ok) atomicAdd(cnt + 0, 1ull);
    if (signed_rule) atomicAdd(cnt + 7, 1ull);
}

// Element-wise comparison with another packed suffix array of the same corpus (e.g. the compiled reference's): where
// the elements differ the two suffixes must be byte-identical (a tie, note N2), anything else is a real difference.
// counters: [0] identical elements [1] different elements, identical suffixes [2] different suffixes
template <typename SAT>
__global__ void __launch_bounds__(256) compare_sa_kernel(VerifyCtx c, const SAT* __restrict__ other, i64 base, i64 count,
                                                         unsigned long long* __restrict__ cnt) {
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    int kind = -1;
    if (j < count) {
        u64 pa;
        const SufRef a = suffix_at<SAT>(c, base + j, &pa);
        const u64 pb = (u64)other[j];
        if (pa == pb) {
            kind = 0;
        } else {
            VerifyCtx c2 = c;
            c2.sa = other - base;  // suffix_at indexes with the global rank
            u64 tmp;
            const SufRef b = suffix_at<SAT>(c2, base + j, &tmp);
            int ca, cb;
            kind = 2;
            if (a.len > 0 && b.len > 0) {
                first_difference(c.text, a, b, &ca, &cb);
                if (ca < 0 && cb < 0) kind = 1;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, kind == k);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(cnt + k, (unsigned long long)__popc(m));
    }
}

void compare_index_sa(const Index& ix, const void* host_sa, cudaStream_t st, i64 out[3]) {
    for (int k = 0; k < 3; ++k) out[k] = 0;
    if (ix.n == 0) return;
    VerifyCtx c{ix.d_sa, ix.n, ix.nd, ix.bits1, ix.mask, ix.d_off, ix.d_text};
    DevBuf<unsigned long long> cnt(3, st);
    CDB_CUDA(cudaMemsetAsync(cnt.p, 0, 24, st));
    const i64 chunk = (i64)1 << 26;  // elements per upload
    DevBuf<u8> buf((size_t)std::min<i64>(chunk, ix.n) * ix.width, st);
    for (i64 base = 0; base < ix.n; base += chunk) {
        const i64 count = std::min<i64>(chunk, ix.n - base);
        CDB_CUDA(cudaMemcpyAsync(buf.p, (const u8*)host_sa + base * ix.width, (size_t)count * ix.width, cudaMemcpyHostToDevice, st));
        const unsigned grid = (unsigned)ceil_div(count, 256);
        if (ix.width == 4)
            compare_sa_kernel<u32><<<grid, 256, 0, st>>>(c, reinterpret_cast<const u32*>(buf.p), base, count, cnt.p);
        else
            compare_sa_kernel<u64><<<grid, 256, 0, st>>>(c, reinterpret_cast<const u64*>(buf.p), base, count, cnt.p);
        CDB_LAUNCH_CHECK();
        CDB_CUDA(cudaStreamSynchronize(st));  // the staging buffer is re-used
    }
    unsigned long long h[3];
    CDB_CUDA(cudaMemcpyAsync(h, cnt.p, 24, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) out[k] = (i64)h[k];
}

void verify_index(const Index& ix, cudaStream_t st, i64 out[8]) {
    for (int k = 0; k < 8; ++k) out[k] = 0;
    if (ix.n == 0) return;
    const bool n1 = ix.mixed && ix.opt.compat_signed && ix.n > ix.chuck_size;
    VerifyCtx c{ix.d_sa, ix.n, ix.nd, ix.bits1, ix.mask, ix.d_off, ix.d_text};
    const size_t words = (size_t)((ix.n + 31) / 32);
    DevBuf<u32> bitmap(words, st);
    DevBuf<unsigned long long> cnt(8, st);
    // Note-N1 layout: the array is walked in slices of at most 2^26 ranks, so the queue of pairs waiting for the
    // signed-rule check (at most one per rank) can never overflow; otherwise one launch covers everything.
    const i64 slice = n1 ? std::min<i64>((ix.n + 31) & ~(i64)31, (i64)1 << 26) : ((ix.n + 31) & ~(i64)31);
    const u64 qcap = n1 ? (u64)slice : 1;
    DevBuf<i64> queue((size_t)qcap * 2, st);
    DevBuf<unsigned long long> qcount(1, st);
    CDB_CUDA(cudaMemsetAsync(bitmap.p, 0, words * 4, st));
    CDB_CUDA(cudaMemsetAsync(cnt.p, 0, 64, st));
    unsigned long long h[8];
    for (i64 i0 = 0; i0 < ix.n; i0 += slice) {
        const i64 i1 = std::min<i64>(ix.n, i0 + slice);
        CDB_CUDA(cudaMemsetAsync(qcount.p, 0, 8, st));
        const unsigned grid = (unsigned)ceil_div(i1 - i0, 256);
        if (ix.width == 4)
            verify_pairs_kernel<u32><<<grid, 256, 0, st>>>(c, n1, bitmap.p, cnt.p, queue.p, qcap, i0, i1, qcount.p);
        else
            verify_pairs_kernel<u64><<<grid, 256, 0, st>>>(c, n1, bitmap.p, cnt.p, queue.p, qcap, i0, i1, qcount.p);
        CDB_LAUNCH_CHECK();
        unsigned long long hq = 0;
        CDB_CUDA(cudaMemcpyAsync(&hq, qcount.p, 8, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        const u64 nq = std::min<u64>(hq, qcap);
        if (nq) {
            const unsigned g2 = (unsigned)ceil_div((i64)nq, 128);
            if (ix.width == 4)
                verify_signed_kernel<u32><<<g2, 128, 0, st>>>(c, ix.chuck_size, queue.p, nq, cnt.p);
            else
                verify_signed_kernel<u64><<<g2, 128, 0, st>>>(c, ix.chuck_size, queue.p, nq, cnt.p);
            CDB_LAUNCH_CHECK();
        }
    }
    CDB_CUDA(cudaMemcpyAsync(h, cnt.p, 64, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 8; ++k) out[k] = (i64)h[k];
}

}  // namespace cdb
