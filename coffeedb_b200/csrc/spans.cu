// K8, batched highlight spans — replaces the occurrence enumeration of ac_automaton::render (src/database.cpp:58-77)
// for MANY (request, document) texts in one launch sequence.
//
// A request is a set of keywords (the keyword list of one key, src/interface.cpp:211-227); a "text" is one document
// that request is to be highlighted in (an object that survived the filter and the span slice, src/database.cpp:401-432).
// For every text the kernels scan the document's bytes, find every occurrence [p, p + len - 1] of every keyword of
// its request and merge them the way render() does: ascending by start, an occurrence extends the open span when it
// starts at or before the span's end (overlap), and opens a new span otherwise — merely touching occurrences stay
// separate (src/database.cpp:66-76; test/test-highlight.py:56-57 relies on it).  This is exactly the reference's
// semantics on every layout (its highlighter scans the text, it never consults the suffix array), also in the note-N1
// layout where query() itself misses occurrences.
//
//   spans_thread_kernel   one THREAD per text, for documents of up to kThreadMaxLen bytes: a sequential sweep with the
//                         keyword compared against an 8-byte big-endian window; 32 texts per warp, every thread reads
//                         its own document (whole sectors, L1 serves the window reloads)
//   spans_warp_kernel     one WARP per long text: 32 positions per step, a warp prefix-max of the occurrence ends
//                         gives every position the furthest end opened before it; persistent warps walk a list of the
//                         long texts
// Both run twice: COUNT (spans per text -> exclusive scan -> CSR offsets) and WRITE.  One host synchronisation per
// batch (the total number of spans sizes the result).  HBM traffic: the documents' bytes once per pass (100-byte
// documents: 4 sectors each), 16 bytes per span written.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "index.cuh"
#include "locate.cuh"
#include "primitives.cuh"

namespace cdb {

constexpr i64 kThreadMaxLen = 1024;  // longer documents go to the warp kernel

struct SpanBatch {
    const u8* text;
    const i64* doc_off;
    i64 nd;
    const u8* kw;           // keyword bytes (padded by 8)
    const i64* kw_off;      // [nkw + 1]
    const i64* req_kw_off;  // [nreq + 1] -> keyword index range of a request
    i64 nkw, nreq;
    const i64* text_req;    // [ntext] request of a text
    const i64* text_doc;    // [ntext] document (doc index) of a text
    i64 ntext;
    const u64* kw_code;     // [nkw] first min(len, 8) bytes, big-endian, right-aligned
};

// first min(m, 8) keyword bytes as a right-aligned big-endian integer; flags empty keywords (src/index.cpp:239-241 for
// the query, and an empty needle would match everywhere)
__global__ void span_kwinfo_kernel(const u8* __restrict__ kw, const i64* __restrict__ kw_off, i64 nkw, u64* __restrict__ code,
                                   int* __restrict__ err) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkw) return;
    const i64 s = kw_off[k], m = kw_off[k + 1] - s;
    if (m <= 0) {
        atomicOr(err, 1);
        code[k] = 0;
        return;
    }
    const int m8 = m < 8 ? (int)m : 8;
    u64 c = 0;
    for (int i = 0; i < m8; ++i) c = (c << 8) | kw[s + i];
    code[k] = c;
}

// furthest end of a keyword occurrence that starts at position p of the document [ds, ds + len), or -1
__device__ __forceinline__ i64 best_end_at(const SpanBatch& b, i64 ds, i64 len, i64 p, i64 k0, i64 k1, u64 w) {
    i64 best = -1;
    for (i64 k = k0; k < k1; ++k) {
        const i64 ks = __ldg(b.kw_off + k), m = __ldg(b.kw_off + k + 1) - ks;
        if (p + m > len) continue;
        const int m8 = m < 8 ? (int)m : 8;
        if ((w >> (64 - 8 * m8)) != __ldg(b.kw_code + k)) continue;
        bool eq = true;
        for (i64 i = 8; i < m && eq; ++i) eq = b.text[ds + p + i] == b.kw[ks + i];
        if (eq && p + m - 1 > best) best = p + m - 1;
    }
    return best;
}

template <bool WRITE>
__global__ void __launch_bounds__(128) spans_thread_kernel(SpanBatch b, u64* __restrict__ cnt, const u64* __restrict__ soff,
                                                           i64* __restrict__ spans, u32* __restrict__ long_list,
                                                           unsigned long long* __restrict__ counters, int* __restrict__ err) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.ntext) return;
    const i64 d = b.text_doc[t], r = b.text_req[t];
    if (d < 0 || d >= b.nd || r < 0 || r >= b.nreq) {
        atomicOr(err, 2);
        if (!WRITE) cnt[t] = 0;
        return;
    }
    const i64 ds = __ldg(b.doc_off + d), len = __ldg(b.doc_off + d + 1) - ds;
    if (len > kThreadMaxLen) {
        if (!WRITE) {
            cnt[t] = 0;  // filled in by the warp kernel
            long_list[atomicAdd(counters, 1ull)] = (u32)t;
        }
        return;
    }
    const i64 k0 = __ldg(b.req_kw_off + r), k1 = __ldg(b.req_kw_off + r + 1);
    u64 n = 0, out = WRITE ? soff[t] : 0;
    bool open = false;
    i64 cb = 0, ce = 0;
    // Rolling window: the document is read ONCE, one aligned 8-byte word per 8 positions (the text is padded, so the word
    // after the document's last byte exists); the 8 bytes from position p on are cut out of two neighbouring words.
    // Reloading the window at every position made every warp-wide load touch 32 different lines — 200 of them per
    // 100-byte document — and the kernel ran at the L1 tag rate: 28 ms per 3.2e7 texts.
    const u64* words = reinterpret_cast<const u64*>(b.text + (ds & ~(i64)7));
    u64 cur = __ldg(words), nxt = __ldg(words + 1);
    i64 wi = 1;
    u32 sh = (u32)(ds & 7) * 8;
    // one keyword per request (the usual case): its length and first bytes stay in registers
    const bool single = k1 - k0 == 1;
    i64 m1 = 0, ks1 = 0;
    u64 code1 = 0;
    int sh1 = 0;
    if (single) {
        ks1 = __ldg(b.kw_off + k0);
        m1 = __ldg(b.kw_off + k0 + 1) - ks1;
        code1 = __ldg(b.kw_code + k0);
        sh1 = 64 - 8 * (m1 < 8 ? (int)m1 : 8);
    }
    for (i64 p = 0; p < len; ++p) {
        const u64 v = sh ? ((cur >> sh) | (nxt << (64 - sh))) : cur;  // little-endian: byte p is the low byte
        const u64 w = ((u64)__byte_perm((u32)v, 0, 0x0123) << 32) | __byte_perm((u32)(v >> 32), 0, 0x0123);
        sh += 8;
        if (sh == 64) {
            sh = 0;
            cur = nxt;
            nxt = __ldg(words + ++wi);
        }
        i64 e = -1;
        if (single) {
            if (p + m1 <= len && (w >> sh1) == code1) {
                bool eq = true;
                for (i64 i = 8; i < m1 && eq; ++i) eq = b.text[ds + p + i] == b.kw[ks1 + i];
                if (eq) e = p + m1 - 1;
            }
        } else {
            e = best_end_at(b, ds, len, p, k0, k1, w);
        }
        if (e < 0) continue;
        if (open && p <= ce) {
            ce = e > ce ? e : ce;
        } else {
            if (open) {
                if (WRITE) {
                    spans[2 * out] = cb;
                    spans[2 * out + 1] = ce;
                    ++out;
                }
                ++n;
            }
            open = true;
            cb = p;
            ce = e;
        }
    }
    if (open) {
        if (WRITE) {
            spans[2 * out] = cb;
            spans[2 * out + 1] = ce;
        }
        ++n;
    }
    if (!WRITE) cnt[t] = n;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) spans_warp_kernel(SpanBatch b, u64* __restrict__ cnt, const u64* __restrict__ soff,
                                                         i64* __restrict__ spans, const u32* __restrict__ long_list,
                                                         const unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const u64 nlong = counters[0];
    const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 j = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < nlong; j += nwarps) {
        const i64 t = long_list[j];
        const i64 d = b.text_doc[t], r = b.text_req[t];
        const i64 ds = __ldg(b.doc_off + d), len = __ldg(b.doc_off + d + 1) - ds;
        const i64 k0 = __ldg(b.req_kw_off + r), k1 = __ldg(b.req_kw_off + r + 1);
        const u64 out = WRITE ? soff[t] : 0;
        i64 carry = -1;  // furthest occurrence end among the positions before this step
        u64 n = 0;
        for (i64 base = 0; base < len; base += 32) {
            const i64 p = base + lane;
            i64 e = -1;
            if (p < len) e = best_end_at(b, ds, len, p, k0, k1, load_be64(b.text, ds + p));
            i64 incl = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const i64 v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o && v > incl) incl = v;
            }
            i64 prev = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) prev = -1;
            const i64 eprev = prev > carry ? prev : carry;  // furthest end opened before p
            const bool start = e >= 0 && p > eprev;
            const u32 bal = __ballot_sync(0xffffffffu, start);
            if (WRITE && start) {
                const u64 idx = n + __popc(bal & lanemask_lt());
                spans[2 * (out + idx)] = p;
                if (idx > 0) spans[2 * (out + idx - 1) + 1] = eprev;  // the previous span ends at the furthest end so far
            }
            n += __popc(bal);
            const i64 last = __shfl_sync(0xffffffffu, incl, 31);
            if (last > carry) carry = last;
        }
        if (WRITE && n > 0 && lane == 0) spans[2 * (out + n - 1) + 1] = carry;
        if (!WRITE && lane == 0) cnt[t] = n;
    }
}

// Everything in device memory.  span_off_out receives ntext + 1 CSR offsets, spans_out 2 * total int64 (begin, end).
void locate_spans_batch_device(const Index& ix, const u8* d_kw, const i64* d_kw_off, i64 nkw, const i64* d_req_kw_off, i64 nreq,
                               const i64* d_text_req, const i64* d_text_doc, i64 ntext, cudaStream_t st, DevBuf<u64>& span_off,
                               DevBuf<i64>& spans, i64* total_out) {
    span_off.alloc((size_t)ntext + 1, st);
    *total_out = 0;
    if (ntext == 0 || nkw == 0) {
        CDB_CUDA(cudaMemsetAsync(span_off.p, 0, ((size_t)ntext + 1) * 8, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        return;
    }
    if (ntext >= ((i64)1 << 32)) throw Error(CDB_ERR_ARG, "cdb_locate_spans_batch: more than 2^32 texts in one batch");
    DevBuf<u64> code((size_t)nkw, st);
    DevBuf<unsigned long long> counters(2, st);  // [0] long texts, [1] error flags
    DevBuf<u32> long_list((size_t)ntext, st);
    CDB_CUDA(cudaMemsetAsync(counters.p, 0, 16, st));
    int* err = reinterpret_cast<int*>(counters.p + 1);
    span_kwinfo_kernel<<<(unsigned)ceil_div(nkw, 256), 256, 0, st>>>(d_kw, d_kw_off, nkw, code.p, err);
    CDB_LAUNCH_CHECK();
    SpanBatch b{ix.d_text, ix.d_off, ix.nd, d_kw, d_kw_off, d_req_kw_off, nkw, nreq, d_text_req, d_text_doc, ntext, code.p};
    const unsigned tgrid = (unsigned)ceil_div(ntext, 128);
    const unsigned wgrid = (unsigned)num_sms() * 4;
    spans_thread_kernel<false><<<tgrid, 128, 0, st>>>(b, span_off.p, nullptr, nullptr, long_list.p, counters.p, err);
    CDB_LAUNCH_CHECK();
    spans_warp_kernel<false><<<wgrid, 256, 0, st>>>(b, span_off.p, nullptr, nullptr, long_list.p, counters.p);
    CDB_LAUNCH_CHECK();
    prim::exclusive_scan<u64>(span_off.p, span_off.p, (u64)ntext, st);
    u64 total = 0;
    unsigned long long hc[2];
    CDB_CUDA(cudaMemcpyAsync(&total, span_off.p + ntext, 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaMemcpyAsync(hc, counters.p, 16, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    const int e = (int)(hc[1] & 0xffffffffu);
    if (e & 1) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");
    if (e & 2) throw Error(CDB_ERR_ARG, "cdb_locate_spans: document or request index out of range");
    *total_out = (i64)total;
    spans.alloc((size_t)(total ? total : 1) * 2, st);
    if (total == 0) return;
    spans_thread_kernel<true><<<tgrid, 128, 0, st>>>(b, nullptr, span_off.p, spans.p, nullptr, counters.p, err);
    CDB_LAUNCH_CHECK();
    if (hc[0]) {
        spans_warp_kernel<true><<<wgrid, 256, 0, st>>>(b, nullptr, span_off.p, spans.p, long_list.p, counters.p);
        CDB_LAUNCH_CHECK();
    }
}

// Host buffers in, host vectors out (cdb_locate_spans / cdb_locate_spans_batch).
void locate_spans_batch(const Index& ix, const u8* kw, const i64* kw_off, i64 nkw, const i64* req_kw_off, i64 nreq,
                        const i64* text_req, const i64* text_doc, i64 ntext, cudaStream_t st, std::vector<i64>& span_off,
                        std::vector<i64>& spans) {
    span_off.assign((size_t)ntext + 1, 0);
    spans.clear();
    for (i64 k = 0; k < nkw; ++k)
        if (kw_off[k + 1] <= kw_off[k]) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");
    if (ntext == 0 || nkw == 0 || ix.n == 0) {
        for (i64 t = 0; t < ntext; ++t)
            if (text_doc[t] < 0 || text_doc[t] >= ix.nd) throw Error(CDB_ERR_ARG, "cdb_locate_spans: document index out of range");
        return;
    }
    const i64 kbytes = kw_off[nkw] - kw_off[0];
    DevBuf<u8> d_kw((size_t)kbytes + 8, st);
    DevBuf<i64> d_koff((size_t)nkw + 1, st), d_rko((size_t)nreq + 1, st), d_treq((size_t)ntext, st), d_tdoc((size_t)ntext, st);
    std::vector<i64> rel((size_t)nkw + 1);
    for (i64 k = 0; k <= nkw; ++k) rel[k] = kw_off[k] - kw_off[0];
    CDB_CUDA(cudaMemsetAsync(d_kw.p + kbytes, 0, 8, st));
    CDB_CUDA(cudaMemcpyAsync(d_kw.p, kw + kw_off[0], (size_t)kbytes, cudaMemcpyHostToDevice, st));
    CDB_CUDA(cudaMemcpyAsync(d_koff.p, rel.data(), (size_t)(nkw + 1) * 8, cudaMemcpyHostToDevice, st));
    CDB_CUDA(cudaMemcpyAsync(d_rko.p, req_kw_off, (size_t)(nreq + 1) * 8, cudaMemcpyHostToDevice, st));
    CDB_CUDA(cudaMemcpyAsync(d_treq.p, text_req, (size_t)ntext * 8, cudaMemcpyHostToDevice, st));
    CDB_CUDA(cudaMemcpyAsync(d_tdoc.p, text_doc, (size_t)ntext * 8, cudaMemcpyHostToDevice, st));
    DevBuf<u64> d_soff;
    DevBuf<i64> d_spans;
    i64 total = 0;
    locate_spans_batch_device(ix, d_kw.p, d_koff.p, nkw, d_rko.p, nreq, d_treq.p, d_tdoc.p, ntext, st, d_soff, d_spans, &total);
    CDB_CUDA(cudaMemcpyAsync(span_off.data(), d_soff.p, (size_t)(ntext + 1) * 8, cudaMemcpyDeviceToHost, st));
    spans.resize((size_t)total * 2);
    if (total) CDB_CUDA(cudaMemcpyAsync(spans.data(), d_spans.p, (size_t)total * 16, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace cdb
