// Micro-benchmark behind DESIGN.md §3.1 (translate_kernel): which L2 cache policies keep a randomly read table
// (one doc range of ids[]) resident while index reads and (id, count) pair stores stream through?  The translate
// kernel at cfg3 misses L2 on 37 % of its ids[] sectors (12 re-fetches of every sector per doc-range pass) although the
// slice is 32 MB and the L2 126 MB.  Each launch: N random 8-byte lookups in a table of T MB + 4 B/entry index stream
// + 16 B/entry pair stores.  Policies: n = no hint, f = evict_first, l = evict_last, m = evict_normal (hinted),
// u = evict_unchanged; persisting set-aside (cudaLimitPersistingL2CacheSize) and an access-policy window are separate axes.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2_microbench.cu -o tools/_build/l2_microbench
// Run under ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum for the DRAM bytes of every launch (same order as printed).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

enum Pol { P_NONE = 0, P_FIRST, P_LAST, P_NORMAL, P_UNCH };

__device__ __forceinline__ uint64_t make_pol(int p) {
    uint64_t r = 0;
    switch (p) {
        case P_FIRST: asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(r)); break;
        case P_LAST: asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(r)); break;
        case P_NORMAL: asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(r)); break;
        case P_UNCH: asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(r)); break;
        default: break;
    }
    return r;
}
__device__ __forceinline__ uint64_t ld64(const uint64_t* p, bool hint, uint64_t pol) {
    uint64_t v;
    if (hint)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    else
        asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld32(const uint32_t* p, bool hint, uint64_t pol) {
    uint32_t v;
    if (hint)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    else
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st128(ulonglong2* p, ulonglong2 v, bool hint, uint64_t pol) {
    if (hint)
        asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v.x), "l"(v.y), "l"(pol) : "memory");
    else
        asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

constexpr int U = 4;

// Segmented variant (what translate_kernel really does): the stream is cut into segments of SEG = 34 entries (one
// pattern's entries inside one doc range); segment s is read from / written to slot (s * odd) mod nseg, i.e. 136-byte
// index reads and 544-byte pair writes at scattered, 8- / 16-byte aligned places.  scat bit 0: index reads, bit 1: stores.
constexpr int SEG = 34;
__global__ void __launch_bounds__(256) l2_seg_kernel(const uint64_t* __restrict__ table, const uint32_t* __restrict__ idx,
                                                     uint64_t nseg_log2, ulonglong2* __restrict__ pairs, int ptab, int pidx, int pst,
                                                     int scat) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t nseg = 1ull << nseg_log2, n = nseg * SEG;
    const uint64_t nround = n / (32 * U);
    const uint64_t poltab = make_pol(ptab), polidx = make_pol(pidx), polst = make_pol(pst);
    for (uint64_t rd = (uint64_t)blockIdx.x * 8 + warp; rd < nround; rd += (uint64_t)gridDim.x * 8) {
        const uint64_t base = rd * (32 * U);
        uint32_t d[U];
        uint64_t v[U], dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t e = base + u * 32 + lane, sg = e / SEG, w = e % SEG;
            const uint64_t slot = ((sg * 2654435761ull) & (nseg - 1)) * SEG + w;
            dst[u] = (scat & 2) ? slot : e;
            d[u] = ld32(idx + ((scat & 1) ? slot : e), pidx != P_NONE, polidx);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld64(table + d[u], ptab != P_NONE, poltab);
#pragma unroll
        for (int u = 0; u < U; ++u) st128(pairs + dst[u], make_ulonglong2(v[u], 1), pst != P_NONE, polst);
    }
}

// stores: 0 = none (lookups summed), 1 = 16-byte pair per entry
__global__ void __launch_bounds__(256) l2_kernel(const uint64_t* __restrict__ table, const uint32_t* __restrict__ idx, uint64_t n,
                                                 ulonglong2* __restrict__ pairs, uint64_t* __restrict__ sink, int ptab, int pidx,
                                                 int pst, int stores) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t nround = n / (32 * U);
    const uint64_t poltab = make_pol(ptab), polidx = make_pol(pidx), polst = make_pol(pst);
    uint64_t acc = 0;
    for (uint64_t rd = (uint64_t)blockIdx.x * 8 + warp; rd < nround; rd += (uint64_t)gridDim.x * 8) {
        const uint64_t base = rd * (32 * U);
        uint32_t d[U];
        uint64_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) d[u] = ld32(idx + base + u * 32 + lane, pidx != P_NONE, polidx);
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld64(table + d[u], ptab != P_NONE, poltab);
        if (stores) {
#pragma unroll
            for (int u = 0; u < U; ++u) st128(pairs + base + u * 32 + lane, make_ulonglong2(v[u], 1), pst != P_NONE, polst);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        }
    }
    if (!stores && acc == 0x123456789abcdefull) sink[0] = acc;
}

int main(int argc, char** argv) {
    const uint64_t NSEG_LOG2 = 21;
    const uint64_t N = (1ull << NSEG_LOG2) * SEG;  // entries per launch: 71.3 M = 1.1 GB of pairs, 285 MB of indices
    const uint64_t TMAX = 128ull << 20;
    int sms = 0, l2 = 0, pmax = 0, wmax = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0));
    CK(cudaDeviceGetAttribute(&pmax, cudaDevAttrMaxPersistingL2CacheSize, 0));
    CK(cudaDeviceGetAttribute(&wmax, cudaDevAttrMaxAccessPolicyWindowSize, 0));
    printf("SMs %d, L2 %.1f MB, max persisting %.1f MB, max window %.1f MB\n", sms, l2 / 1048576.0, pmax / 1048576.0, wmax / 1048576.0);
    uint64_t *table, *sink;
    uint32_t* idx;
    ulonglong2* pairs;
    CK(cudaMalloc(&table, TMAX));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMalloc(&idx, N * 4));
    CK(cudaMalloc(&pairs, N * 16));
    CK(cudaMemset(table, 1, TMAX));
    std::vector<uint32_t> h(N);
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = sms * 3;
    auto run = [&](const char* name, int tmb, int ptab, int pidx, int pst, int stores) {
        for (int w = 0; w < 1; ++w) l2_kernel<<<grid, 256, 0, st>>>(table, idx, N, pairs, sink, ptab, pidx, pst, stores);
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < 3; ++r) l2_kernel<<<grid, 256, 0, st>>>(table, idx, N, pairs, sink, ptab, pidx, pst, stores);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 3;
        printf("T=%3d MB %-34s %7.3f ms  %.2f SM-cycles/entry\n", tmb, name, ms, ms * 1e-3 * 1.965e9 * sms / (double)N);
        fflush(stdout);
    };
    auto run_seg = [&](const char* name, int tmb, int ptab, int pidx, int pst, int scat) {
        l2_seg_kernel<<<grid, 256, 0, st>>>(table, idx, NSEG_LOG2, pairs, ptab, pidx, pst, scat);
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < 3; ++r) l2_seg_kernel<<<grid, 256, 0, st>>>(table, idx, NSEG_LOG2, pairs, ptab, pidx, pst, scat);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 3;
        printf("T=%3d MB %-34s %7.3f ms  %.2f SM-cycles/entry\n", tmb, name, ms, ms * 1e-3 * 1.965e9 * sms / (double)N);
        fflush(stdout);
    };
    const bool quick = argc > 1;  // any argument: only the segmented runs at 16 / 32 MB
    for (int tmb : {8, 16, 32, 64}) {
        const uint64_t T = ((uint64_t)tmb << 20) / 8;
        uint64_t x = 88172645463325252ull + tmb;
        for (auto& v : h) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            v = (uint32_t)(x % T);
        }
        CK(cudaMemcpy(idx, h.data(), N * 4, cudaMemcpyHostToDevice));
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
        if (tmb == 16 || tmb == 32) {
            for (int scat = 0; scat < 4; ++scat) {
                char nm[96];
                snprintf(nm, sizeof nm, "seg34 scat=%d tab=last st=first", scat);
                run_seg(nm, tmb, P_LAST, P_FIRST, P_FIRST, scat);
                snprintf(nm, sizeof nm, "seg34 scat=%d no hints", scat);
                run_seg(nm, tmb, P_NONE, P_NONE, P_NONE, scat);
                snprintf(nm, sizeof nm, "seg34 scat=%d tab=last only", scat);
                run_seg(nm, tmb, P_LAST, P_NONE, P_NONE, scat);
            }
        }
        if (quick) continue;
        run("lookup only, no hints", tmb, P_NONE, P_NONE, P_NONE, 0);
        run("lookup only, tab=last idx=first", tmb, P_LAST, P_FIRST, P_NONE, 0);
        run("no hints", tmb, P_NONE, P_NONE, P_NONE, 1);
        run("tab=last idx=first st=first", tmb, P_LAST, P_FIRST, P_FIRST, 1);
        run("tab=last idx=none st=none", tmb, P_LAST, P_NONE, P_NONE, 1);
        run("tab=none idx=first st=first", tmb, P_NONE, P_FIRST, P_FIRST, 1);
        run("tab=normal idx=first st=first", tmb, P_NORMAL, P_FIRST, P_FIRST, 1);
        run("tab=last idx=unch st=unch", tmb, P_LAST, P_UNCH, P_UNCH, 1);
        run("tab=none idx=unch st=unch", tmb, P_NONE, P_UNCH, P_UNCH, 1);
        for (int pmb : {32, 64, 96}) {
            size_t want = (size_t)pmb << 20;
            if (want > (size_t)pmax) want = pmax;
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            char nm[96];
            snprintf(nm, sizeof nm, "set-aside %d: tab=last st=first", pmb);
            run(nm, tmb, P_LAST, P_FIRST, P_FIRST, 1);
            // access-policy window over the table: hits persist, everything else is a normal access
            cudaStreamAttrValue a;
            memset(&a, 0, sizeof a);
            a.accessPolicyWindow.base_ptr = table;
            a.accessPolicyWindow.num_bytes = (size_t)tmb << 20;
            a.accessPolicyWindow.hitRatio = 1.0f;
            a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a));
            snprintf(nm, sizeof nm, "set-aside %d + window, no hints", pmb);
            run(nm, tmb, P_NONE, P_NONE, P_NONE, 1);
            snprintf(nm, sizeof nm, "set-aside %d + window, st=first", pmb);
            run(nm, tmb, P_NONE, P_FIRST, P_FIRST, 1);
            a.accessPolicyWindow.num_bytes = 0;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &a));
            CK(cudaCtxResetPersistingL2Cache());
        }
    }
    return 0;
}
