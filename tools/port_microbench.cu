// Micro-benchmark behind DESIGN.md §3.1: translate_kernel is bound by the per-SM L1 -> crossbar request port
// (l1tex__m_l1tex2xbar_req_cycles_active 72 %): one cycle per looked-up ids[] sector and one per 32 bytes stored.
// Question: do the pair stores leave that port when they are staged in shared memory and written with bulk async
// copies (TMA, cp.async.bulk.global.shared::cta) instead of STG?
//   lookup      random 8-byte lookups in a 32 MB table (L2 resident), results summed (no stores)
//   lookup+stg  the same + one coalesced 16-byte (id, 1) pair store per lookup
//   lookup+tma  the same, pairs staged in shared memory and stored with one bulk copy per 32 (or 128) pairs
//   stg / tma   stores only
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/port_microbench.cu -o tools/_build/port_microbench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

// L2 policies as in translate_kernel: the table is kept (evict_last), index reads and pair stores stream (evict_first)
__device__ __forceinline__ uint64_t pol_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t pol_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t ldg_na(const uint64_t* p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ldg_idx(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_pair(ulonglong2* p, ulonglong2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v.x), "l"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(s), "r"(bytes),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int U = 4;  // lookups per lane and round: a warp round = 128 entries = 2 KB of pairs

// MODE 0 lookups only, 1 lookups + STG, 2 lookups + TMA (PIECE pairs per bulk copy), 3 STG only, 4 TMA only
template <int MODE, int PIECE>
__global__ void __launch_bounds__(256) port_kernel(const uint64_t* __restrict__ table, const uint32_t* __restrict__ idx,
                                                   uint64_t n, ulonglong2* __restrict__ pairs, uint64_t* __restrict__ sink) {
    __shared__ __align__(128) ulonglong2 stage[8][2][32 * U];  // per warp, double buffered: 2 x 2 KB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t nround = n / (32 * U);
    uint64_t acc = 0;
    int buf = 0;
    const uint64_t keep = pol_last(), stream = pol_first();
    for (uint64_t rd = (uint64_t)blockIdx.x * 8 + warp; rd < nround; rd += (uint64_t)gridDim.x * 8) {
        const uint64_t base = rd * (32 * U);
        uint64_t v[U];
        if (MODE <= 2) {
            uint32_t d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) d[u] = ldg_idx(idx + base + u * 32 + lane, stream);
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_na(table + d[u], keep);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = base + u;
        }
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        } else if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int u = 0; u < U; ++u) stg_pair(pairs + base + u * 32 + lane, make_ulonglong2(v[u], 1), stream);
        } else {
            if (MODE == 2 || MODE == 4) {
                // the buffer written two rounds ago must have been read by its bulk copies
                bulk_wait_read1();  // bulk groups are per thread: every lane waits for its own copies
                __syncwarp();
#pragma unroll
                for (int u = 0; u < U; ++u) stage[warp][buf][u * 32 + lane] = make_ulonglong2(v[u], 1);
                fence_async();
                __syncwarp();
                constexpr int NP = 32 * U / PIECE;  // pieces per round
                if (lane < NP) bulk_store(pairs + base + lane * PIECE, &stage[warp][buf][lane * PIECE], PIECE * 16, stream);
                bulk_commit();  // one (possibly empty) group per lane and round
                buf ^= 1;
            }
        }
    }
    if (MODE == 2 || MODE == 4) {
        bulk_wait_read0();
        __syncwarp();
    }
    if (MODE == 0 && acc == 0x123456789abcdefull) sink[0] = acc;
}

int main() {
    const uint64_t T = 4u << 20, N = 1ull << 27;  // 2^27 entries: 2 GB of pairs, 512 MB of indices
    std::vector<uint32_t> h(N);
    uint64_t x = 88172645463325252ull;
    for (auto& v : h) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        v = (uint32_t)(x % T);
    }
    uint64_t *table, *sink;
    uint32_t* idx;
    ulonglong2* pairs;
    CK(cudaMalloc(&table, T * 8));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMalloc(&idx, N * 4));
    CK(cudaMalloc(&pairs, N * 16));
    CK(cudaMemset(table, 1, T * 8));
    CK(cudaMemcpy(idx, h.data(), N * 4, cudaMemcpyHostToDevice));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, auto launch) {
        for (int w = 0; w < 2; ++w) launch();
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 5;
        printf("%-16s %8.3f ms  %7.2f G entries/s  %.2f SM-cycles per entry at 1.965 GHz\n", name, ms, (double)N / ms / 1e6,
               ms * 1e-3 * 1.965e9 * sms / (double)N);
    };
    for (int occ : {3, 6}) {
        const int grid = sms * occ;
        printf("-- %d CTAs of 256 threads per SM\n", occ);
        time("lookup", [&] { port_kernel<0, 32><<<grid, 256>>>(table, idx, N, pairs, sink); });
        time("lookup+stg", [&] { port_kernel<1, 32><<<grid, 256>>>(table, idx, N, pairs, sink); });
        time("lookup+tma32", [&] { port_kernel<2, 32><<<grid, 256>>>(table, idx, N, pairs, sink); });
        time("lookup+tma128", [&] { port_kernel<2, 128><<<grid, 256>>>(table, idx, N, pairs, sink); });
        time("stg", [&] { port_kernel<3, 32><<<grid, 256>>>(table, idx, N, pairs, sink); });
        time("tma32", [&] { port_kernel<4, 32><<<grid, 256>>>(table, idx, N, pairs, sink); });
        time("tma128", [&] { port_kernel<4, 128><<<grid, 256>>>(table, idx, N, pairs, sink); });
    }
    return 0;
}
