// Micro-benchmark behind DESIGN.md §3.1 (translate_kernel): what does one random 8-byte table lookup cost on a B200 SM,
// by access path?  A 32 MB table (one doc range of ids[]) stays L2-resident; 2^26 random indices are read coalesced.
//   ldg       one LDG.64 per lane, 32 distinct 128-byte lines per warp instruction (what translate_kernel does)
//   ldg1      the same lookups issued as 32 single-lane LDGs (one line per instruction)
//   tex       tex1Dfetch<uint2> through a texture object over the same table
//   lds       the table slice in shared memory (16K entries), random LDS.64
// Build + run (on a GPU box):  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_microbench.cu -o tools/_build/gather_microbench && tools/_build/gather_microbench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e = (x);                                                              \
        if (e != cudaSuccess) {                                                           \
            fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

constexpr int U = 4;

__device__ __forceinline__ uint64_t ldg_na(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) lookup_kernel(const uint64_t* __restrict__ table, cudaTextureObject_t tex,
                                                      const uint32_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    uint64_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base + (uint64_t)(U - 1) * blockDim.x < n; base += stride) {
        uint32_t d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) d[u] = idx[base + (uint64_t)u * blockDim.x];
        uint64_t v[U];
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_na(table + d[u]);
        } else if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = 0;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (lane == k) v[u] = ldg_na(table + d[u]);
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint2 t = tex1Dfetch<uint2>(tex, (int)d[u]);
                v[u] = ((uint64_t)t.y << 32) | t.x;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u];
    }
    if (acc == 0x123456789abcdefull) out[0] = acc;  // keep the loads alive
}

__global__ void __launch_bounds__(256) lds_kernel(const uint64_t* __restrict__ table, const uint32_t* __restrict__ idx,
                                                  uint64_t n, uint64_t* __restrict__ out) {
    extern __shared__ uint64_t s[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) s[i] = table[i];
    __syncthreads();
    uint64_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base + (uint64_t)(U - 1) * blockDim.x < n; base += stride) {
#pragma unroll
        for (int u = 0; u < U; ++u) acc += s[idx[base + (uint64_t)u * blockDim.x] & 16383];
    }
    if (acc == 0x123456789abcdefull) out[0] = acc;
}

// the store side: 16-byte pairs written (a) coalesced, (b) to 32 different rows per warp instruction
template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(const uint32_t* __restrict__ idx, uint64_t n, ulonglong2* __restrict__ out,
                                                    uint64_t nout) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t dst = MODE == 0 ? i % nout : ((uint64_t)idx[i] * 2654435761ull) % nout;
        out[dst] = make_ulonglong2(i, 1);
    }
}

int main() {
    const uint64_t T = 4u << 20, N = 1ull << 26;
    std::vector<uint32_t> h(N);
    uint64_t x = 88172645463325252ull;
    for (auto& v : h) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        v = (uint32_t)(x % T);
    }
    uint64_t *table, *out;
    uint32_t* idx;
    ulonglong2* pairs;
    const uint64_t NOUT = 1ull << 26;  // 1 GB of pairs
    CK(cudaMalloc(&table, T * 8));
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&idx, N * 4));
    CK(cudaMalloc(&pairs, NOUT * 16));
    CK(cudaMemset(table, 1, T * 8));
    CK(cudaMemcpy(idx, h.data(), N * 4, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<uint2>();
    rd.res.linear.sizeInBytes = T * 8;
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, auto launch, double per) {
        for (int w = 0; w < 2; ++w) launch();
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 5;
        printf("%-10s %8.3f ms  %7.2f G/s  %.2f SM-cycles per element at 1.965 GHz\n", name, ms, per / ms / 1e6,
               ms * 1e-3 * 1.965e9 * sms / per);
        CK(cudaGetLastError());
    };
    for (int occ : {3, 6, 8}) {
        const int grid = sms * occ;
        printf("-- %d CTAs of 256 threads per SM\n", occ);
        time("ldg", [&] { lookup_kernel<0><<<grid, 256>>>(table, tex, idx, N, out); }, (double)N);
        time("ldg1", [&] { lookup_kernel<1><<<grid, 256>>>(table, tex, idx, N, out); }, (double)N);
        time("tex", [&] { lookup_kernel<2><<<grid, 256>>>(table, tex, idx, N, out); }, (double)N);
    }
    CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    time("lds", [&] { lds_kernel<<<sms, 256, 131072>>>(table, idx, N, out); }, (double)N);
    time("lds x4", [&] { lds_kernel<<<sms, 1024, 131072>>>(table, idx, N, out); }, (double)N);
    time("st coal", [&] { store_kernel<0><<<sms * 8, 256>>>(idx, N, pairs, NOUT); }, (double)N);
    time("st rand", [&] { store_kernel<1><<<sms * 8, 256>>>(idx, N, pairs, NOUT); }, (double)N);
    return 0;
}
