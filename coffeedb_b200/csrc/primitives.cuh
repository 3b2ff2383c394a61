// Small device-wide primitives shared by build and locate: exclusive scan (u64 out) and reductions.
#pragma once
#include "common.cuh"

namespace cdb {
namespace prim {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

// block-wide exclusive scan of one u64 per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ u64 block_exclusive_scan_u64(u64 v, u64* total, u64* warp_sums /* [32] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u64 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protects warp_sums against a previous use
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    u64 woff = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        u64 s = warp_sums[w];
        if (w < warp) woff += s;
        tot += s;
    }
    *total = tot;
    return woff + incl - v;
}

// 32-bit variant (counts inside one tile)
__device__ __forceinline__ u32 block_exclusive_scan_u32(u32 v, u32* total, u32* warp_sums /* [32] shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protects warp_sums against a previous use
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    u32 woff = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        u32 s = warp_sums[w];
        if (w < warp) woff += s;
        tot += s;
    }
    *total = tot;
    return woff + incl - v;
}

template <typename Tin>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const Tin* __restrict__ in, u64 n, u64* __restrict__ bsum) {
    __shared__ u64 ws[32];
    u64 base = (u64)blockIdx.x * SCAN_TILE;
    u64 s = 0;
#pragma unroll
    for (int r = 0; r < SCAN_IPT; ++r) {
        u64 i = base + (u64)r * SCAN_THREADS + threadIdx.x;
        if (i < n) s += (u64)in[i];
    }
    u64 tot;
    block_exclusive_scan_u64(s, &tot, ws);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of bsum[nb]; bsum[nb] receives the grand total
static __global__ void __launch_bounds__(1024) scan_blocksums_kernel(u64* bsum, u64 nb) {
    __shared__ u64 ws[32];
    __shared__ u64 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < nb; base += 1024) {
        u64 i = base + threadIdx.x;
        u64 v = i < nb ? bsum[i] : 0;
        u64 tot;
        u64 ex = block_exclusive_scan_u64(v, &tot, ws);
        u64 carry = carry_s;
        if (i < nb) bsum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry_s;
}

template <typename Tin>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const Tin* in, u64 n, const u64* __restrict__ bsum, u64* out) {
    __shared__ u64 ws[32];
    // thread-contiguous items so that the per-thread prefix is a running sum
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_IPT;
    u64 v[SCAN_IPT];
    u64 s = 0;
#pragma unroll
    for (int r = 0; r < SCAN_IPT; ++r) {
        u64 i = base + r;
        v[r] = i < n ? (u64)in[i] : 0;
        s += v[r];
    }
    u64 tot;
    u64 ex = block_exclusive_scan_u64(s, &tot, ws) + bsum[blockIdx.x];
#pragma unroll
    for (int r = 0; r < SCAN_IPT; ++r) {
        u64 i = base + r;
        if (i < n) out[i] = ex;
        ex += v[r];
    }
}

// out[i] = sum(in[0..i)), out[n] = total (out has n+1 elements).  In-place allowed when Tin is 8 bytes wide.
template <typename Tin>
inline void exclusive_scan(const Tin* in, u64* out, u64 n, cudaStream_t st) {
    if (n == 0) {
        CDB_CUDA(cudaMemsetAsync(out, 0, 8, st));
        return;
    }
    u64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf<u64> bsum(nb + 1, st);
    scan_reduce_kernel<Tin><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, bsum.p);
    CDB_LAUNCH_CHECK();
    scan_blocksums_kernel<<<1, 1024, 0, st>>>(bsum.p, nb);
    CDB_LAUNCH_CHECK();
    scan_apply_kernel<Tin><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, bsum.p, out);
    CDB_LAUNCH_CHECK();
    CDB_CUDA(cudaMemcpyAsync(out + n, bsum.p + nb, 8, cudaMemcpyDeviceToDevice, st));
}

}  // namespace prim
}  // namespace cdb
