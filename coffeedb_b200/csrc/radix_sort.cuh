// Device-wide stable LSD radix sort of (u64 key, value) pairs — "onesweep" organisation:
//   * one up-front kernel histograms every 8-bit digit of every pass (keys read once),
//   * each pass is ONE kernel: a tile of 4096 keys is ranked inside the CTA (warp-level match_any multi-split,
//     staged through shared memory so that global stores leave in digit-contiguous runs) and its global
//     offsets come from a decoupled look-back over per-tile status words — no separate scan kernel and no
//     second read of the keys.
// Per pass and element the traffic is therefore read (8+sizeof V) + write (8+sizeof V) bytes: the HBM
// roofline of a pass over n pairs is n * 2 * (8 + sizeof V) / BW.  No tensor cores: pure integer movement.
//
// This is the engine under K2/K3 of SURVEY.md §2a (the reference's MSD radix + std::sort leaves,
// src/index.cpp:75-128) and under the large-interval doc sort of locate (src/index.cpp:294-315).
#pragma once
#include <vector>

#include "common.cuh"

namespace cdb {
namespace rs {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int THREADS = 256;  // == RADIX: thread d owns digit d during the look-back
constexpr int WARPS = THREADS / 32;
constexpr int IPT = 16;
constexpr int TILE = THREADS * IPT;
constexpr int MAX_PASSES = 8;

constexpr u64 FLAG_LOCAL = 1ull << 62;
constexpr u64 FLAG_INCL = 2ull << 62;
constexpr u64 VALUE_MASK = (1ull << 62) - 1;

struct NoValue {};

struct PassDesc {
    int shift[MAX_PASSES];
    u32 mask[MAX_PASSES];
    int npass;
};

// ---- histogram of all digits of all passes --------------------------------------------------------------
static __global__ void __launch_bounds__(512) hist_kernel(const u64* __restrict__ keys, u64 n, PassDesc pd,
                                                   u64* __restrict__ ghist) {
    __shared__ u32 sh[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < pd.npass * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    // two keys per 16-byte load
    const u64 npair = n >> 1;
    const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += stride) {
        ulonglong2 kk = k2[i];
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p) {
            if (p < pd.npass) {
                atomicAdd(&sh[p * RADIX + ((u32)(kk.x >> pd.shift[p]) & pd.mask[p])], 1u);
                atomicAdd(&sh[p * RADIX + ((u32)(kk.y >> pd.shift[p]) & pd.mask[p])], 1u);
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        u64 k = keys[n - 1];
        for (int p = 0; p < pd.npass; ++p) atomicAdd(&sh[p * RADIX + ((u32)(k >> pd.shift[p]) & pd.mask[p])], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pd.npass * RADIX; i += blockDim.x) {
        u32 c = sh[i];
        if (c) atomicAdd(reinterpret_cast<unsigned long long*>(&ghist[i]), (unsigned long long)c);
    }
}

template <typename V>
struct ValueTraits {
    static constexpr bool has = true;
    static constexpr int bytes = sizeof(V);
};
template <>
struct ValueTraits<NoValue> {
    static constexpr bool has = false;
    static constexpr int bytes = 0;
};

template <typename V>
constexpr size_t onesweep_smem_bytes() {
    return (size_t)TILE * 8 + (size_t)TILE * ValueTraits<V>::bytes;
}

// ---- one pass -----------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(THREADS, 3) onesweep_kernel(const u64* __restrict__ kin, u64* __restrict__ kout,
                                                           const V* __restrict__ vin, V* __restrict__ vout, u64 n,
                                                           int shift, u32 dmask, const u64* __restrict__ digit_base,
                                                           u64* status, u32* tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* skeys = reinterpret_cast<u64*>(smem_raw);
    __shared__ u16 wc[WARPS][RADIX];  // per-warp digit counts (<= 512), then exclusive offsets (<= 4096)
    __shared__ u32 bin_start[RADIX];
    __shared__ u64 gbase[RADIX];
    __shared__ u32 warp_tot[WARPS];
    __shared__ u32 s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < WARPS * RADIX; i += THREADS) (&wc[0][0])[i] = 0;
    __syncthreads();
    const u64 tile = s_tile;
    const u64 tile_base = tile * TILE;
    const u32 nvalid = (u32)((n - tile_base) < (u64)TILE ? (n - tile_base) : (u64)TILE);

    // warp-striped load: item r of lane l in warp w is element w*512 + r*32 + l of the tile (stable order)
    u64 key[IPT];
    const u32 wbase = warp * (32 * IPT) + lane;
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        u32 li = wbase + r * 32;
        key[r] = li < nvalid ? ld_stream_u64(kin + tile_base + li) : ~0ull;
    }
    // rank inside the warp, digit by digit (match_any multi-split); counters are per warp
    u32 rank[IPT];
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        u32 d = (u32)(key[r] >> shift) & dmask;
        // lanes holding the same digit: 8 ballots (ALU pipe).  MATCH.ANY does this in one instruction but runs on the
        // ADU pipe, which the ncu capture showed 56 % busy and the kernel's limiter (profiles/README.md).
        u32 peers = 0xffffffffu;
#pragma unroll
        for (int bit = 0; bit < RADIX_BITS; ++bit) {
            const bool one = (d >> bit) & 1u;
            const u32 bal = __ballot_sync(0xffffffffu, one);
            peers &= one ? bal : ~bal;
        }
        int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader) {
            old = wc[warp][d];
            wc[warp][d] = (u16)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // thread d: exclusive scan of digit d's count over the warps; total = this tile's count of digit d
    u32 run = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        u32 c = wc[w][tid];
        wc[w][tid] = (u16)run;
        run += c;
    }
    u64* my_status = status + tile * RADIX + tid;
    st_relaxed_u64(my_status, (tile == 0 ? FLAG_INCL : FLAG_LOCAL) | (u64)run);
    // block-wide exclusive scan of the digit totals -> start of each digit inside the tile
    u32 incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    u32 woff = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) woff += (w < warp) ? warp_tot[w] : 0;
    const u32 bstart = woff + incl - run;
    bin_start[tid] = bstart;
    // decoupled look-back: sum the counts of digit d over all earlier tiles
    u64 excl = 0;
    if (tile > 0) {
        for (i64 t = (i64)tile - 1; t >= 0; --t) {
            const u64* sp = status + (u64)t * RADIX + tid;
            u64 v;
            do {
                v = ld_relaxed_u64(sp);
            } while ((v >> 62) == 0);
            excl += v & VALUE_MASK;
            if ((v >> 62) == 2) break;
        }
        st_relaxed_u64(my_status, FLAG_INCL | (excl + run));
    }
    gbase[tid] = digit_base[tid] + excl - (u64)bstart;
    __syncthreads();
    // scatter into shared memory in final tile order
#pragma unroll
    for (int r = 0; r < IPT; ++r) {
        u32 d = (u32)(key[r] >> shift) & dmask;
        u32 pos = bin_start[d] + wc[warp][d] + rank[r];
        skeys[pos] = key[r];
        rank[r] = pos;
    }
    if constexpr (ValueTraits<V>::has) {
        V* svals = reinterpret_cast<V*>(smem_raw + (size_t)TILE * 8);
        V val[IPT];
#pragma unroll
        for (int r = 0; r < IPT; ++r) {
            u32 li = wbase + r * 32;
            if (li < nvalid) val[r] = vin[tile_base + li];
        }
#pragma unroll
        for (int r = 0; r < IPT; ++r) {
            u32 li = wbase + r * 32;
            if (li < nvalid) svals[rank[r]] = val[r];
        }
    }
    __syncthreads();
    // digit-contiguous runs leave as coalesced stores
    if constexpr (ValueTraits<V>::has) {
        const V* svals = reinterpret_cast<const V*>(smem_raw + (size_t)TILE * 8);
#pragma unroll 4
        for (u32 i = tid; i < nvalid; i += THREADS) {
            u64 k = skeys[i];
            u64 dst = gbase[(u32)(k >> shift) & dmask] + i;
            kout[dst] = k;
            vout[dst] = svals[i];
        }
    } else {
#pragma unroll 4
        for (u32 i = tid; i < nvalid; i += THREADS) {
            u64 k = skeys[i];
            kout[gbase[(u32)(k >> shift) & dmask] + i] = k;
        }
    }
}

struct SortStats {
    int passes_run = 0;
    int passes_skipped = 0;
};

// Sorts n pairs by key bits [begin_bit, end_bit) (stable).  Input in (k0, v0); (k1, v1) are scratch of the
// same size.  Returns 0 if the sorted data ended in (k0, v0), 1 if in (k1, v1).  Passes whose digit is
// constant over all keys are skipped.  Synchronises the stream once (to read the histogram).
template <typename V>
int radix_sort_pairs(u64* k0, u64* k1, V* v0, V* v1, u64 n, int begin_bit, int end_bit, cudaStream_t st,
                     SortStats* stats = nullptr, V* final_vout = nullptr, bool* used_final = nullptr) {
    if (n <= 1 || end_bit <= begin_bit) return 0;
    PassDesc pd;
    pd.npass = 0;
    for (int b = begin_bit; b < end_bit; b += RADIX_BITS) {
        int bits = end_bit - b < RADIX_BITS ? end_bit - b : RADIX_BITS;
        pd.shift[pd.npass] = b;
        pd.mask[pd.npass] = (1u << bits) - 1;
        pd.npass++;
    }
    for (int p = pd.npass; p < MAX_PASSES; ++p) {
        pd.shift[p] = 0;
        pd.mask[p] = 0;
    }
    const u64 ntiles = (n + TILE - 1) / TILE;
    DevBuf<u64> ghist((size_t)MAX_PASSES * RADIX * 2, st);  // [hist | digit_base]
    DevBuf<u64> status((size_t)ntiles * RADIX, st);
    DevBuf<u32> counter(MAX_PASSES, st);
    CDB_CUDA(cudaMemsetAsync(ghist.p, 0, ghist.bytes(), st));
    CDB_CUDA(cudaMemsetAsync(counter.p, 0, counter.bytes(), st));
    {
        u64 want = (n / 2 + 511) / 512;
        int grid = (int)(want < (u64)(num_sms() * 4) ? (want ? want : 1) : (u64)(num_sms() * 4));
        hist_kernel<<<grid, 512, 0, st>>>(k0, n, pd, ghist.p);
        CDB_LAUNCH_CHECK();
    }
    std::vector<u64> h((size_t)MAX_PASSES * RADIX);
    CDB_CUDA(cudaMemcpyAsync(h.data(), ghist.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
    CDB_CUDA(cudaStreamSynchronize(st));
    std::vector<u64> base((size_t)MAX_PASSES * RADIX, 0);
    bool run[MAX_PASSES];
    for (int p = 0; p < pd.npass; ++p) {
        u64 acc = 0;
        run[p] = true;
        for (int d = 0; d < RADIX; ++d) {
            u64 c = h[(size_t)p * RADIX + d];
            if (c == n) run[p] = false;
            base[(size_t)p * RADIX + d] = acc;
            acc += c;
        }
    }
    u64* dbase = ghist.p + (size_t)MAX_PASSES * RADIX;
    CDB_CUDA(cudaMemcpyAsync(dbase, base.data(), base.size() * 8, cudaMemcpyHostToDevice, st));
    static bool attr_set = false;
    const size_t smem = onesweep_smem_bytes<V>();
    CDB_CUDA(cudaFuncSetAttribute(onesweep_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)attr_set;
    int cur = 0;
    u64* kb[2] = {k0, k1};
    V* vb[2] = {v0, v1};
    // final_vout: the last pass that runs scatters the values straight into the caller's destination (the suffix-array
    // slice of a chunk) instead of the ping-pong buffer; *used_final tells whether any pass ran
    int last_run = -1, nrun = 0;
    for (int p = 0; p < pd.npass; ++p)
        if (run[p]) {
            last_run = p;
            ++nrun;
        }
    // final_vout may be one of the two ping-pong buffers (a chunk sorts between its workspace and its suffix-array
    // range).  If the natural ping-pong already ends there nothing special happens; if it ends in the other one the
    // last pass would have to scatter in place, so the passes run naturally and one copy follows.
    bool copy_after = false;
    if (ValueTraits<V>::has && final_vout != nullptr && (final_vout == v0 || final_vout == v1)) {
        V* natural_end = (nrun % 2) ? v1 : v0;
        if (final_vout != natural_end) copy_after = nrun > 0;
        if (used_final) *used_final = nrun > 0;
        final_vout = nullptr;  // plain ping-pong
        if (copy_after) final_vout = nullptr;
    } else if (used_final) {
        *used_final = final_vout != nullptr && last_run >= 0;
    }
    V* const alias_dest = copy_after ? ((nrun % 2) ? v0 : v1) : nullptr;
    for (int p = 0; p < pd.npass; ++p) {
        if (!run[p]) {
            if (stats) stats->passes_skipped++;
            continue;
        }
        CDB_CUDA(cudaMemsetAsync(status.p, 0, status.bytes(), st));
        V* vdst = (final_vout != nullptr && p == last_run) ? final_vout : vb[cur ^ 1];
        onesweep_kernel<V><<<(unsigned)ntiles, THREADS, smem, st>>>(kb[cur], kb[cur ^ 1], vb[cur], vdst, n,
                                                                    pd.shift[p], pd.mask[p], dbase + (size_t)p * RADIX,
                                                                    status.p, counter.p + p);
        CDB_LAUNCH_CHECK();
        cur ^= 1;
        if (stats) stats->passes_run++;
    }
    if (copy_after) {
        if constexpr (ValueTraits<V>::has)
            CDB_CUDA(cudaMemcpyAsync(alias_dest, vb[cur], n * sizeof(V), cudaMemcpyDeviceToDevice, st));
    }
    return cur;
}

}  // namespace rs
}  // namespace cdb
