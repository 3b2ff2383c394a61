// coffeedb_b200 — shared device/host helpers.  sm_100a only (B200): 148 SMs, 32 B DRAM sectors, 126 MB L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/coffeedb_b200.h"

namespace cdb {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

constexpr int kNumSMs = 148;  // B200: only the fallback of num_sms() when the device cannot be asked

// SM count of the current device (persistent grids are sized from it)
inline int num_sms() {
    static thread_local int cached_dev = -1, cached = kNumSMs;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMs;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

struct Error : std::runtime_error {
    cdb_status code;
    Error(cdb_status c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear sticky-less errors
        cdb_status c = (e == cudaErrorMemoryAllocation) ? CDB_ERR_NOMEM : CDB_ERR_CUDA;
        throw Error(c, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
                           std::to_string(line) + ")");
    }
}
#define CDB_CUDA(x) ::cdb::cuda_check((x), #x, __FILE__, __LINE__)
// every kernel launch of this library goes through this macro: it checks the launch and counts it
// (cdb_launch_count() — bench.py reports the count as gpu_launches)
extern std::atomic<unsigned long long> g_launches;
#define CDB_LAUNCH_CHECK()                                                             \
    do {                                                                               \
        ::cdb::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
        ::cdb::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__);    \
    } while (0)

// per-thread timing of the last locate call (CUDA events on the launching stream), milliseconds
struct LocateStats {
    float search_ms = 0, gather_ms = 0, large_ms = 0, tail_ms = 0, translate_ms = 0, total_ms = 0, listing_ms = 0;
    long long npat = 0, total_pairs = 0, total_occ = 0, nlarge = 0, nlisted = 0, listed_pairs = 0;
};
extern thread_local LocateStats g_locate_stats;

// Stream-ordered temporary device buffer (cudaMallocAsync pool): allocation is cheap after warm-up and the
// calls are re-entrant, which the locate path needs (several host threads query one index concurrently).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; s = o.s;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    void alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream;
        n = count;
        size_t bytes = (count ? count : 1) * sizeof(T);
        CDB_CUDA(cudaMallocAsync((void**)&p, bytes, stream));
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    T* detach() { T* q = p; p = nullptr; n = 0; return q; }
    ~DevBuf() { release(); }
    size_t bytes() const { return n * sizeof(T); }
};

// Large, long-lived or build-sized device buffer: plain cudaMalloc / cudaFree.  Growing the stream-ordered pool by
// tens of gigabytes goes through the virtual-memory-management path and costs seconds at the 10 GB configuration;
// it would also keep the build workspace inside the pool after the build.
template <typename T>
struct BigBuf {
    T* p = nullptr;
    size_t n = 0;
    BigBuf() {}
    explicit BigBuf(size_t count) { alloc(count); }
    BigBuf(const BigBuf&) = delete;
    BigBuf& operator=(const BigBuf&) = delete;
    void alloc(size_t count) {
        release();
        n = count;
        if (const char* e = getenv("CDB_BIGBUF_POOL")) pooled = atoi(e) != 0;  // experiment knob
        if (pooled)
            CDB_CUDA(cudaMallocAsync((void**)&p, (count ? count : 1) * sizeof(T), nullptr));
        else
            CDB_CUDA(cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    T* detach() { T* q = p; p = nullptr; n = 0; return q; }
    ~BigBuf() { release(); }
    bool pooled = false;
};

// Per host thread and device: one non-blocking stream and a set of timing events, created on first use and kept (a
// query() per request must not pay stream and event creation every time).
struct ThreadCtx {
    static constexpr int kEvents = 10;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // device-to-host copies that run under the next part's kernels (cdb_filter)
    cudaEvent_t ev[kEvents] = {};
    ~ThreadCtx() {
        if (!stream) return;
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        cudaStreamDestroy(stream);
        if (prev >= 0) cudaSetDevice(prev);
    }
};
ThreadCtx& thread_ctx(int device);  // capi.cu; the device must be current

inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------- device side
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// 64-bit load that bypasses L1 allocation: for streams read exactly once (suffix-array intervals, sort passes)
__device__ __forceinline__ u64 ld_stream_u64(const u64* p) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_stream_u32(const u32* p) {
    u32 v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 ld_stream(const u64* p) { return ld_stream_u64(p); }
__device__ __forceinline__ u32 ld_stream(const u32* p) { return ld_stream_u32(p); }
__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
// L2 eviction-priority policies (createpolicy) and loads/stores that carry one: translate_kernel keeps the slice of
// ids[] it is working on resident (evict_last) while rows stream through (evict_first).
__device__ __forceinline__ u64 l2_policy_evict_last() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_first() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_normal() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_unchanged() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 ld_hint_u64(const u64* p, u64 policy) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_hint_v2(i64* p, longlong2 v, u64 policy) {
    asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(p), "l"(v.x), "l"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ u32 ld_hint_u32(const u32* p, u64 policy) {
    u32 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_stream_u32(u32* p, u32 v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// streaming 8-byte store (written once, read by a later kernel)
__device__ __forceinline__ void st_stream_u64(u64* p, u64 v) {
    asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// relaxed gpu-scope load/store of one 64-bit word: the decoupled look-back status words
__device__ __forceinline__ u64 ld_relaxed_u64(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 8 bytes of text starting at an arbitrary byte address, returned big-endian (first byte most significant)
// so that unsigned integer comparison == memcmp order.  Reads the two aligned 8-byte words that cover the
// window; the text buffer is padded so that both are in bounds.
__device__ __forceinline__ u64 load_be64(const u8* text, i64 pos) {
    const u64* w = reinterpret_cast<const u64*>(text + (pos & ~(i64)7));
    u64 lo = __ldg(w), hi = __ldg(w + 1);
    u32 sh = (u32)(pos & 7) * 8;
    u64 v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;  // little-endian: byte at pos is the low byte
    u32 a = (u32)v, b = (u32)(v >> 32);
    a = __byte_perm(a, 0, 0x0123);
    b = __byte_perm(b, 0, 0x0123);
    return ((u64)a << 32) | b;
}

}  // namespace cdb
