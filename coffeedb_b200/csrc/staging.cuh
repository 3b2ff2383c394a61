// Loader staging.  The reference's loader keeps one std::string per field and hands string_index::add a
// view of it (src/database.cpp:173-275, 263-264: 32 B of std::string + 16 B of string_view per document, and the bytes stay
// on the host for good).  Here cdb_add streams the bytes straight into page-locked chunks; a chunk that fills up leaves for
// the device at once on the staging's own copy stream while the loader fills the next one, so that cdb_build finds the text
// (all but the last chunk) already in HBM and only stitches the chunks together device-to-device.  Without a CUDA device
// the chunks are plain host memory and nothing is uploaded (cdb_add has to work on a box that only inserts).
#pragma once
#include <cstring>
#include <vector>

#include "common.cuh"

namespace cdb {

class TextStaging {
public:
    TextStaging() {}
    TextStaging(const TextStaging&) = delete;
    TextStaging& operator=(const TextStaging&) = delete;
    ~TextStaging() { clear(); }

    size_t size() const { return total_; }
    size_t uploaded_bytes() const {
        size_t s = 0;
        for (const Chunk& c : chunks_)
            if (c.d) s += c.used;
        return s;
    }

    // device the full chunks are sent to (-1: the device current at the first upload)
    void set_device(int device) { device_ = device; }

    void append(const u8* p, size_t len) {
        while (len) {
            if (chunks_.empty() || chunks_.back().used == chunks_.back().cap) {
                if (!chunks_.empty()) upload(chunks_.back());
                new_chunk();
            }
            Chunk& c = chunks_.back();
            const size_t m = std::min(len, c.cap - c.used);
            std::memcpy(c.h + c.used, p, m);
            c.used += m;
            total_ += m;
            p += m;
            len -= m;
        }
    }

    // back to an earlier size (cdb_add_many rolling back after a failure)
    void truncate(size_t new_size) {
        while (total_ > new_size) {
            Chunk& c = chunks_.back();
            const size_t drop = std::min(c.used, total_ - new_size);
            c.used -= drop;
            total_ -= drop;
            if (drop && c.d) free_device_copy(c);  // its device copy no longer matches
            if (c.used == 0 && total_ > new_size) {
                free_chunk(c);
                chunks_.pop_back();
            }
        }
    }

    // all bytes, in order, to d_dst (device memory of `size()` bytes) on stream st
    void assemble(void* d_dst, cudaStream_t st) {
        size_t o = 0;
        for (Chunk& c : chunks_) {
            if (!c.used) continue;
            if (c.d) {
                CDB_CUDA(cudaStreamWaitEvent(st, c.done, 0));
                CDB_CUDA(cudaMemcpyAsync((u8*)d_dst + o, c.d, c.used, cudaMemcpyDefault, st));
            } else {
                CDB_CUDA(cudaMemcpyAsync((u8*)d_dst + o, c.h, c.used, cudaMemcpyHostToDevice, st));
            }
            o += c.used;
        }
    }

    // after the build has consumed the text: the per-chunk device copies go; the host bytes stay (keep_host_copy)
    void release_device_copies() {
        for (Chunk& c : chunks_)
            if (c.d) free_device_copy(c);
    }

    void clear() {
        for (Chunk& c : chunks_) free_chunk(c);
        chunks_.clear();
        total_ = 0;
        if (stream_) {
            cudaStreamDestroy(stream_);
            stream_ = nullptr;
        }
    }

private:
    struct Chunk {
        u8* h = nullptr;
        size_t cap = 0, used = 0;
        bool pinned = false;
        void* d = nullptr;  // device copy of a full chunk (sent when it filled up)
        cudaEvent_t done = nullptr;
        int ddev = -1;
    };
    static constexpr size_t kFirstChunk = (size_t)1 << 20, kMaxChunk = (size_t)64 << 20;

    static bool have_device() {
        static const bool v = [] {
            int cnt = 0;
            const bool ok = cudaGetDeviceCount(&cnt) == cudaSuccess && cnt > 0;
            if (!ok) cudaGetLastError();
            return ok;
        }();
        return v;
    }

    void new_chunk() {
        Chunk c;
        c.cap = chunks_.empty() ? kFirstChunk : std::min(kMaxChunk, chunks_.back().cap * 2);
        if (have_device() && cudaHostAlloc((void**)&c.h, c.cap, cudaHostAllocDefault) == cudaSuccess) {
            c.pinned = true;
        } else {
            cudaGetLastError();
            c.h = static_cast<u8*>(std::malloc(c.cap));
            if (!c.h) throw std::bad_alloc();
        }
        chunks_.push_back(c);
    }

    // best effort: a chunk that cannot be uploaded now (no device, no memory) simply goes up at build time
    void upload(Chunk& c) {
        if (!have_device() || !c.pinned || c.d || !c.used) return;
        int prev = -1;
        cudaGetDevice(&prev);
        int dev = device_ >= 0 ? device_ : prev;
        if (dev != prev && cudaSetDevice(dev) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        bool ok = true;
        if (!stream_) ok = cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaMalloc(&c.d, c.used) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(c.d, c.h, c.used, cudaMemcpyHostToDevice, stream_) == cudaSuccess;
        ok = ok && cudaEventRecord(c.done, stream_) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            if (c.d) cudaFree(c.d);
            if (c.done) cudaEventDestroy(c.done);
            c.d = nullptr;
            c.done = nullptr;
        } else {
            c.ddev = dev;
        }
        if (dev != prev && prev >= 0) cudaSetDevice(prev);
    }

    void free_device_copy(Chunk& c) {
        if (c.done) {
            cudaEventSynchronize(c.done);
            cudaEventDestroy(c.done);
        }
        if (c.d) cudaFree(c.d);
        c.d = nullptr;
        c.done = nullptr;
    }

    void free_chunk(Chunk& c) {
        free_device_copy(c);
        if (c.h) {
            if (c.pinned)
                cudaFreeHost(c.h);
            else
                std::free(c.h);
        }
        c.h = nullptr;
    }

    std::vector<Chunk> chunks_;
    size_t total_ = 0;
    int device_ = -1;
    cudaStream_t stream_ = nullptr;
};

}  // namespace cdb
