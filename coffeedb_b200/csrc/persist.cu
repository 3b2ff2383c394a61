// Suffix-array persistence.  The reference never persists its index: it is rebuilt from the raw/
// files at every start and every `build` (src/server.cpp:43-44, src/database.cpp:170-282), so build time is restart
// latency.  cdb_save writes the finished packed array next to the corpus; cdb_build_or_load uploads the staged corpus
// as cdb_build does and, when the file's key matches — a 64-bit hash of text, doc_off and ids, the geometry and the
// compat flag — streams the array back into device memory instead of sorting.
//   file = header | packed suffix array (n * width bytes, the reference's element format, note-N1 layout applied)
#include <cstdio>
#include <cstring>

#include "index.cuh"
#include "persist.cuh"

namespace cdb {

struct SavedHeader {
    char magic[8];  // "CDBSA001"
    u64 corpus_hash;
    i64 n, nd;
    int32_t width, bits1, bits2, compat_signed;
    u64 sa_bytes;
};
static const char kMagic[8] = {'C', 'D', 'B', 'S', 'A', '0', '0', '1'};
constexpr size_t kIoChunk = (size_t)64 << 20;

struct PinnedChunk {
    void* p = nullptr;
    PinnedChunk() { CDB_CUDA(cudaHostAlloc(&p, kIoChunk, cudaHostAllocDefault)); }
    ~PinnedChunk() { cudaFreeHost(p); }
};

void save_index(const Index& ix, const char* path, cudaStream_t st) {
    SavedHeader h{};
    std::memcpy(h.magic, kMagic, 8);
    h.corpus_hash = corpus_hash(ix, ix.n, st);
    h.n = ix.n;
    h.nd = ix.nd;
    h.width = ix.width;
    h.bits1 = ix.bits1;
    h.bits2 = ix.bits2;
    h.compat_signed = ix.opt.compat_signed;
    h.sa_bytes = (u64)ix.n * (u64)ix.width;
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) throw Error(CDB_ERR_ARG, std::string("cdb_save: cannot open ") + tmp);
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    PinnedChunk buf;
    for (u64 o = 0; ok && o < h.sa_bytes; o += kIoChunk) {
        const size_t m = (size_t)std::min<u64>(kIoChunk, h.sa_bytes - o);
        if (cudaMemcpyAsync(buf.p, (const u8*)ix.d_sa + o, m, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            std::fclose(f);
            std::remove(tmp.c_str());
            throw Error(CDB_ERR_CUDA, "cdb_save: device read failed");
        }
        ok = std::fwrite(buf.p, 1, m, f) == m;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok || std::rename(tmp.c_str(), path) != 0) {
        std::remove(tmp.c_str());
        throw Error(CDB_ERR_ARG, std::string("cdb_save: writing ") + path + " failed");
    }
}

bool SavedArrayFile::try_load(Index& ix, cudaStream_t st) const {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    SavedHeader h{};
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, kMagic, 8) == 0 && h.n == ix.n && h.nd == ix.nd &&
              h.width == ix.width && h.bits1 == ix.bits1 && h.bits2 == ix.bits2 && h.compat_signed == ix.opt.compat_signed &&
              h.sa_bytes == (u64)ix.n * (u64)ix.width;
    if (ok) ok = h.corpus_hash == corpus_hash(ix, ix.n, st);  // stale file (the corpus changed): build instead
    if (!ok) {
        std::fclose(f);
        return false;
    }
    void* d_sa = nullptr;
    CDB_CUDA(cudaMalloc(&d_sa, h.sa_bytes ? (size_t)h.sa_bytes : 1));
    PinnedChunk buf;
    for (u64 o = 0; ok && o < h.sa_bytes; o += kIoChunk) {
        const size_t m = (size_t)std::min<u64>(kIoChunk, h.sa_bytes - o);
        ok = std::fread(buf.p, 1, m, f) == m &&
             cudaMemcpyAsync((u8*)d_sa + o, buf.p, m, cudaMemcpyHostToDevice, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
    }
    std::fclose(f);
    if (!ok) {  // truncated file: fall back to the build
        cudaGetLastError();
        cudaFree(d_sa);
        return false;
    }
    ix.d_sa = d_sa;
    return true;
}

}  // namespace cdb
