#pragma once
#include <string>

#include "index.cuh"

namespace cdb {
// persist.cu
void save_index(const Index& ix, const char* path, cudaStream_t st);
struct SavedArrayFile : SavedArraySource {
    std::string path;
    explicit SavedArrayFile(const char* p) : path(p) {}
    bool try_load(Index& ix, cudaStream_t st) const override;
};
}  // namespace cdb
