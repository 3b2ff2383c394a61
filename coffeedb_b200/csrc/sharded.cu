// One string_index over several GPUs of ONE process (SURVEY.md §8e behind the drop-in boundary).  The reference is a single
// process (src/database.cpp:170-282, 387-393), so a sharded index has to sit behind the same add / build / query calls:
//
//   add      documents are staged on the host in call order (doc index = call order, src/index.cpp:174-177)
//   build    the staged documents are cut into contiguous doc-index ranges of about equal bytes, one per device; every
//            shard is an ordinary Index built by its own host thread, all at once — a suffix never leaves its document,
//            so the shards share nothing
//   locate   the packed pattern batch is uploaded to the first shard's device and copied device-to-device to the others
//            (NVLink where the devices are peers); every shard locates the batch on its own documents; the shards then
//            exchange their per-pattern row offsets (device-to-device again) and each one writes its part of every row
//            straight into the caller's result in page-locked host memory, at
//                global_row_off[q] + sum of the row lengths of the shards before it,
//            i.e. row q = the shard rows concatenated in shard order = ascending doc index = string_index::query on the
//            whole corpus (src/index.cpp:316-322).  Counts are per document and a document lives in one shard: nothing
//            is summed.  Each device uses its own PCIe link for its part of the result.
// One worker thread per shard lives as long as the handle (its stream and events stay warm); a call hands every worker
// its phase functions and the workers meet at host barriers between the phases.  Calls on one handle are serialised.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "index.cuh"
#include "locate.cuh"
#include "sharded.cuh"

namespace cdb {

struct ShardWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, quit = false, done = true;
    std::exception_ptr err;
};

class HostBarrier {
public:
    explicit HostBarrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != gen; });
        }
    }

private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

struct ShardedIndex {
    cdb_options opt{};
    std::vector<int> devs;
    std::vector<cdb_index*> shard;
    std::vector<i64> doc_begin;  // [nshards + 1]
    std::vector<u8> text;
    std::vector<i64> off{0};
    std::vector<i64> ids;
    bool built = false;
    std::mutex call_mu;
    std::vector<std::unique_ptr<ShardWorker>> workers;

    ~ShardedIndex() {
        for (auto& w : workers) {
            {
                std::lock_guard<std::mutex> lk(w->mu);
                w->quit = true;
            }
            w->cv.notify_all();
            if (w->th.joinable()) w->th.join();
        }
        for (cdb_index* s : shard) cdb_destroy(s);
    }

    void start_workers() {
        for (size_t g = 0; g < devs.size(); ++g) {
            workers.emplace_back(new ShardWorker());
            ShardWorker* w = workers.back().get();
            w->th = std::thread([w] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(w->mu);
                        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
                        if (w->quit) return;
                        job = std::move(w->job);
                        w->has_job = false;
                    }
                    std::exception_ptr e;
                    try {
                        job();
                    } catch (...) {
                        e = std::current_exception();
                    }
                    {
                        std::lock_guard<std::mutex> lk(w->mu);
                        w->err = e;
                        w->done = true;
                    }
                    w->cv.notify_all();
                }
            });
        }
    }

    // fn(g) on worker g for every shard at once; the first failure is rethrown after all have finished
    void run_all(const std::function<void(int)>& fn) {
        for (size_t g = 0; g < workers.size(); ++g) {
            ShardWorker* w = workers[g].get();
            {
                std::lock_guard<std::mutex> lk(w->mu);
                w->job = [fn, g] { fn((int)g); };
                w->has_job = true;
                w->done = false;
                w->err = nullptr;
            }
            w->cv.notify_all();
        }
        std::exception_ptr first;
        for (auto& w : workers) {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return w->done; });
            if (w->err && !first) first = w->err;
        }
        if (first) std::rethrow_exception(first);
    }
};

// A phase that failed on one shard must not leave the others waiting at the next barrier: every worker runs every phase
// inside this guard and keeps walking to the barriers; the first error is rethrown at the end of the call.
struct PhaseGuard {
    std::mutex mu;
    std::exception_ptr err;
    bool failed() {
        std::lock_guard<std::mutex> lk(mu);
        return (bool)err;
    }
    template <typename F>
    void run(F&& f) {
        if (failed()) return;
        try {
            f();
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu);
            if (!err) err = std::current_exception();
        }
    }
};

// Every lane group copies one row part: shard `me` owns ro[me][q] .. ro[me][q+1] of its own pairs; the part lands in the
// host result at sum_g ro[g][q] (the global offset of row q) + the lengths of the shards before `me`.
__global__ void __launch_bounds__(256) shard_scatter_kernel(const i64* __restrict__ ro_all, int nshards, int me, i64 npat,
                                                            const i64* __restrict__ pairs, i64* __restrict__ host_row_off,
                                                            i64* __restrict__ host_pairs) {
    const int lane = threadIdx.x & 31;
    const i64 q = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q > npat) return;
    i64 goff = 0, before = 0;
    for (int g = 0; g < nshards; ++g) {
        const i64 a = ro_all[(i64)g * (npat + 1) + q];
        goff += a;
        if (g < me && q < npat) before += ro_all[(i64)g * (npat + 1) + q + 1] - a;
    }
    if (me == 0 && lane == 0) host_row_off[q] = goff;
    if (q == npat) return;
    const i64 s = ro_all[(i64)me * (npat + 1) + q], e = ro_all[(i64)me * (npat + 1) + q + 1];
    const longlong2* src = reinterpret_cast<const longlong2*>(pairs) + s;
    longlong2* dst = reinterpret_cast<longlong2*>(host_pairs) + goff + before;
    for (i64 i = lane; i < e - s; i += 32) dst[i] = src[i];
}

ShardedIndex* sharded_create(const int32_t* devices, int32_t ndev, const cdb_options* opts) {
    std::unique_ptr<ShardedIndex> s(new ShardedIndex());
    if (opts)
        s->opt = *opts;
    else {
        s->opt.device = -1;
        s->opt.compat_signed = 1;
        s->opt.workspace_bytes = 0;
        s->opt.keep_host_copy = 0;
    }
    s->devs.assign(devices, devices + ndev);
    s->start_workers();
    return s.release();
}

void sharded_destroy(ShardedIndex* s) { delete s; }

int sharded_count(const ShardedIndex* s) { return (int)s->devs.size(); }

cdb_index* sharded_shard(const ShardedIndex* s, int g, i64* doc_begin, i64* doc_end) {
    if (!s->built || g < 0 || g >= (int)s->shard.size()) throw Error(CDB_ERR_STATE, "index has not been built");
    if (doc_begin) *doc_begin = s->doc_begin[g];
    if (doc_end) *doc_end = s->doc_begin[g + 1];
    return s->shard[g];
}

void sharded_add_many(ShardedIndex* s, const i64* ids, const u8* text, const i64* doc_off, i64 nd) {
    if (s->built) throw Error(CDB_ERR_STATE, "the sharded index has been built: add to a new one");
    const size_t t0 = s->text.size(), o0 = s->off.size(), i0 = s->ids.size();
    try {
        const i64 base = (i64)t0 - doc_off[0];
        if (doc_off[nd] > doc_off[0]) s->text.insert(s->text.end(), text + doc_off[0], text + doc_off[nd]);
        s->ids.insert(s->ids.end(), ids, ids + nd);
        s->off.reserve(o0 + (size_t)nd);
        for (i64 d = 1; d <= nd; ++d) s->off.push_back(base + doc_off[d]);
    } catch (...) {
        s->text.resize(t0);
        s->off.resize(o0);
        s->ids.resize(i0);
        throw;
    }
}

void sharded_build(ShardedIndex* s) {
    std::lock_guard<std::mutex> call(s->call_mu);
    if (s->built) throw Error(CDB_ERR_STATE, "the sharded index has been built already");
    const int D = (int)s->devs.size();
    const i64 nd = (i64)s->ids.size();
    const i64 total = (i64)s->text.size();
    // contiguous doc ranges of about equal bytes (a document is never split)
    s->doc_begin.assign(D + 1, nd);
    s->doc_begin[0] = 0;
    for (int g = 1; g < D; ++g) {
        const i64 want = total / D * g;
        const i64 d = (i64)(std::lower_bound(s->off.begin(), s->off.end(), want) - s->off.begin());
        s->doc_begin[g] = std::max(s->doc_begin[g - 1], std::min(d, nd));
    }
    s->shard.assign(D, nullptr);
    PhaseGuard guard;
    s->run_all([&](int g) {
        guard.run([&] {
            cdb_options o = s->opt;
            o.device = s->devs[g];
            cdb_index* ix = nullptr;
            if (cdb_create(&o, &ix) != CDB_OK) throw Error(CDB_ERR_NOMEM, cdb_last_error());
            s->shard[g] = ix;
            const i64 lo = s->doc_begin[g], hi = s->doc_begin[g + 1];
            if (hi > lo) {
                const cdb_status st = cdb_add_many(ix, s->ids.data() + lo, s->text.data(), s->off.data() + lo, hi - lo);
                if (st != CDB_OK) throw Error(st, cdb_last_error());
            }
            const cdb_status st = cdb_build(ix);
            if (st != CDB_OK) throw Error(st, cdb_last_error());
        });
    });
    if (guard.err) {
        for (cdb_index*& x : s->shard) {
            cdb_destroy(x);
            x = nullptr;
        }
        std::rethrow_exception(guard.err);
    }
    s->built = true;
    if (!s->opt.keep_host_copy) {
        std::vector<u8>().swap(s->text);
        std::vector<i64>().swap(s->off);
        s->off.push_back(0);
    }
}

// out_row_off / out_pairs: page-locked host buffers provided through `alloc` once the total is known
void sharded_locate(ShardedIndex* s, const u8* pat, const i64* pat_off, i64 npat,
                    const std::function<void(i64 total_pairs, i64** row_off, i64** pairs)>& alloc, i64* total_pairs_out,
                    i64* total_occ_out) {
    std::lock_guard<std::mutex> call(s->call_mu);
    if (!s->built) throw Error(CDB_ERR_STATE, "index has not been built");
    const int D = (int)s->devs.size();
    const i64 p0 = npat ? pat_off[0] : 0, pbytes = npat ? pat_off[npat] - p0 : 0;
    std::vector<i64> rel((size_t)npat + 1, 0);
    for (i64 q = 0; q <= npat && npat; ++q) rel[q] = pat_off[q] - p0;
    std::vector<cdb_device_result> res((size_t)D);
    for (auto& r : res) std::memset(&r, 0, sizeof(r));
    std::vector<u8*> d_pat((size_t)D, nullptr);
    std::vector<i64*> d_off((size_t)D, nullptr);
    std::vector<cudaStream_t> streams((size_t)D, nullptr);
    cudaEvent_t uploaded = nullptr;
    i64* h_row_off = nullptr;
    i64* h_pairs = nullptr;
    i64 total_pairs = 0, total_occ = 0;
    HostBarrier bar(D);
    PhaseGuard guard;
    s->run_all([&](int g) {
        const Index* ix = reinterpret_cast<const Index*>(s->shard[g]);
        const int dev = ix->device;
        cudaSetDevice(dev);
        cudaStream_t st = nullptr;
        i64* ro_all = nullptr;
        // ---- A: the batch reaches the first shard's device
        guard.run([&] {
            st = thread_ctx(dev).stream;
            streams[g] = st;
            CDB_CUDA(cudaMallocAsync((void**)&d_pat[g], (size_t)pbytes + 8, st));
            CDB_CUDA(cudaMallocAsync((void**)&d_off[g], (size_t)(npat + 1) * 8, st));
            if (g == 0) {
                if (pbytes) CDB_CUDA(cudaMemcpyAsync(d_pat[0], pat + p0, (size_t)pbytes, cudaMemcpyHostToDevice, st));
                CDB_CUDA(cudaMemcpyAsync(d_off[0], rel.data(), (size_t)(npat + 1) * 8, cudaMemcpyHostToDevice, st));
                CDB_CUDA(cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming));
                CDB_CUDA(cudaEventRecord(uploaded, st));
            }
        });
        bar.wait();
        // ---- B: device-to-device broadcast, then every shard locates the batch on its own documents
        guard.run([&] {
            if (g > 0) {
                const int dev0 = reinterpret_cast<const Index*>(s->shard[0])->device;
                CDB_CUDA(cudaStreamWaitEvent(st, uploaded, 0));
                if (pbytes) CDB_CUDA(cudaMemcpyPeerAsync(d_pat[g], dev, d_pat[0], dev0, (size_t)pbytes, st));
                CDB_CUDA(cudaMemcpyPeerAsync(d_off[g], dev, d_off[0], dev0, (size_t)(npat + 1) * 8, st));
            }
            if (npat) locate_device(*ix, d_pat[g], d_off[g], npat, st, &res[g]);
        });
        bar.wait();
        // ---- C: the result buffers, once the total is known
        if (g == 0) {
            guard.run([&] {
                for (int k = 0; k < D; ++k) {
                    total_pairs += res[k].total_pairs;
                    total_occ += res[k].total_occurrences;
                }
                alloc(total_pairs, &h_row_off, &h_pairs);
                if (npat == 0) h_row_off[0] = 0;
            });
        }
        bar.wait();
        // ---- D: row offsets of all shards come over device-to-device, every shard writes its part of every row
        guard.run([&] {
            if (npat == 0) return;
            CDB_CUDA(cudaMallocAsync((void**)&ro_all, (size_t)D * (npat + 1) * 8, st));
            for (int k = 0; k < D; ++k) {
                const int devk = reinterpret_cast<const Index*>(s->shard[k])->device;
                CDB_CUDA(cudaMemcpyPeerAsync(ro_all + (size_t)k * (npat + 1), dev, res[k].row_off, devk, (size_t)(npat + 1) * 8, st));
            }
            const i64 threads = (npat + 1) * 32;
            shard_scatter_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, st>>>(ro_all, D, g, npat, res[g].pairs, h_row_off, h_pairs);
            CDB_LAUNCH_CHECK();
            CDB_CUDA(cudaStreamSynchronize(st));
        });
        bar.wait();  // nobody frees a row_off array a peer is still copying
        if (st) {
            if (ro_all) cudaFreeAsync(ro_all, st);
            if (d_pat[g]) cudaFreeAsync(d_pat[g], st);
            if (d_off[g]) cudaFreeAsync(d_off[g], st);
            cdb_device_result_free(&res[g]);
            cudaStreamSynchronize(st);
        }
        if (g == 0 && uploaded) cudaEventDestroy(uploaded);
    });
    if (guard.err) std::rethrow_exception(guard.err);
    *total_pairs_out = total_pairs;
    *total_occ_out = total_occ;
}

}  // namespace cdb
