// coffeedb_b200::micro_batcher — coalesces concurrent single-keyword queries into device batches.
//
// The reference only ever issues batches of one: `string_index::query(keyword)` is called once per keyword, from up to
// max(8, hw-1) httplib worker threads at a time (src/database.cpp:387-393 under a shared lock, src/interface.cpp:79-86
// for the keyword list of one key).  On the GPU one keyword costs a full launch sequence (~150 us) while 10^6 keywords
// cost ~10 ms, so the drop-in `query()` groups whatever is waiting — the "group commit" scheme of database logs:
//
//   * a caller appends its keyword to the open batch; the first one in becomes the batch's leader;
//   * the leader waits until fewer than `max_in_flight` batches are on the device (no timer: with an idle device it
//     goes at once, so a lone query pays no extra latency), closes the batch and runs it through the backend;
//   * callers that arrive meanwhile fill the next batch; followers sleep until their batch is done and then read
//     their own row out of the shared result (no copy by the leader).
//
// Header-only and CUDA-free: `Backend` is any callable
//     std::shared_ptr<Result> backend(const std::string& bytes, const std::vector<int64_t>& off)
// (keyword q = bytes[off[q], off[q+1]); throws on failure) whose Result exposes `const int64_t* row_off` ([n+1]) and
// `const int64_t* pairs` ((id, count) pairs, as cdb_result).  capi.cu instantiates it over the locate path
// (cdb_query); tests/host/test_batcher.cpp drives it with a host-only stand-in.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace coffeedb_b200 {

struct micro_batcher_stats {
    uint64_t queries = 0;   // keywords submitted
    uint64_t batches = 0;   // backend calls
    uint64_t largest = 0;   // keywords in the largest batch
};

template <class Backend, class Result>
class micro_batcher {
public:
    // One keyword's answer: a view of its row inside the batch result, which it keeps alive.
    struct row_view {
        std::shared_ptr<Result> owner;
        const int64_t* pairs = nullptr;  // 2 * count int64: (id, count), ascending doc index
        int64_t count = 0;
    };

    explicit micro_batcher(Backend backend, size_t max_batch = 1 << 16, int max_in_flight = 1,
                           std::chrono::microseconds linger = std::chrono::microseconds(0))
        : backend_(std::move(backend)),
          max_batch_(max_batch ? max_batch : 1),
          max_in_flight_(max_in_flight > 0 ? max_in_flight : 1),
          linger_(linger) {}

    micro_batcher(const micro_batcher&) = delete;
    micro_batcher& operator=(const micro_batcher&) = delete;

    // Blocks until no caller is inside query(); the owner must not start new queries while destroying.
    ~micro_batcher() {
        std::unique_lock<std::mutex> lk(mu_);
        idle_cv_.wait(lk, [&] { return active_ == 0; });
    }

    // string_index::query(keyword) for one keyword (src/index.cpp:237-326), coalesced with concurrent callers.
    row_view query(std::string_view keyword) {
        // the reference rejects an empty keyword before touching the index (src/index.cpp:239-241); doing it here keeps
        // one bad request from failing the strangers that share its batch
        if (keyword.empty()) throw std::runtime_error("Empty keywords are not allowed");
        std::unique_lock<std::mutex> lk(mu_);
        ++active_;
        leave_guard leave{this, &lk};
        if (!open_) open_ = std::make_shared<batch>();
        std::shared_ptr<batch> b = open_;
        // strong guarantee: the keyword's bytes and offset go in first (either may throw bad_alloc, and a slot without
        // an offset would shift every later member's row); the slot is counted last
        b->off.reserve(b->off.size() + 1);
        b->bytes.append(keyword.data(), keyword.size());
        b->off.push_back((int64_t)b->bytes.size());  // cannot throw: reserved above
        const size_t mine = b->n++;
        ++stats_.queries;
        if (b->n >= max_batch_) {
            open_.reset();  // full: later arrivals start the next batch
            b->cv.notify_all();
        }
        if (mine == 0) {
            slot_cv_.wait(lk, [&] { return in_flight_ < max_in_flight_; });
            if (linger_.count() > 0 && b->n < max_batch_) {
                b->cv.wait_for(lk, linger_, [&] { return b->n >= max_batch_; });
                // the lock was dropped while lingering: another leader may have taken the slot
                slot_cv_.wait(lk, [&] { return in_flight_ < max_in_flight_; });
            }
            if (open_ == b) open_.reset();  // close: nobody joins from here on
            ++in_flight_;
            ++stats_.batches;
            if (b->n > stats_.largest) stats_.largest = b->n;
            lk.unlock();
            try {
                b->result = backend_(b->bytes, b->off);
                if (!b->result) throw std::runtime_error("micro_batcher: backend returned no result");
            } catch (...) {
                b->error = std::current_exception();
            }
            lk.lock();
            --in_flight_;
            b->done = true;
            slot_cv_.notify_all();
            b->cv.notify_all();
        } else {
            b->cv.wait(lk, [&] { return b->done; });
        }
        if (b->error) std::rethrow_exception(b->error);
        row_view v;
        v.owner = b->result;
        const int64_t lo = b->result->row_off[mine], hi = b->result->row_off[mine + 1];
        v.pairs = b->result->pairs + 2 * lo;
        v.count = hi - lo;
        return v;
    }

    micro_batcher_stats statistics() const {
        std::lock_guard<std::mutex> lk(mu_);
        return stats_;
    }

private:
    struct batch {
        std::string bytes;
        std::vector<int64_t> off{0};
        size_t n = 0;
        bool done = false;
        std::shared_ptr<Result> result;
        std::exception_ptr error;
        std::condition_variable cv;  // followers: done; leader (with linger): full
    };
    struct leave_guard {  // --active_ under the lock on every way out of query()
        micro_batcher* self;
        std::unique_lock<std::mutex>* lk;
        ~leave_guard() {
            if (!lk->owns_lock()) lk->lock();
            if (--self->active_ == 0) self->idle_cv_.notify_all();
        }
    };

    Backend backend_;
    const size_t max_batch_;
    const int max_in_flight_;
    const std::chrono::microseconds linger_;
    mutable std::mutex mu_;
    std::condition_variable slot_cv_, idle_cv_;
    std::shared_ptr<batch> open_;
    int in_flight_ = 0;
    int active_ = 0;
    micro_batcher_stats stats_;
};

}  // namespace coffeedb_b200
