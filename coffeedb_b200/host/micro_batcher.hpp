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
//   * callers that arrive meanwhile fill the next batch; followers wait until their batch is done and then read
//     their own row out of the shared result (no copy by the leader).  They wait WITHOUT the queue's mutex — a short
//     spin on the batch's flag, then the batch's own condition variable — and read the finished batch without any lock:
//     when 64 followers woke up through one mutex, handing it from one to the next cost more than the device batch.
//
// Header-only and CUDA-free: `Backend` is any callable
//     std::shared_ptr<Result> backend(const std::string& bytes, const std::vector<int64_t>& off)
// (keyword q = bytes[off[q], off[q+1]); throws on failure) whose Result exposes `const int64_t* row_off` ([n+1]) and
// `const int64_t* pairs` ((id, count) pairs, as cdb_result).  capi.cu instantiates it over the locate path
// (cdb_query); tests/host/test_batcher.cpp drives it with a host-only stand-in.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
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
        idle_cv_.wait(lk, [&] { return active_.load(std::memory_order_acquire) == 0; });
    }

    // string_index::query(keyword) for one keyword (src/index.cpp:237-326), coalesced with concurrent callers.
    row_view query(std::string_view keyword) {
        // the reference rejects an empty keyword before touching the index (src/index.cpp:239-241); doing it here keeps
        // one bad request from failing the strangers that share its batch
        if (keyword.empty()) throw std::runtime_error("Empty keywords are not allowed");
        active_.fetch_add(1, std::memory_order_acq_rel);
        leave_guard leave{this};
        std::unique_lock<std::mutex> lk(mu_);
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
            slot_cv_.notify_all();
            lk.unlock();
            {
                std::lock_guard<std::mutex> bl(b->mu);  // pairs with the followers' predicate check: no lost wake-up
                b->done.store(true, std::memory_order_release);
            }
            b->done_cv.notify_all();
        } else {
            lk.unlock();
            // the batch is on the device for a few tens of microseconds: look at its flag for about that long before sleeping
            const auto t_give_up = std::chrono::steady_clock::now() + std::chrono::microseconds(80);
            for (int spin = 0; !b->done.load(std::memory_order_acquire); ++spin) {
                if ((spin & 31) == 31) {
                    if (std::chrono::steady_clock::now() > t_give_up) break;
                    std::this_thread::yield();
                }
            }
            if (!b->done.load(std::memory_order_acquire)) {
                std::unique_lock<std::mutex> bl(b->mu);
                b->done_cv.wait(bl, [&] { return b->done.load(std::memory_order_acquire); });
            }
        }
        // from here on the batch is immutable: read without a lock
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
        std::atomic<bool> done{false};
        std::shared_ptr<Result> result;
        std::exception_ptr error;
        std::condition_variable cv;       // leader (with linger): full — under the queue's mutex
        std::mutex mu;                    // followers that gave up spinning sleep here, not on the queue's mutex
        std::condition_variable done_cv;
    };
    struct leave_guard {  // --active_ on every way out of query(); the destructor of the queue waits for 0
        micro_batcher* self;
        ~leave_guard() {
            if (self->active_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> lk(self->mu_);
                self->idle_cv_.notify_all();
            }
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
    std::atomic<int> active_{0};
    micro_batcher_stats stats_;
};

}  // namespace coffeedb_b200
