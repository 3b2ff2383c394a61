// Host-only test of coffeedb_b200::micro_batcher (coffeedb_b200/host/micro_batcher.hpp) with a stand-in backend: no
// CUDA, no library.  Checks that every caller gets exactly its own row, that concurrent callers are coalesced, the
// batch-size and in-flight limits, the linger window, and error delivery (the reference's empty-keyword message for
// the offender only; a backend failure for every member of the failed batch and nobody else).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>

#include "../../coffeedb_b200/host/micro_batcher.hpp"

using namespace coffeedb_b200;

#define REQUIRE(c)                                                         \
    do {                                                                   \
        if (!(c)) {                                                        \
            std::fprintf(stderr, "%s:%d: REQUIRE(%s)\n", __FILE__, __LINE__, #c); \
            std::exit(1);                                                  \
        }                                                                  \
    } while (0)

struct fake_result {
    std::vector<int64_t> ro, pr;
    const int64_t* row_off = nullptr;
    const int64_t* pairs = nullptr;
};

static uint64_t hash_of(std::string_view s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    return h;
}

// row of keyword kw: (hash + i, i + 1) for i < len(kw) % 7
static std::vector<std::pair<int64_t, int64_t>> expected_row(std::string_view kw) {
    std::vector<std::pair<int64_t, int64_t>> r;
    const uint64_t h = hash_of(kw) >> 8;
    for (size_t i = 0; i < kw.size() % 7; ++i) r.emplace_back((int64_t)(h + i), (int64_t)i + 1);
    return r;
}

struct fake_backend {
    std::atomic<int>* running;
    std::atomic<int>* max_running;
    std::atomic<uint64_t>* largest;
    int sleep_us;
    std::shared_ptr<fake_result> operator()(const std::string& bytes, const std::vector<int64_t>& off) const {
        const int now = ++*running;
        int seen = max_running->load();
        while (now > seen && !max_running->compare_exchange_weak(seen, now)) {
        }
        const uint64_t n = off.size() - 1;
        uint64_t big = largest->load();
        while (n > big && !largest->compare_exchange_weak(big, n)) {
        }
        std::this_thread::sleep_for(std::chrono::microseconds(sleep_us));
        auto res = std::make_shared<fake_result>();
        res->ro.push_back(0);
        bool boom = false;
        for (size_t q = 0; q + 1 < off.size(); ++q) {
            std::string_view kw(bytes.data() + off[q], (size_t)(off[q + 1] - off[q]));
            if (kw.find("BOOM") != std::string_view::npos) boom = true;
            for (auto& p : expected_row(kw)) {
                res->pr.push_back(p.first);
                res->pr.push_back(p.second);
            }
            res->ro.push_back((int64_t)res->pr.size() / 2);
        }
        res->row_off = res->ro.data();
        res->pairs = res->pr.data();
        --*running;
        if (boom) throw std::runtime_error("device failure");
        return res;
    }
};

using batcher = micro_batcher<fake_backend, fake_result>;

static bool same(const batcher::row_view& v, std::string_view kw) {
    auto e = expected_row(kw);
    if ((size_t)v.count != e.size()) return false;
    for (size_t i = 0; i < e.size(); ++i)
        if (v.pairs[2 * i] != e[i].first || v.pairs[2 * i + 1] != e[i].second) return false;
    return true;
}

static std::string random_keyword(std::mt19937_64& rng) {
    std::string s(1 + rng() % 12, 'a');
    for (auto& c : s) c = (char)('a' + rng() % 26);
    return s;
}

int main() {
    std::atomic<int> running{0}, max_running{0};
    std::atomic<uint64_t> largest{0};

    {  // 1. many threads, one batch in flight: every row is the caller's own, and batches form
        batcher mb(fake_backend{&running, &max_running, &largest, 200}, 1 << 16, 1);
        const int T = 16, Q = 400;
        std::atomic<int> bad{0};
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                std::mt19937_64 rng(1000 + t);
                for (int i = 0; i < Q; ++i) {
                    const std::string kw = random_keyword(rng);
                    if (!same(mb.query(kw), kw)) ++bad;
                }
            });
        for (auto& x : th) x.join();
        const auto st = mb.statistics();
        REQUIRE(bad == 0);
        REQUIRE(st.queries == (uint64_t)T * Q);
        REQUIRE(st.batches < st.queries / 2);  // 16 callers against a 200 us backend: far fewer calls than queries
        REQUIRE(st.largest > 1 && st.largest <= (uint64_t)T);
        REQUIRE(max_running == 1);
        std::printf("coalescing: %llu queries in %llu batches (largest %llu)\n", (unsigned long long)st.queries,
                    (unsigned long long)st.batches, (unsigned long long)st.largest);
    }
    {  // 2. a single caller is never delayed or grouped
        batcher mb(fake_backend{&running, &max_running, &largest, 0});
        for (int i = 0; i < 100; ++i) REQUIRE(same(mb.query("single" + std::to_string(i)), "single" + std::to_string(i)));
        REQUIRE(mb.statistics().batches == 100 && mb.statistics().largest == 1);
    }
    {  // 3. batch-size and in-flight limits
        max_running = 0;
        largest = 0;
        batcher mb(fake_backend{&running, &max_running, &largest, 300}, 4, 2);
        std::atomic<int> bad{0};
        std::vector<std::thread> th;
        for (int t = 0; t < 24; ++t)
            th.emplace_back([&, t] {
                std::mt19937_64 rng(2000 + t);
                for (int i = 0; i < 100; ++i) {
                    const std::string kw = random_keyword(rng);
                    if (!same(mb.query(kw), kw)) ++bad;
                }
            });
        for (auto& x : th) x.join();
        REQUIRE(bad == 0);
        REQUIRE(largest <= 4 && mb.statistics().largest <= 4);
        REQUIRE(max_running <= 2);
    }
    {  // 4. errors: the empty keyword fails alone and at once; a backend failure reaches the whole batch, later ones work
        batcher mb(fake_backend{&running, &max_running, &largest, 2000});
        try {
            mb.query("");
            REQUIRE(!"empty keyword accepted");
        } catch (const std::runtime_error& e) {
            REQUIRE(std::string(e.what()) == "Empty keywords are not allowed");
        }
        REQUIRE(mb.statistics().queries == 0);
        std::atomic<int> failed{0}, ok{0};
        std::vector<std::thread> th;
        // the first caller occupies the device for 2 ms; the other seven (one of them poisonous) share the next batch
        th.emplace_back([&] { REQUIRE(same(mb.query("first"), "first")); });
        while (mb.statistics().batches == 0) std::this_thread::yield();  // "first" is on the device, alone
        for (int t = 0; t < 7; ++t)
            th.emplace_back([&, t] {
                const std::string kw = t == 3 ? "kaBOOM" : "kw" + std::to_string(t);
                try {
                    REQUIRE(same(mb.query(kw), kw));
                    ++ok;
                } catch (const std::runtime_error& e) {
                    REQUIRE(std::string(e.what()) == "device failure");
                    ++failed;
                }
            });
        for (auto& x : th) x.join();
        REQUIRE(failed + ok == 7 && failed >= 1);  // at least the offender; its batch mates with it
        REQUIRE(same(mb.query("after"), "after"));
    }
    {  // 5. linger: callers arriving within the window share one batch even with an idle device
        batcher mb(fake_backend{&running, &max_running, &largest, 0}, 1 << 16, 1, std::chrono::milliseconds(200));
        std::vector<std::thread> th;
        for (int t = 0; t < 4; ++t) th.emplace_back([&, t] { REQUIRE(same(mb.query("linger" + std::to_string(t)), "linger" + std::to_string(t))); });
        for (auto& x : th) x.join();
        REQUIRE(mb.statistics().batches <= 2);
    }
    {  // 6. a full batch goes at once, without waiting out the linger window
        batcher mb(fake_backend{&running, &max_running, &largest, 0}, 2, 1, std::chrono::seconds(30));
        std::thread a([&] { REQUIRE(same(mb.query("aa"), "aa")); });
        std::thread b([&] { REQUIRE(same(mb.query("bbb"), "bbb")); });
        a.join();
        b.join();
        REQUIRE(mb.statistics().batches == 1 && mb.statistics().largest == 2);
    }
    std::printf("micro_batcher ok\n");
    return 0;
}
