// The reference server's calling pattern, without the server: T worker threads (httplib's pool is max(8, hw-1),
// httplib.h:97-101), each calling string_index::query(keyword) — one keyword per call (src/database.cpp:387-393) —
// against one index.  Prints queries/s and how many device batches the coalescing queue (cdb_query) issued.
//   g++ -std=c++20 -O2 -pthread tools/query_pool_bench.cpp -o tools/_build/query_pool_bench \
//       -Lcoffeedb_b200 -lcoffeedb_b200 -Wl,-rpath,$PWD/coffeedb_b200
//   tools/_build/query_pool_bench [docs=100000] [doc_bytes=100] [keyword_bytes=4] [queries_per_thread=2000]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>

#include "../coffeedb_b200/host/string_index.hpp"

int main(int argc, char** argv) {
    const int64_t nd = argc > 1 ? atoll(argv[1]) : 100000;
    const int doclen = argc > 2 ? atoi(argv[2]) : 100;
    const int m = argc > 3 ? atoi(argv[3]) : 4;
    const int per_thread = argc > 4 ? atoi(argv[4]) : 2000;
    std::mt19937_64 rng(12345);
    coffeedb_b200::string_index ix;
    {
        std::string doc((size_t)doclen, 'a');
        for (int64_t d = 0; d < nd; ++d) {
            for (auto& c : doc) c = (char)('a' + rng() % 26);
            ix.add(1000 + d, doc);
        }
    }
    const auto tb = std::chrono::steady_clock::now();
    ix.build();
    std::printf("index: %lld docs x %d B, built in %.1f ms\n", (long long)nd, doclen,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb).count());
    auto keyword = [m](std::mt19937_64& r) {
        std::string kw((size_t)m, 'a');
        for (auto& c : kw) c = (char)('a' + r() % 26);
        return kw;
    };
    for (int i = 0; i < 50; ++i) ix.query(keyword(rng));  // warm-up
    for (int T : {1, 8, 16, 32, 64, 128}) {
        uint64_t q0 = 0, b0 = 0, l0 = 0, q1 = 0, b1 = 0, l1 = 0;
        cdb_query_stats(ix.handle(), &q0, &b0, &l0);
        std::atomic<int64_t> pairs{0};
        std::atomic<int> warm{0};
        std::atomic<bool> go{false};
        std::vector<std::thread> th;
        // a server's worker threads are long-lived: every thread first issues a few untimed calls (its stream, its first
        // turn as a batch leader), then all of them start the timed part together
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                std::mt19937_64 r(777 + t);
                for (int i = 0; i < 20; ++i) ix.query(keyword(r));
                warm.fetch_add(1);
                while (!go.load()) std::this_thread::yield();
                int64_t mine = 0;
                for (int i = 0; i < per_thread; ++i) mine += (int64_t)ix.query(keyword(r)).size();
                pairs += mine;
            });
        while (warm.load() < T) std::this_thread::yield();
        cdb_query_stats(ix.handle(), &q0, &b0, &l0);
        const auto t0 = std::chrono::steady_clock::now();
        go.store(true);
        for (auto& x : th) x.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        cdb_query_stats(ix.handle(), &q1, &b1, &l1);
        std::printf("%3d threads: %9.0f queries/s  (%llu queries in %llu device batches, largest %llu; %.1f pairs per query)\n", T,
                    (double)T * per_thread / s, (unsigned long long)(q1 - q0), (unsigned long long)(b1 - b0),
                    (unsigned long long)l1, (double)pairs / ((double)T * per_thread));
    }
    return 0;
}
