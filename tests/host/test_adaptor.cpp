// Host-side test of the C++ adaptor (coffeedb_b200/host/string_index.hpp).  Reads like the reference's own
// black-box tests: README.md:64-109 known answers, test/test-string.py's brute-force count property and
// test/test-highlight.py's marker splicing, at a small scale.
//   test_adaptor nogpu   -> only what works without a device: construction, add, and build() failing loudly
//   test_adaptor gpu     -> the full suite
// With -DCDB_TEST_REFERENCE_BASE=\"/root/reference/src/index.h\" the adaptor derives from the reference's own
// abstract `index` class (the drop-in arrangement of INTEGRATION.md).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <atomic>
#include <random>
#include <thread>
#include <stdexcept>
#include <string>
#include <string_view>
#ifdef CDB_TEST_REFERENCE_BASE
#include <format>
#include CDB_TEST_REFERENCE_BASE
using base_t = index;
#else
#include "../../coffeedb_b200/host/string_index.hpp"
using base_t = coffeedb_b200::index_base;
#endif
#include "../../coffeedb_b200/host/string_index.hpp"

using sindex = coffeedb_b200::basic_string_index<base_t>;
using result_t = std::vector<std::pair<int64_t, int64_t>>;

#define REQUIRE(c)                                                         \
    do {                                                                   \
        if (!(c)) {                                                        \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            std::exit(1);                                                  \
        }                                                                  \
    } while (0)

static int64_t brute(const std::string& doc, const std::string& kw) {  // test-string.py:14-19 (overlapping)
    int64_t c = 0;
    for (size_t p = doc.find(kw); p != std::string::npos; p = doc.find(kw, p + 1)) ++c;
    return c;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "gpu";
    if (mode == "nogpu") {
        std::unique_ptr<base_t> p = std::make_unique<sindex>();  // dynamic type probes of database.cpp:147,257,323
        REQUIRE(dynamic_cast<sindex*>(p.get()) != nullptr);
        static_cast<sindex*>(p.get())->add(1, "abc");
        if (cdb_device_count() == 0) {
            bool threw = false;
            try {
                p->build();
            } catch (const std::runtime_error& e) {
                threw = std::string(e.what()).find("no CPU fallback") != std::string::npos;
            }
            REQUIRE(threw);
        }
        std::puts("adaptor nogpu ok");
        return 0;
    }
    {  // README.md:64-109
        sindex ix;
        ix.add(100, "3010103");
        ix.add(101, "301022");
        ix.build();
        REQUIRE((ix.query("010") == result_t{{100, 2}, {101, 1}}));
        REQUIRE((ix.query("0") == result_t{{100, 3}, {101, 2}}));
        REQUIRE(ix.query("9").empty());
        bool threw = false;
        try {
            ix.query("");
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()) == "Empty keywords are not allowed";
        }
        REQUIRE(threw);
        auto sp = ix.spans({"010"}, {0});
        REQUIRE(sindex::render("3010103", sp[0], "<b>", "</b>") == "3<b>01010</b>3");  // README.md:109
    }
    {  // test-string.py at reduced scale, through the base-class pointer as database.cpp:387-393 calls it
        std::mt19937_64 rng(7);
        std::vector<std::string> docs(300);
        for (auto& d : docs) {
            d.resize(500);
            for (auto& c : d) c = (char)('a' + rng() % 26);
        }
        std::unique_ptr<base_t> p = std::make_unique<sindex>();
        for (size_t i = 0; i < docs.size(); ++i) static_cast<sindex*>(p.get())->add(1000 + (int64_t)i * 3, docs[i]);
        p->build();
        std::vector<std::string> kws;
        for (int t = 0; t < 50; ++t) {
            std::string kw(1 + t % 3, 'a');
            for (auto& c : kw) c = (char)('a' + rng() % 26);
            kws.push_back(kw);
        }
        auto rows = static_cast<const sindex*>(p.get())->query_batch(kws);
        std::vector<result_t> wants(kws.size());
        for (size_t t = 0; t < kws.size(); ++t) {
            result_t& want = wants[t];
            for (size_t i = 0; i < docs.size(); ++i)
                if (int64_t c = brute(docs[i], kws[t])) want.emplace_back(1000 + (int64_t)i * 3, c);
            REQUIRE(p->query(kws[t]) == want);
            REQUIRE(rows[t] == want);
        }
        // the server's calling pattern: a pool of worker threads, one keyword per call (httplib.h:97-101); concurrent
        // calls share device batches (cdb_query) and every caller still gets exactly its own row
        std::atomic<int> bad{0};
        std::vector<std::thread> workers;
        for (int w = 0; w < 8; ++w)
            workers.emplace_back([&, w] {
                for (int rep = 0; rep < 4; ++rep)
                    for (size_t t = 0; t < kws.size(); ++t) {
                        const size_t k = (t + (size_t)w * 7) % kws.size();
                        if (p->query(kws[k]) != wants[k]) ++bad;
                    }
            });
        for (auto& w : workers) w.join();
        REQUIRE(bad == 0);
        uint64_t nq = 0, nb = 0, largest = 0;
        REQUIRE(cdb_query_stats(static_cast<const sindex*>(p.get())->handle(), &nq, &nb, &largest) == CDB_OK);
        REQUIRE(nq == kws.size() + 8 * 4 * kws.size());
        REQUIRE(nb >= 1 && nb <= nq && largest >= 1 && largest <= 8);
        std::printf("coalesced queries: %llu in %llu batches (largest %llu)\n", (unsigned long long)nq,
                    (unsigned long long)nb, (unsigned long long)largest);
    }
    {  // $correlation composition (src/interface.cpp:79-146): OR of three keywords of one key, AND with a second key
        namespace cor = coffeedb_b200::correlation;
        std::mt19937_64 rng(11);
        std::vector<std::string> title(400), body(400);
        for (size_t i = 0; i < title.size(); ++i) {
            title[i].resize(40);
            body[i].resize(200);
            for (auto& c : title[i]) c = (char)('a' + rng() % 4);
            for (auto& c : body[i]) c = (char)('a' + rng() % 4);
        }
        sindex t, b;
        for (size_t i = 0; i < title.size(); ++i) {
            t.add(5000 - (int64_t)i, title[i]);  // ids descend: list order (doc index) != id order
            b.add(5000 - (int64_t)i, body[i]);
        }
        t.build();
        b.build();
        const std::vector<std::string> kws = {"abc", "ca", "dddd"};
        cor::list got = cor::and_merge(cor::or_of_rows(b.query_batch(kws)), cor::or_of_rows(t.query_batch({"ab"})));
        cor::list want;
        for (size_t i = title.size(); i-- > 0;) {  // ascending id
            int64_t cb = 0;
            for (auto& k : kws) cb += brute(body[i], k);
            const int64_t ct = brute(title[i], "ab");
            if (cb && ct) want.emplace_back(5000 - (int64_t)i, cb + ct);
        }
        REQUIRE(got == want);
        cor::finish(got, true, 3, 6);
        for (size_t i = 0; i < got.size(); ++i) {
            REQUIRE(got[i].second >= 3 && got[i].second < 6);
            REQUIRE(i == 0 || got[i - 1].second >= got[i].second);
        }
    }
    {  // the same surface over several shards (cdb_sharded_*): three shards on the devices of this box
        std::vector<int32_t> devs;
        for (int g = 0; g < 3; ++g) devs.push_back(g % cdb_device_count());
        coffeedb_b200::sharded_string_index sh(devs);
        sindex one;
        std::mt19937_64 rng(99);
        std::vector<std::string> docs(500);
        for (size_t i = 0; i < docs.size(); ++i) {
            docs[i].resize(20 + rng() % 200);
            for (auto& c : docs[i]) c = (char)('a' + rng() % 4);
            sh.add(9000 - (int64_t)i, docs[i]);
            one.add(9000 - (int64_t)i, docs[i]);
        }
        sh.build();
        one.build();
        for (const char* kw : {"a", "ab", "abcd", "dddd", "zz"}) REQUIRE(sh.query(kw) == one.query(kw));
        REQUIRE(sh.query_batch({"ab", "ca"}) == one.query_batch({"ab", "ca"}));
    }
    std::puts("adaptor gpu ok");
    return 0;
}
