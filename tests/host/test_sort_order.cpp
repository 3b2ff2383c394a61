// coffeedb_b200/host/std_sort_order.hpp against the real std::sort: the permutation of (id, $correlation) pairs that
// `std::sort(answer.begin(), answer.end(), [](auto x, auto y) { return x.second > y.second; })` (src/interface.cpp:143-146)
// produces on an id-ascending input must equal the permutation of handles the restatement produces.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <random>
#include <utility>
#include <vector>

#include "../../coffeedb_b200/host/std_sort_order.hpp"

using Pair = std::pair<int64_t, int64_t>;

static long checked = 0;

static bool check(const std::vector<int64_t>& corr, const char* what) {
    const int n = (int)corr.size();
    std::vector<Pair> v(n);
    for (int i = 0; i < n; ++i) v[i] = {1000 + 3 * (int64_t)i, corr[i]};  // ids ascending, as filter() hands them over
    std::sort(v.begin(), v.end(), [](auto x, auto y) { return x.second > y.second; });
    std::vector<uint32_t> p(n);
    std::iota(p.begin(), p.end(), 0u);
    coffeedb_b200::sort_order::std_sort_order(p.data(), n, [&](uint32_t a, uint32_t b) { return corr[a] > corr[b]; });
    for (int i = 0; i < n; ++i) {
        if (v[i].first != 1000 + 3 * (int64_t)p[i]) {
            std::printf("MISMATCH %s n=%d at %d: std::sort has id %lld, restatement %lld\n", what, n, i, (long long)v[i].first,
                        (long long)(1000 + 3 * (int64_t)p[i]));
            return false;
        }
    }
    ++checked;
    return true;
}

// input on which median-of-three quicksort degenerates (Musser's construction): drives introsort into its heap-sort fallback
static std::vector<int64_t> killer(int n) {
    std::vector<int64_t> a(n);
    const int k = n / 2;
    for (int i = 1; i <= k; ++i) {
        if (i % 2 == 1) {
            a[i - 1] = i;
            a[i] = k + i;
        }
        a[k + i - 1] = 2 * i;
    }
    return a;
}

int main() {
    std::mt19937_64 rng(20260101);
    bool ok = true;
    // every length up to 1100 with all-equal keys (the table of the device path), and with 0/1 keys
    for (int n = 0; n <= 1100 && ok; ++n) {
        ok = ok && check(std::vector<int64_t>(n, 1), "constant");
        std::vector<int64_t> c(n);
        for (auto& x : c) x = 1 + (int64_t)(rng() % 2);
        ok = ok && check(c, "two values");
    }
    // few-valued (the realistic case: counts 1, sometimes 2 or 3), at many lengths
    for (int rep = 0; rep < 4000 && ok; ++rep) {
        const int n = (int)(rng() % 1500);
        std::vector<int64_t> c(n);
        const int mode = rep % 5;
        for (auto& x : c) {
            const uint64_t r = rng();
            if (mode == 0) x = 1 + (r % 100 == 0);
            else if (mode == 1) x = 1 + (int64_t)(r % 3);
            else if (mode == 2) x = (int64_t)(r % 1000);
            else if (mode == 3) x = (int64_t)(r % 7) - 3;
            else x = (int64_t)r;
        }
        ok = ok && check(c, "random");
    }
    // sorted, reversed, organ pipe, killer sequences (heap-sort fallback)
    for (int n : {17, 33, 100, 257, 1000, 4096, 20000, 100000}) {
        std::vector<int64_t> c(n);
        std::iota(c.begin(), c.end(), 0);
        ok = ok && check(c, "ascending");
        std::reverse(c.begin(), c.end());
        ok = ok && check(c, "descending");
        for (int i = 0; i < n; ++i) c[i] = std::min(i, n - 1 - i);
        ok = ok && check(c, "organ pipe");
        ok = ok && check(killer(n), "killer");
        auto neg = killer(n);
        for (auto& x : neg) x = -x;
        ok = ok && check(neg, "killer negated");
    }
    if (!ok) return 1;
    std::printf("sort_order ok: %ld sequences\n", checked);
    return 0;
}
