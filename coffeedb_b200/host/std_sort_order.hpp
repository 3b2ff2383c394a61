// The permutation std::sort produces — restated so that it can run inside a CUDA kernel.
//
// filter() ends with `std::sort(answer.begin(), answer.end(), [](auto x, auto y) { return x.second > y.second; })`
// (src/interface.cpp:143-146).  The sort is unstable and most $correlation values tie (a keyword rarely occurs twice in
// a document), so WHICH objects a `span` (src/interface.cpp:196-209) returns is decided by the permutation libstdc++'s
// introsort happens to apply to the id-ascending input.  A device-side filter that applies `span` before the result
// leaves the GPU therefore has to reproduce that permutation exactly.  The algorithm below is the published introsort
// of libstdc++ (bits/stl_algo.h of GCC 13, the toolchain the reference is built with here: median-of-three pivot moved
// to the front, unguarded Hoare partition, recursion on the right part / iteration on the left, depth limit 2*floor(lg n)
// with a heap sort beyond it, threshold 16, final insertion sort), expressed over an array of element handles `p[0..n)`
// and a strict-weak `less(a, b)` on handles.  It only depends on the outcome of comparisons, so it yields the same
// permutation as std::sort on the elements themselves — tests/host/test_sort_order.cpp checks that against the real
// std::sort on random, constant, few-valued, sorted, reversed and median-of-three-killer inputs.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CDB_HD __host__ __device__
#else
#define CDB_HD
#endif

namespace coffeedb_b200 {
namespace sort_order {

template <typename H>
CDB_HD inline void hswap(H& a, H& b) {
    const H t = a;
    a = b;
    b = t;
}

// ---- heap sort of p[first, last) (the depth-limit fallback: partial_sort(first, last, last)) ------------------------------
template <typename H, typename Less>
CDB_HD inline void push_heap_(H* p, int first, int hole, int top, H value, Less less) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(p[first + parent], value)) {
        p[first + hole] = p[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    p[first + hole] = value;
}

template <typename H, typename Less>
CDB_HD inline void adjust_heap_(H* p, int first, int hole, int len, H value, Less less) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(p[first + child], p[first + child - 1])) --child;
        p[first + hole] = p[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        p[first + hole] = p[first + child - 1];
        hole = child - 1;
    }
    push_heap_(p, first, hole, top, value, less);
}

template <typename H, typename Less>
CDB_HD inline void heap_sort_(H* p, int first, int last, Less less) {
    const int len = last - first;
    if (len >= 2) {  // make_heap
        for (int parent = (len - 2) / 2;; --parent) {
            const H value = p[first + parent];
            adjust_heap_(p, first, parent, len, value, less);
            if (parent == 0) break;
        }
    }
    for (int end = last; end - first > 1;) {  // sort_heap: pop the maximum to the back
        --end;
        const H value = p[end];
        p[end] = p[first];
        adjust_heap_(p, first, 0, end - first, value, less);
    }
}

// ---- insertion sorts ------------------------------------------------------------------------------------------------------
template <typename H, typename Less>
CDB_HD inline void unguarded_linear_insert_(H* p, int last, Less less) {
    const H val = p[last];
    int next = last - 1;
    while (less(val, p[next])) {
        p[last] = p[next];
        last = next;
        --next;
    }
    p[last] = val;
}

template <typename H, typename Less>
CDB_HD inline void insertion_sort_(H* p, int first, int last, Less less) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(p[i], p[first])) {
            const H val = p[i];
            for (int j = i; j > first; --j) p[j] = p[j - 1];  // move_backward(first, i, i + 1)
            p[first] = val;
        } else {
            unguarded_linear_insert_(p, i, less);
        }
    }
}

// ---- the sort ---------------------------------------------------------------------------------------------------------------
// Reorders the handles p[0..n) the way std::sort(first, last, comp) reorders its elements, where less(a, b) == comp(*a, *b).
template <typename H, typename Less>
CDB_HD inline void std_sort_order(H* p, int n, Less less) {
    if (n <= 0) return;
    constexpr int kThreshold = 16;
    int lg = 0;
    while ((n >> (lg + 1)) != 0) ++lg;  // floor(log2 n)
    // introsort loop with an explicit stack: the right part of every partition is pushed (the recursive call), the loop
    // continues on the left part; the parts are disjoint, so the order in which they are processed does not matter
    struct Frame {
        int first, last, depth;
    };
    Frame stack[72];
    int sp = 0;
    stack[sp++] = Frame{0, n, 2 * lg};
    while (sp > 0) {
        Frame f = stack[--sp];
        int first = f.first, last = f.last, depth = f.depth;
        while (last - first > kThreshold) {
            if (depth == 0) {
                heap_sort_(p, first, last, less);
                break;
            }
            --depth;
            // median of p[first+1], p[mid], p[last-1] goes to p[first]
            const int mid = first + (last - first) / 2;
            const int a = first + 1, b = mid, c = last - 1;
            if (less(p[a], p[b])) {
                if (less(p[b], p[c])) hswap(p[first], p[b]);
                else if (less(p[a], p[c])) hswap(p[first], p[c]);
                else hswap(p[first], p[a]);
            } else if (less(p[a], p[c])) {
                hswap(p[first], p[a]);
            } else if (less(p[b], p[c])) {
                hswap(p[first], p[c]);
            } else {
                hswap(p[first], p[b]);
            }
            // unguarded partition of [first+1, last) around the pivot p[first]
            int lo = first + 1, hi = last;
            for (;;) {
                while (less(p[lo], p[first])) ++lo;
                --hi;
                while (less(p[first], p[hi])) --hi;
                if (!(lo < hi)) break;
                hswap(p[lo], p[hi]);
                ++lo;
            }
            stack[sp++] = Frame{lo, last, depth};  // __introsort_loop(cut, last, depth_limit)
            last = lo;
        }
    }
    // final insertion sort
    if (n > kThreshold) {
        insertion_sort_(p, 0, kThreshold, less);
        for (int i = kThreshold; i != n; ++i) unguarded_linear_insert_(p, i, less);
    } else {
        insertion_sort_(p, 0, n, less);
    }
}

}  // namespace sort_order
}  // namespace coffeedb_b200
