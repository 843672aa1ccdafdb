// gnu_sort.h -- the order libstdc++'s std::sort leaves an index array in, restated for host AND device code.
//
// The reference sorts bokeh rows and columns with std::sort and a "greater by value" index comparator
// (src/zoic.cpp:317, :381).  Where values tie (the zero pixels outside the aperture shape, or equal pixels of a
// real photograph) the resulting order is whatever the library's introsort does for that input: not specified by
// the language, but fully determined by the algorithm.  libstdc++ is not vendored in /root/reference (it is the
// toolchain's: GCC 13.3, bits/stl_algo.h + bits/stl_heap.h); this file restates its published algorithm --
// introsort: median-of-three quicksort down to 16-element runs, heapsort once 2*floor(log2 n) partitions deep,
// one final insertion sort -- so that the tables built on the GPU (bokeh_build.cu) equal the
// tables of the compiled reference entry for entry, ties included.  tests/test_gnu_sort.py checks it against the
// toolchain's own std::sort on random, tie-heavy and quicksort-adversarial inputs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZSORT_HD __host__ __device__ inline
#else
#define ZSORT_HD inline
#endif

namespace zoicb {
namespace gnusort {

// "a comes before b": index a's value is greater (descending by value), as the reference's comparator
template <typename Index>
struct Before {
    const float* v;
    ZSORT_HD bool operator()(Index a, Index b) const { return v[a] > v[b]; }
};

template <typename Index, typename Cmp>
ZSORT_HD void push_heap_value(Index* first, long hole, long top, Index value, Cmp before) {
    long parent = (hole - 1) / 2;
    while (hole > top && before(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <typename Index, typename Cmp>
ZSORT_HD void adjust_heap(Index* first, long hole, long len, Index value, Cmp before) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (before(first[child], first[child - 1])) --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_value(first, hole, top, value, before);
}

template <typename Index, typename Cmp>
ZSORT_HD void heap_sort(Index* first, long len, Cmp before) {
    // partial_sort(first, last, last): make_heap, (empty select phase), sort_heap
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            Index value = first[parent];
            adjust_heap(first, parent, len, value, before);
            if (parent == 0) break;
            --parent;
        }
    }
    long last = len;
    while (last > 1) {
        --last;
        Index value = first[last];
        first[last] = first[0];
        adjust_heap(first, 0, last, value, before);
    }
}

template <typename Index, typename Cmp>
ZSORT_HD void unguarded_linear_insert(Index* last, Cmp before) {
    Index value = *last;
    Index* next = last - 1;
    while (before(value, *next)) {
        *last = *next;
        last = next;
        --next;
    }
    *last = value;
}

template <typename Index, typename Cmp>
ZSORT_HD void insertion_sort(Index* first, Index* last, Cmp before) {
    if (first == last) return;
    for (Index* i = first + 1; i != last; ++i) {
        if (before(*i, *first)) {
            Index value = *i;
            for (Index* p = i; p != first; --p) *p = *(p - 1);
            *first = value;
        } else {
            unguarded_linear_insert(i, before);
        }
    }
}

template <typename Index, typename Cmp>
ZSORT_HD Index* partition_pivot(Index* first, Index* last, Cmp before) {
    Index* mid = first + (last - first) / 2;
    Index *a = first + 1, *b = mid, *c = last - 1, *m;
    // median of (a, b, c) goes to *first
    if (before(*a, *b)) {
        if (before(*b, *c)) m = b;
        else if (before(*a, *c)) m = c;
        else m = a;
    } else if (before(*a, *c)) m = a;
    else if (before(*b, *c)) m = c;
    else m = b;
    { Index t = *first; *first = *m; *m = t; }
    Index* lo = first + 1;
    Index* hi = last;
    for (;;) {
        while (before(*lo, *first)) ++lo;
        --hi;
        while (before(*first, *hi)) --hi;
        if (!(lo < hi)) return lo;
        Index t = *lo; *lo = *hi; *hi = t;
        ++lo;
    }
}

// std::sort(first, first + n, before)
template <typename Index, typename Cmp>
ZSORT_HD void sort(Index* first, long n, Cmp before) {
    if (n <= 0) return;
    constexpr long kThreshold = 16;
    // introsort loop; the recursion on the right part is replaced by an explicit stack (device code):
    // every entry is a right part still to do, at most one per partition depth
    struct Todo { Index* first; Index* last; int depth; };
    Todo stack[72];
    int sp = 0;
    int lg = 0;
    for (long m = n; m > 1; m >>= 1) ++lg;
    Index* lo = first;
    Index* hi = first + n;
    int depth = 2 * lg;
    for (;;) {
        while (hi - lo > kThreshold) {
            if (depth == 0) {
                heap_sort(lo, (long)(hi - lo), before);
                break;
            }
            --depth;
            Index* cut = partition_pivot(lo, hi, before);
            stack[sp].first = cut; stack[sp].last = hi; stack[sp].depth = depth;
            ++sp;
            hi = cut;
        }
        if (sp == 0) break;
        --sp;
        lo = stack[sp].first; hi = stack[sp].last; depth = stack[sp].depth;
    }
    // final insertion sort
    if (n > kThreshold) {
        insertion_sort(first, first + kThreshold, before);
        for (Index* i = first + kThreshold; i != first + n; ++i) unguarded_linear_insert(i, before);
    } else {
        insertion_sort(first, first + n, before);
    }
}

}  // namespace gnusort
}  // namespace zoicb
