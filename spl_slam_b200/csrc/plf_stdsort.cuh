// plf_stdsort.cuh -- libstdc++'s std::sort (bits/stl_algo.h), replayed on a permutation, for one device thread.
//
// Lineextractor::ComputeLsdWithLbd sorts the lines of an octave with std::sort by `a.response > b.response`
// (src/Lineextractor.cc:175, include/Lineextractor.h:66-71).  std::sort is not stable: which of several lines with EQUAL
// response survive the quota cut, and in which order they come out, is defined only by the exact sequence of comparisons
// and moves of libstdc++'s introsort (median-of-3 moved to the front, unguarded Hoare partition, ranges of <= 16 left to a
// final insertion sort, heapsort once the depth budget 2 * floor(log2 n) is spent).  k_line_select ranks lines in parallel
// when all responses are distinct (then every correct sort agrees) and calls this on ONE thread when there are ties.
// Checked against the real std::sort through the compiled reference (tests/test_gpu_parity.py::test_line_response_ties).
#pragma once

namespace plf_stdsort {

#define PLF_SS_LT(a, b) (k[a] > k[b])

template <typename I>
__host__ __device__ inline void unguarded_linear_insert(const float* k, I* p, int last)
{
    const int val = p[last];
    int next = last - 1;
    while (PLF_SS_LT(val, p[next])) { p[last] = p[next]; last = next; --next; }
    p[last] = val;
}

template <typename I>
__host__ __device__ inline void insertion_sort(const float* k, I* p, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (PLF_SS_LT(p[i], p[first])) {
            const int val = p[i];
            for (int j = i; j > first; --j) p[j] = p[j - 1];
            p[first] = val;
        } else
            unguarded_linear_insert(k, p, i);
    }
}

template <typename I>
__host__ __device__ inline void adjust_heap(const float* k, I* p, int first, int hole, int len, int value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (PLF_SS_LT(p[first + child], p[first + child - 1])) child--;
        p[first + hole] = p[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        p[first + hole] = p[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;   // __push_heap
    while (hole > top && PLF_SS_LT(p[first + parent], value)) {
        p[first + hole] = p[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    p[first + hole] = value;
}

template <typename I>
__host__ __device__ inline void heap_sort(const float* k, I* p, int first, int last)   // __partial_sort(first, last, last)
{
    const int len = last - first;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; parent--) {
            adjust_heap(k, p, first, parent, len, p[first + parent]);
            if (parent == 0) break;
        }
    while (last - first > 1) {
        --last;
        const int value = p[last];
        p[last] = p[first];
        adjust_heap(k, p, first, 0, last - first, value);
    }
}

// p[0..n) must hold 0..n-1; k[] are the keys; on return p is the order std::sort leaves the records in
template <typename I>
__host__ __device__ inline void sort_desc(const float* k, int n, I* p)
{
    if (n == 0) return;
    int lg = 0;
    for (int v = n; v > 1; v >>= 1) lg++;
    // __introsort_loop recurses on the right part and loops on the left; the two parts are disjoint, so an explicit stack
    // in any order leaves the same array.  Depth <= 2 lg n + 1 entries.
    int st_first[72], st_last[72], st_depth[72], sp = 0;
    st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
    while (sp) {
        --sp;
        int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
        while (last - first > 16) {
            if (depth == 0) { heap_sort(k, p, first, last); break; }
            --depth;
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            int m;
            if (PLF_SS_LT(p[a], p[b])) m = PLF_SS_LT(p[b], p[c]) ? b : PLF_SS_LT(p[a], p[c]) ? c : a;
            else m = PLF_SS_LT(p[a], p[c]) ? a : PLF_SS_LT(p[b], p[c]) ? c : b;
            I t = p[first]; p[first] = p[m]; p[m] = t;
            int lo = first + 1, hi = last;
            for (;;) {
                while (PLF_SS_LT(p[lo], p[first])) ++lo;
                --hi;
                while (PLF_SS_LT(p[first], p[hi])) --hi;
                if (!(lo < hi)) break;
                t = p[lo]; p[lo] = p[hi]; p[hi] = t;
                ++lo;
            }
            st_first[sp] = lo; st_last[sp] = last; st_depth[sp] = depth; sp++;
            last = lo;
        }
    }
    if (n > 16) {
        insertion_sort(k, p, 0, 16);
        for (int i = 16; i != n; ++i) unguarded_linear_insert(k, p, i);
    } else
        insertion_sort(k, p, 0, n);
}

#undef PLF_SS_LT

} // namespace plf_stdsort
