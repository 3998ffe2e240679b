// SumTree pieces shared by sumtree.cu and replay.cu: round-to-nearest arithmetic in the dtype of the draw and the root-to-leaf
// descent `get(t, v)` of RLCore's SumTree (call site src/algorithms/dqns/common.jl:18).
#pragma once
#include "common.cuh"

namespace zoo {

template <typename T> struct Rn;
template <> struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

struct SumTreeView {
    const float *tree;  // 1-based heap
    long long tree_len, nparents, capacity, first;
};

// get(t, v): v <= tree[left] ? go left : (v -= tree[left]; go right); returns the 1-based logical index and the leaf priority
template <typename T>
__device__ __forceinline__ long long sumtree_descend(const SumTreeView &t, T v, float *prio) {
    long long node = 1;
    for (;;) {
        const long long l = 2 * node;
        if (l > t.tree_len) break;
        const float tl = __ldg(t.tree + l);
        if (v <= (T)tl) node = l;
        else { v = Rn<T>::sub(v, (T)tl); node = l + 1; }
        if (node > t.tree_len) { node = t.tree_len; break; }
    }
    *prio = t.tree[node];
    const long long k = node - t.nparents;
    return k >= t.first ? k - t.first + 1 : k + t.capacity - t.first + 1;
}

}  // namespace zoo
