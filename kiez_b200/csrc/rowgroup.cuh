// Register-resident row processing for the memory-bound (n, c) kernels: a row of c
// values is held by a group of G lanes (G = 8/16/32, 32/G rows per warp) with E values
// per lane (element e = t*G + lane_in_group, so every load is coalesced).  Reductions are
// xor-shuffles inside the group; the top-k sort is a bitonic network over lanes x registers
// keyed by (value, position) with NaN last -- the order numpy's argsort/argpartition give
// (kiez/hubness_reduction/base.py:80-87).  No shared memory, no block barriers.  (Rows of <= 16
// values take the thread-per-row kernels of rescale.cu instead, and k <= 16 of a wider row are
// selected by arg-min rounds, rescale.cu warp_select_topk.)
#pragma once
#include "common.cuh"

namespace kb2 {

constexpr int RG_POS_PAD = 0x7fffffff;

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ double group_min(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// value of element `e` of the row (held in register e / G of lane e % G of the group)
template <int G, int E>
__device__ __forceinline__ double group_element(const double (&x)[E], int e, int lane) {
    double v = 0.0;
#pragma unroll
    for (int t = 0; t < E; ++t)
        if (t == e / G) v = x[t];
    return __shfl_sync(FULL_MASK, v, (lane & ~(G - 1)) + (e % G));
}

// ascending bitonic sort of the G*E (key, pos) pairs of each group
template <int G, int E>
__device__ __forceinline__ void group_sort(double (&key)[E], int (&pos)[E], int gl) {
    constexpr int N = G * E;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= G) {
                const int ts = stride / G;
#pragma unroll
                for (int t = 0; t < E; ++t) {
                    if ((t & ts) == 0) {
                        const bool up = (((t * G) & size) == 0);
                        const bool a_less = pair_less(key[t], pos[t], key[t | ts], pos[t | ts]);
                        if (a_less != up) {
                            const double k0 = key[t]; key[t] = key[t | ts]; key[t | ts] = k0;
                            const int p0 = pos[t]; pos[t] = pos[t | ts]; pos[t | ts] = p0;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < E; ++t) {
                    const double ok = __shfl_xor_sync(FULL_MASK, key[t], stride);
                    const int op = __shfl_xor_sync(FULL_MASK, pos[t], stride);
                    const int e = t * G + gl;
                    const bool up = ((e & size) == 0);
                    const bool lower = ((gl & stride) == 0);
                    const bool keep_min = (lower == up);
                    const bool self_less = pair_less(key[t], pos[t], ok, op);
                    if (keep_min != self_less) { key[t] = ok; pos[t] = op; }
                }
            }
        }
    }
}

}  // namespace kb2
