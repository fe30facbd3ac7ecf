// Running top-`cap` selection shared by the tcgen05 and the SIMT candidate-search
// kernels.  One thread owns one query row (the TMEM lane it reads); each row has
// a sorted list of its `cap` best (key, column) pairs in shared memory and the
// owner keeps the list's current worst key `tau` in a register.  After the list
// has filled, an element survives the `key < tau` test with probability
// ~cap/position, so the common case is one compare per element; survivors are
// inserted by the whole warp cooperating on the owner's list (cost independent
// of which lane owns the row, no divergence).  Ties keep the earlier column,
// like sklearn's heap (utils/_heap.pyx:46 rejects val >= heap_max).
#pragma once
#include "common.cuh"

namespace kb2 {

struct RowLists {
    float *keys;   // [rows][cap] ascending
    int *cols;     // [rows][cap]
    int cap;
};

__device__ __forceinline__ void lists_reset(const RowLists &L, int row_begin, int rows, int lane) {
    // called by one warp for its own `rows` rows
    for (int i = lane; i < rows * L.cap; i += 32) {
        L.keys[row_begin * L.cap + i] = INFINITY;
        L.cols[row_begin * L.cap + i] = -1;
    }
    __syncwarp();
}

// Insert (nv, ncol) into the sorted list of `row`; all 32 lanes participate.
// Returns the new worst key of the list (valid in every lane).
static __device__ __noinline__ float list_insert(const RowLists &L, int row, float nv, int ncol,
                                             int lane) {
    float *k = L.keys + (size_t)row * L.cap;
    int *c = L.cols + (size_t)row * L.cap;
    constexpr int MAXT = 4;   // cap <= 128
    float nk[MAXT];
    int nc[MAXT];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const int p = lane + 32 * t;
        if (p < L.cap) {
            const float kp = k[p];
            const int cp = c[p];
            const float km = (p > 0) ? k[p - 1] : -INFINITY;
            const int cm = (p > 0) ? c[p - 1] : -1;
            if (kp <= nv) { nk[t] = kp; nc[t] = cp; }            // stays in place
            else if (km <= nv) { nk[t] = nv; nc[t] = ncol; }     // insertion point
            else { nk[t] = km; nc[t] = cm; }                     // shifted right by one
        }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
        const int p = lane + 32 * t;
        if (p < L.cap) { k[p] = nk[t]; c[p] = nc[t]; }
    }
    __syncwarp();
    return k[L.cap - 1];
}

// Offer NV consecutive columns [col0, col0+NV) of this thread's row.
// v[j] must already be +inf for masked columns.  `row` is the thread's list row,
// `tau` its current threshold (-inf for rows that do not exist).
template <int NV>
__device__ __forceinline__ void select_chunk(const RowLists &L, int row, const float (&v)[NV],
                                             int col0, float &tau, int lane) {
    float mn = v[0];
#pragma unroll
    for (int j = 1; j < NV; ++j) mn = fminf(mn, v[j]);
    if (!__any_sync(FULL_MASK, mn < tau)) return;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        unsigned need = __ballot_sync(FULL_MASK, v[j] < tau);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const float nv = __shfl_sync(FULL_MASK, v[j], src);
            const int r = __shfl_sync(FULL_MASK, row, src);
            const float t = list_insert(L, r, nv, col0 + j, lane);
            if (lane == src) tau = t;
        }
    }
}

}  // namespace kb2
