// Running top-`cap` selection shared by the tcgen05 and the SIMT candidate-search
// kernels.  One thread owns one query row (the TMEM lane it reads).  Per row, shared
// memory holds `cap` sorted best entries followed by an append buffer of `B` slots;
// the owner keeps the list's worst key `tau` and the buffer fill `cnt` in registers.
//
//   per element : key < tau ?  -> the owner appends (key, col) to its buffer: one
//                 predicated 64-bit store, no cross-lane traffic, no divergence
//   per 4 elems : any row's buffer nearly full? -> the whole warp merges that row's
//                 buffer into its list by rank (every entry counts the entries below it and
//                 is scattered to its position), keeps the best `cap`, tightens tau
//
// Between merges tau is stale, so a few more elements pass than with an exact
// threshold, but each costs a store instead of a serialized sorted insert: after the
// list has filled an element passes with probability ~cap/position, and the number of
// merges per row is ~log(m/cap)/log(1 + B/(2 cap)).
// Entries are packed (order-preserving key bits << 32 | column): ties keep the lower
// column, like sklearn's heap, which rejects val >= heap_max (utils/_heap.pyx:46).
#pragma once
#include "common.cuh"

namespace kb2 {

typedef unsigned long long ent_t;

__device__ __forceinline__ ent_t pack_entry(float key, int col) {
    unsigned u = __float_as_uint(key);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;           // unsigned order == float order
    return ((ent_t)u << 32) | (unsigned)col;
}
__device__ __forceinline__ float entry_key(ent_t e) {
    unsigned u = (unsigned)(e >> 32);
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}
__device__ __forceinline__ int entry_col(ent_t e) { return (int)(unsigned)(e & 0xFFFFFFFFu); }
// +inf key, column -1
constexpr ent_t EMPTY_ENTRY = ((ent_t)0xFF800000u << 32) | 0xFFFFFFFFull;

struct RowLists {
    ent_t *ent;     // [rows][stride]: [0,cap) sorted list, [cap, cap+B) append buffer
    int cap;
    int B;          // buffer slots, >= LISTS_MIN_SLOTS
    int stride;     // cap + B entries per row
};

constexpr int LISTS_GROUP = 4;        // elements offered between two buffer-full checks
constexpr int LISTS_MIN_SLOTS = 8;    // >= 2 * LISTS_GROUP
// default: cap/2 rounded up to the group size, within [LISTS_MIN_SLOTS, 64] (list_merge handles
// cap <= 128 and B <= 64)
__host__ __device__ inline int lists_buffer_slots(int cap) {
    int b = ((cap / 2 + LISTS_GROUP - 1) / LISTS_GROUP) * LISTS_GROUP;
    return b < LISTS_MIN_SLOTS ? LISTS_MIN_SLOTS : (b > 64 ? 64 : b);
}
__host__ __device__ inline int lists_stride(int cap, int B) { return cap + B; }
__host__ __device__ inline size_t lists_bytes(int rows, int cap, int B) {
    return (size_t)rows * lists_stride(cap, B) * sizeof(ent_t);
}

// called by one warp for its own `rows` rows (only the sorted part needs a reset)
__device__ __forceinline__ void lists_reset(const RowLists &L, int row_begin, int rows, int lane) {
    for (int r = 0; r < rows; ++r)
        for (int p = lane; p < L.cap; p += 32) L.ent[(size_t)(row_begin + r) * L.stride + p] = EMPTY_ENTRY;
    __syncwarp();
}

// Ascending bitonic sort of 32*R entries held as x[r] in lane `lane` <-> element r*32+lane.
template <int R>
__device__ __forceinline__ void warp_sort_entries(ent_t (&x)[R], int lane) {
    constexpr int N = 32 * R;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rs) == 0) {
                        const bool up = (((r * 32) & size) == 0);
                        const ent_t a = x[r], b = x[r | rs];
                        const bool sw = (a > b) == up;
                        x[r] = sw ? b : a;
                        x[r | rs] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const ent_t other = __shfl_xor_sync(FULL_MASK, x[r], stride);
                    const int e = r * 32 + lane;
                    const bool up = ((e & size) == 0);
                    const bool lower = ((lane & stride) == 0);
                    const bool keep_min = (lower == up);
                    const ent_t mn = x[r] < other ? x[r] : other;
                    const ent_t mx = x[r] < other ? other : x[r];
                    x[r] = keep_min ? mn : mx;
                }
            }
        }
    }
}

// Merge the append buffer (first `cnt` slots valid, unsorted) of one row into its sorted list BY
// RANK: every entry computes the position it has in the sorted union and is scattered there.
//   list entry i (the list is sorted):   position = i + #{buffer entries below it}
//   buffer entry:                        position = #{list entries below it}  (binary search)
//                                                 + #{buffer entries below it}
// Entries are unique ((key, column) pairs; the +inf / -1 padding of a list that is not full yet
// sits at distinct indices i and only moves up), so the positions are a permutation and the
// first `cap` of them are written exactly once.  The counting loops are independent compares
// against broadcast shared-memory reads: ~8 instructions per buffered entry and no dependent
// shuffle chain.  The bitonic sort + merge over lanes x registers this replaces was 15 (cap 16)
// to 29 (cap 112) dependent 64-bit shuffle steps, ~3-5 k cycles per merge, and took 42 % of the
// row-epilogue warps' time at C4 (ncu source view, profiles/r02_ab_experiments.md block R) --
// those warps are the critical path of the dual-direction kernel.
// All 32 lanes participate; returns the list's new worst key in every lane.
// NL = list entries per lane (cap <= 32 NL), NB = buffer entries per lane (B <= 32 NB).
template <int NL, int NB>
static __device__ __noinline__ float list_merge(ent_t *e, int cap, int cnt, int lane) {
    const ent_t *buf = e + cap;
    ent_t xl[NL], xb[NB];
    int pl[NL], pb[NB];
#pragma unroll
    for (int t = 0; t < NL; ++t) {
        const int i = lane + 32 * t;
        xl[t] = (i < cap) ? e[i] : EMPTY_ENTRY;
        pl[t] = (i < cap) ? i : (1 << 30);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int j = lane + 32 * u;
        const bool have = j < cnt;
        xb[u] = have ? buf[j] : EMPTY_ENTRY;
        int lo = 0;
        if (have) {                                   // entries of the sorted list below xb[u]
            int hi = cap;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (e[mid] < xb[u]) lo = mid + 1; else hi = mid;
            }
        }
        pb[u] = have ? lo : (1 << 30);
    }
    for (int j = 0; j < cnt; ++j) {                   // warp-uniform trip count, broadcast reads
        const ent_t y = buf[j];
#pragma unroll
        for (int t = 0; t < NL; ++t) pl[t] += (y < xl[t]) ? 1 : 0;
#pragma unroll
        for (int u = 0; u < NB; ++u) pb[u] += (y < xb[u]) ? 1 : 0;
    }
    __syncwarp();                                     // every read of the old contents is done
#pragma unroll
    for (int t = 0; t < NL; ++t)
        if (pl[t] < cap) e[pl[t]] = xl[t];
#pragma unroll
    for (int u = 0; u < NB; ++u)
        if (pb[u] < cap) e[pb[u]] = xb[u];
    __syncwarp();
    return entry_key(e[cap - 1]);
}

// cap <= 128, B <= 64 (lists_buffer_slots and the screen kernel's plan stay within that)
__device__ __forceinline__ float list_merge_dispatch(const RowLists &L, int row, int cnt, int lane) {
    ent_t *e = L.ent + (size_t)row * L.stride;
    if (L.cap <= 32)
        return L.B <= 32 ? list_merge<1, 1>(e, L.cap, cnt, lane) : list_merge<1, 2>(e, L.cap, cnt, lane);
    if (L.cap <= 64)
        return L.B <= 32 ? list_merge<2, 1>(e, L.cap, cnt, lane) : list_merge<2, 2>(e, L.cap, cnt, lane);
    return L.B <= 32 ? list_merge<4, 1>(e, L.cap, cnt, lane) : list_merge<4, 2>(e, L.cap, cnt, lane);
}

// Merge every row of this warp whose buffer fill satisfies `want` (warp-uniform loop).
__device__ __forceinline__ void merge_rows(const RowLists &L, int row, float &tau, int &cnt,
                                           bool want, int lane) {
    unsigned need = __ballot_sync(FULL_MASK, want);
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int r = __shfl_sync(FULL_MASK, row, src);
        const int c = __shfl_sync(FULL_MASK, cnt, src);
        __syncwarp();
        const float t = list_merge_dispatch(L, r, c, lane);
        if (lane == src) { tau = t; cnt = 0; }
    }
}

// Offer NV (multiple of LISTS_GROUP) consecutive columns [col0, col0+NV) of this thread's row.
// v[j] must already be +inf for masked columns; `tau` = -inf for rows that do not exist.
template <int NV>
__device__ __forceinline__ void select_chunk(const RowLists &L, int row, const float (&v)[NV],
                                             int col0, float &tau, int &cnt, int lane) {
    // chunk minimum as 4 independent chains (the lone epilogue warp of a scheduler is
    // latency-bound: a single 31-deep dependent chain would cost ~5 clk per link)
    float m4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        m4[q] = v[q * (NV / 4)];
#pragma unroll
        for (int j = 1; j < NV / 4; ++j) m4[q] = fminf(m4[q], v[q * (NV / 4) + j]);
    }
    const float mn = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
    if (!__any_sync(FULL_MASK, mn < tau)) return;
    ent_t *buf = L.ent + (size_t)row * L.stride + L.cap;
    if constexpr (NV == 32) {
        // Sparse chunk (the steady state: a fraction of a survivor per 32 x 32 chunk): every
        // lane's survivors fit its buffer, so they are appended without the per-group votes --
        // 3 votes per passing chunk instead of 10.  v[j] with a runtime j comes from a 5-level
        // select tree (no local memory).  Dense chunks (list fill phase) take the loop below.
        unsigned int m = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) m |= (v[j] < tau) ? (1u << j) : 0u;
        if (!__any_sync(FULL_MASK, cnt + __popc(m) > L.B)) {
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                float s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                for (int i = 0; i < 16; ++i) s16[i] = (j & 16) ? v[i + 16] : v[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) s8[i] = (j & 8) ? s16[i + 8] : s16[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) s4[i] = (j & 4) ? s8[i + 4] : s8[i];
#pragma unroll
                for (int i = 0; i < 2; ++i) s2[i] = (j & 2) ? s4[i + 2] : s4[i];
                buf[cnt] = pack_entry((j & 1) ? s2[1] : s2[0], col0 + j);
                ++cnt;
            }
            const bool full = cnt > L.B - LISTS_GROUP;
            if (__any_sync(FULL_MASK, full)) merge_rows(L, row, tau, cnt, full, lane);
            return;
        }
    }
#pragma unroll
    for (int g = 0; g < NV / LISTS_GROUP; ++g) {
        // most groups of a passing chunk still hold no survivor: skip them with one vote
        bool any_pass = false;
#pragma unroll
        for (int j = LISTS_GROUP * g; j < LISTS_GROUP * (g + 1); ++j) any_pass |= (v[j] < tau);
        if (!__any_sync(FULL_MASK, any_pass)) continue;
#pragma unroll
        for (int j = LISTS_GROUP * g; j < LISTS_GROUP * (g + 1); ++j) {
            if (v[j] < tau) {
                buf[cnt] = pack_entry(v[j], col0 + j);
                ++cnt;
            }
        }
        // a buffer with fewer than LISTS_GROUP free slots could overflow in the next group
        const bool full = cnt > L.B - LISTS_GROUP;
        if (__any_sync(FULL_MASK, full)) merge_rows(L, row, tau, cnt, full, lane);
    }
}

// After the last column: fold what is left in the buffers into the lists.
__device__ __forceinline__ void lists_flush(const RowLists &L, int row, float &tau, int &cnt,
                                            int lane) {
    merge_rows(L, row, tau, cnt, cnt > 0, lane);
    __syncwarp();
}

}  // namespace kb2
