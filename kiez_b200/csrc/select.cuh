// Running top-`cap` selection shared by the tcgen05 and the SIMT candidate-search
// kernels.  One thread owns one query row (the TMEM lane it reads).  Per row, shared
// memory holds `cap` sorted best entries followed by an append buffer of `B` slots;
// the owner keeps the list's worst key `tau` and the buffer fill `cnt` in registers.
//
//   per element : key < tau ?  -> the owner appends (key, col) to its buffer: one
//                 predicated 64-bit store, no cross-lane traffic, no divergence
//   per 4 elems : any row's buffer nearly full? -> the whole warp merges that row's
//                 buffer into its list (by rank: every entry counts the entries below it and
//                 is scattered to its position; long lists with a full buffer: a bitonic
//                 network over lanes x registers), keeps the best `cap`, tightens tau
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

constexpr int LISTS_RANK_MAX_CNT = 12;   // see list_merge_dispatch

struct RowLists {
    ent_t *ent;     // [rows][stride]: [0,cap) sorted list, [cap, cap+B) append buffer
    int cap;
    int B;          // buffer slots, >= LISTS_MIN_SLOTS
    int stride;     // cap + B entries per row
    int rank_max = LISTS_RANK_MAX_CNT;   // lists > 32: merge by rank up to this many buffered entries
};

constexpr int LISTS_GROUP = 4;        // elements offered between two buffer-full checks
constexpr int LISTS_MIN_SLOTS = 8;    // >= 2 * LISTS_GROUP
// default: cap/2 rounded up to the group size, within [LISTS_MIN_SLOTS, 64] (the merges handle
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

// Explicit shared-space accesses of the row records: through the generic `ent_t *` the compiler
// emits LD.E / ST.E with a descriptor (two R2UR per access) whose latency a dependent chain pays
// in full.
__device__ __forceinline__ uint32_t lists_saddr(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ ent_t lds_entry(uint32_t saddr) {
    ent_t v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_entry(uint32_t saddr, ent_t v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(saddr), "l"(v) : "memory");
}

// Merge the append buffer (first `cnt` slots valid, unsorted) of one row into its sorted list BY
// RANK: every entry computes the position it has in the sorted union and is scattered there.
//   list entry i (the list is sorted):   position = i + #{buffer entries below it}
//   buffer entry j:                      position = #{list entries below it} + #{buffer entries below it}
// One loop over the buffered entries does all of it: entry j is broadcast from shared memory,
// every lane compares it with its own list / buffer entries (independent compares), and ONE
// ballot per list register counts the list entries below it -- the lane that owns entry j keeps
// that count.  No shuffles, no dependent chain: ~12 instructions per buffered entry for lists of
// <= 32.  Entries are unique ((key, column) pairs; the +inf / -1 padding of a list that is not
// full yet sits at distinct indices i and only moves up), so the positions are a permutation
// and the first `cap` of them are written exactly once.
// History (profiles/r02_ab_experiments.md blocks R, S): the bitonic sort16 + merge this replaces
// for short lists was 15 dependent 64-bit shuffle steps, ~5 k cycles per merge on a warp that
// shares its scheduler, 42 % of the row-epilogue warps' time at C4 -- and those warps are the
// critical path of the dual-direction kernel.  A first rank merge that located the buffered
// entries by binary search was no faster: 4-5 dependent generic loads of ~500 cycles each.
// All 32 lanes participate; returns the list's new worst key in every lane.
// NL = list entries per lane (cap <= 32 NL), NB = buffer entries per lane (cnt <= 32 NB).
template <int NL, int NB>
static __device__ __noinline__ float list_merge_rank(ent_t *e, int cap, int cnt, int lane) {
    const uint32_t e_s = lists_saddr(e), buf_s = e_s + (uint32_t)cap * 8u;
    ent_t xl[NL], xb[NB];
    int pl[NL], pb[NB];
    bool vl[NL];
#pragma unroll
    for (int t = 0; t < NL; ++t) {
        const int i = lane + 32 * t;
        vl[t] = i < cap;
        xl[t] = vl[t] ? lds_entry(e_s + (uint32_t)i * 8u) : EMPTY_ENTRY;
        pl[t] = vl[t] ? i : (1 << 30);
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        const int j = lane + 32 * u;
        const bool have = j < cnt;
        xb[u] = have ? lds_entry(buf_s + (uint32_t)j * 8u) : EMPTY_ENTRY;
        pb[u] = have ? 0 : (1 << 30);
    }
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {                   // warp-uniform trip count
        const ent_t y = lds_entry(buf_s + (uint32_t)j * 8u);      // broadcast
        int below = 0;                                // list entries below y (the same in every lane)
#pragma unroll
        for (int t = 0; t < NL; ++t) {
            const bool up = y < xl[t];                // y lands below my list entry: it moves up
            pl[t] += up ? 1 : 0;
            below += __popc(__ballot_sync(FULL_MASK, vl[t] && !up));   // unique entries: !up <=> xl < y
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            pb[u] += (y < xb[u]) ? 1 : 0;
            pb[u] += (lane == (j & 31) && u == (j >> 5)) ? below : 0;  // the owner of entry j
        }
    }
    __syncwarp();                                     // every read of the old contents is done
#pragma unroll
    for (int t = 0; t < NL; ++t)
        if (pl[t] < cap) sts_entry(e_s + (uint32_t)pl[t] * 8u, xl[t]);
#pragma unroll
    for (int u = 0; u < NB; ++u)
        if (pb[u] < cap) sts_entry(e_s + (uint32_t)pb[u] * 8u, xb[u]);
    __syncwarp();
    return entry_key(lds_entry(e_s + (uint32_t)(cap - 1) * 8u));
}

// Ascending bitonic MERGE of 32*R entries whose first half is ascending and whose second
// half is descending (log2(32R) compare-exchange steps instead of a full sort).
template <int R>
__device__ __forceinline__ void warp_merge_entries(ent_t (&x)[R], int lane) {
#pragma unroll
    for (int stride = 16 * R; stride > 0; stride >>= 1) {
        if (stride >= 32) {
            const int rs = stride >> 5;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((r & rs) == 0) {
                    const ent_t a = x[r], b = x[r | rs];
                    x[r] = a < b ? a : b;
                    x[r | rs] = a < b ? b : a;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const ent_t other = __shfl_xor_sync(FULL_MASK, x[r], stride);
                const bool lower = ((lane & stride) == 0);
                const ent_t mn = x[r] < other ? x[r] : other;
                const ent_t mx = x[r] < other ? other : x[r];
                x[r] = lower ? mn : mx;
            }
        }
    }
}

// Long lists with a well-filled buffer (cap > 32, cnt > LISTS_RANK_MAX_CNT): the rank merge
// costs ~(5 NL + 3 NB + 4) instructions per buffered entry, which a bitonic network over lanes x
// registers undercuts once there are many of them (measured: C2 / C3, whose dense 36-64 slot
// buffers ran 20-150 % slower through the rank merge alone, block S).  The list (ascending,
// <= 16R entries) fills the first half of a 32R-entry register tile, the buffer is sorted
// DESCENDING into the tail of the second half (the rest is +inf), and one bitonic merge yields
// the ascending union; the best `cap` go back.  R >= 2: cap <= 16R, B <= 32*RB.
template <int R, int RB>
static __device__ __noinline__ float list_merge_bitonic(ent_t *e, int cap, int cnt, int lane) {
    ent_t x[R];
    constexpr int TAIL = 32 * R - 32 * RB;               // first buffer position
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int idx = r * 32 + lane;
        const int j = idx - TAIL;
        x[r] = (idx < cap) ? e[idx] : ((j >= 0 && j < cnt) ? e[cap + j] : EMPTY_ENTRY);
    }
    // descending sort of the buffer registers = ascending sort of the complements
    ent_t y[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) y[r] = ~x[R - RB + r];
    warp_sort_entries<RB>(y, lane);
#pragma unroll
    for (int r = 0; r < RB; ++r) x[R - RB + r] = ~y[r];
    warp_merge_entries<R>(x, lane);
    ent_t worst = EMPTY_ENTRY;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int idx = r * 32 + lane;
        if (idx < cap) e[idx] = x[r];
        const ent_t w = __shfl_sync(FULL_MASK, x[r], (cap - 1) & 31);
        if (r == ((cap - 1) >> 5)) worst = w;
    }
    __syncwarp();
    return entry_key(worst);
}

// cap <= 128, B <= 64.  Lists of <= 32 always merge by rank; longer ones only while the buffer
// holds few entries (the flush at the end of an index range, sparse steady state).
__device__ __forceinline__ float list_merge_dispatch(const RowLists &L, int row, int cnt, int lane) {
    ent_t *e = L.ent + (size_t)row * L.stride;
    if (L.cap <= 32)
        return L.B <= 32 ? list_merge_rank<1, 1>(e, L.cap, cnt, lane)
                         : list_merge_rank<1, 2>(e, L.cap, cnt, lane);
    if (L.cap <= 64) {
        if (cnt <= L.rank_max)
            return cnt <= 32 ? list_merge_rank<2, 1>(e, L.cap, cnt, lane)
                             : list_merge_rank<2, 2>(e, L.cap, cnt, lane);
        return L.B <= 32 ? list_merge_bitonic<4, 1>(e, L.cap, cnt, lane)
                         : list_merge_bitonic<4, 2>(e, L.cap, cnt, lane);
    }
    if (cnt <= L.rank_max)
        return cnt <= 32 ? list_merge_rank<4, 1>(e, L.cap, cnt, lane)
                         : list_merge_rank<4, 2>(e, L.cap, cnt, lane);
    return L.B <= 32 ? list_merge_bitonic<8, 1>(e, L.cap, cnt, lane)
                     : list_merge_bitonic<8, 2>(e, L.cap, cnt, lane);
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
    // this row's append buffer (shared-space address: plain STS instead of generic stores)
    const uint32_t buf_s = lists_saddr(L.ent) + (uint32_t)(row * L.stride + L.cap) * 8u;
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
                sts_entry(buf_s + (uint32_t)cnt * 8u, pack_entry((j & 1) ? s2[1] : s2[0], col0 + j));
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
                sts_entry(buf_s + (uint32_t)cnt * 8u, pack_entry(v[j], col0 + j));
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
