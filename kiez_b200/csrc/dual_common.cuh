// Column side of the dual-direction kernels (knn_fused.cu: 3xTF32, knn_screen.cu: 1xTF32
// screen): every accumulator element whose column key  x_key[row] - 2 <x,y>  beats the
// column's threshold tau_col[col] is appended to that column's buffer in global memory.
#pragma once
#include "tc_common.cuh"

namespace kb2 {

constexpr int EMIT_Q = 48;       // pending emits per column-epilogue warp; flushed above 16
constexpr int EMIT_WORDS = 4 * EMIT_Q * 2 + 4 * EMIT_Q / 4;   // floats: keys + cols + lanes(bytes)

struct FusedParams {
    const float *x_key;          // [nq] row-side selection term
    const float *tau_col;        // [ny]
    unsigned int *col_cnt;       // [ny] rows emitted so far (may exceed col_cap: overflow)
    ent_t *col_buf;              // [ny][col_cap] packed (column key, row)
    int col_cap;
    int row_id_base;             // added to the emitted row ids (the pass covers a row segment)
};

// The emit queue of one column-epilogue warp (shared memory) + the previous drain, whose slot
// claims (atomicAdd results) are still in flight: they are consumed by the NEXT drain, so a warp
// never waits one L2 round trip (~1.5 k cycles) per drain.  That wait used to land in the tile
// time of all 16 epilogue warps of the pair through tmem_empty.
struct EmitQueue {
    float *key;
    int *col;
    unsigned char *lane;
    int n;                       // pending emits (warp-uniform)
    ent_t pend_ent;              // this lane's entry of the previous drain: packed (column key, row)
    int pend_col;                // its column, -1 = none
    unsigned int pend_pos;       // its slot: result of the atomicAdd issued by the previous drain
};

__device__ __forceinline__ void emit_queue_init(EmitQueue &Q) {
    Q.n = 0;
    Q.pend_ent = 0;
    Q.pend_col = -1;
    Q.pend_pos = 0;
}

// Write the entries whose slots the previous drain claimed.
__device__ __forceinline__ void emit_retire(const FusedParams &FP, EmitQueue &Q) {
    if (Q.pend_col >= 0 && Q.pend_pos < (unsigned int)FP.col_cap)
        FP.col_buf[(size_t)Q.pend_col * FP.col_cap + Q.pend_pos] = Q.pend_ent;
    Q.pend_col = -1;
}

// Drain up to 32 pending emits of a warp: lane i claims a slot of its entry's column buffer
// (32 independent atomics in flight) and keeps (column key, row) for emit_retire; the rest of
// the queue moves down.
__device__ __forceinline__ void emit_flush(const FusedParams &FP, EmitQueue &Q, int64_t row_base,
                                           int lane) {
    __syncwarp();
    emit_retire(FP, Q);
    const int take = min(Q.n, 32);
    float key = 0.f;
    int col = 0, owner = 0;
    if (lane < take) { key = Q.key[lane]; col = Q.col[lane]; owner = Q.lane[lane]; }
    float key2 = 0.f;
    int col2 = 0, owner2 = 0;
    const bool more = lane + 32 < Q.n;
    if (more) { key2 = Q.key[lane + 32]; col2 = Q.col[lane + 32]; owner2 = Q.lane[lane + 32]; }
    if (lane < take) {
        Q.pend_pos = atomicAdd(FP.col_cnt + col, 1u);
        Q.pend_ent = pack_entry(key, (int)(row_base + owner) + FP.row_id_base);
        Q.pend_col = col;
    }
    __syncwarp();
    if (more) { Q.key[lane] = key2; Q.col[lane] = col2; Q.lane[lane] = (unsigned char)owner2; }
    __syncwarp();
    Q.n -= take;
}

// tau_col[c0 .. c0+BN) -> registers, lane l holds columns t*32 + l; -inf masks a column.  The
// callers stage the NEGATED values in shared memory (column_tile adds them with a packed FMA).
template <int BN>
__device__ __forceinline__ void load_taucol(const float *__restrict__ tau_col, int64_t c0,
                                            int64_t y_end, int lane, float (&treg)[BN / 32]) {
#pragma unroll
    for (int t = 0; t < BN / 32; ++t) {
        const int64_t col = c0 + t * 32 + lane;
        treg[t] = (col < y_end) ? __ldg(tau_col + col) : -INFINITY;
    }
}

// One accumulator tile (this warp's 32 TMEM lanes x BN columns), column side.
// `ntk` = this warp's shared copy of the tile's NEGATED thresholds (-tau_col; +inf masks a
// column), xk = x_key of this thread's row (+inf for rows that do not exist: they never emit),
// row_base = row of lane 0.
template <int BN>
__device__ __forceinline__ void column_tile(const FusedParams &FP, EmitQueue &Q, const float *ntk,
                                            uint32_t taddr, int64_t c0, float xk, int64_t row_base,
                                            int lane) {
    constexpr int NCH = BN / 32;
    const float nxk = -xk;
    const uint32_t ntk_saddr = smem_u32(ntk);
    auto process = [&](const uint32_t (&r)[32], int ch) {
        // g = -2 acc - tau_col;  emit when  key_x - 2 acc < tau_col  <=>  g < -key_x.  The g are
        // not kept (they are recomputed on the rare slow path): the accumulator registers of the
        // chunk stay the only 32-wide array that is live across the test.  Four at a time:
        // one broadcast LDS.128 of the negated thresholds + two packed FFMA2 (key_finish4).
        auto g4 = [&](int j, float &g0, float &g1, float &g2, float &g3) {
            key_finish4(ntk_saddr + (uint32_t)(ch * 32 + j) * 4u, r[j], r[j + 1], r[j + 2], r[j + 3],
                        g0, g1, g2, g3);
        };
        float m4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float a0, a1, a2, a3, b0, b1, b2, b3;
            g4(8 * q, a0, a1, a2, a3);
            g4(8 * q + 4, b0, b1, b2, b3);
            m4[q] = fminf(fminf(fminf(a0, a1), fminf(a2, a3)), fminf(fminf(b0, b1), fminf(b2, b3)));
        }
        const float gmin = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
        if (!__any_sync(FULL_MASK, gmin < nxk)) return;
        // Slow path (some lane has a survivor).  Every emit used to cost the CTA pair a chain of
        // ~13 dependent warp votes (8 groups, 4 ballots each where one passed), and a late
        // epilogue warp holds the accumulator buffer of all 16.  Now: each lane builds its own
        // 32-bit pass mask without votes; per round ONE ballot ranks the lanes that still hold
        // a survivor and each of them queues its lowest one (r[j] with a runtime j comes from a
        // 5-level select tree, not local memory).  Rounds = most survivors in one lane, almost
        // always 1.  The queue is drained 17-48 entries at a time so that the global atomics
        // that assign the column-buffer slots are in flight together.
        unsigned int m = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            float g0, g1, g2, g3;
            g4(j, g0, g1, g2, g3);
            m |= (g0 < nxk) ? (1u << j) : 0u;
            m |= (g1 < nxk) ? (2u << j) : 0u;
            m |= (g2 < nxk) ? (4u << j) : 0u;
            m |= (g3 < nxk) ? (8u << j) : 0u;
        }
        for (;;) {
            const bool have = m != 0;
            const unsigned int has = __ballot_sync(FULL_MASK, have);
            if (has == 0) break;
            if (have) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                for (int i = 0; i < 16; ++i) s16[i] = (j & 16) ? r[i + 16] : r[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) s8[i] = (j & 8) ? s16[i + 8] : s16[i];
#pragma unroll
                for (int i = 0; i < 4; ++i) s4[i] = (j & 4) ? s8[i + 4] : s8[i];
#pragma unroll
                for (int i = 0; i < 2; ++i) s2[i] = (j & 2) ? s4[i + 2] : s4[i];
                const uint32_t val = (j & 1) ? s2[1] : s2[0];
                const int slot = Q.n + __popc(has & ((1u << lane) - 1));
                Q.key[slot] = fmaf(-2.f, __uint_as_float(val), xk);
                Q.col[slot] = (int)(c0 + ch * 32 + j);
                Q.lane[slot] = (unsigned char)lane;
            }
            Q.n += __popc(has);
            if (Q.n > EMIT_Q - 32) emit_flush(FP, Q, row_base, lane);
        }
    };
    // (Keeping the next chunk's tcgen05.ld in flight while this one is tested -- with and without
    // a larger register budget -- measured no faster, profiles/r02_ab_experiments.md block M: two
    // epilogue warps per scheduler already hide the TMEM latency.)
    uint32_t ra[32];
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
        tmem_ld_32x32b_x32(taddr + ch * 32, ra);
        tmem_ld_wait();
        process(ra, ch);
    }
}

}  // namespace kb2
