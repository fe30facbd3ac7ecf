// Dual-direction candidate search: ONE pass over the row x column tiles produces
//   (F) row-wise:    for every row its `cap` best columns    (lists in shared memory, as knn_tc2.cu)
//   (R) column-wise: for every column every row whose key beats the column's threshold
//                    tau_col[col], appended to a per-column buffer in global memory
// so that kiez's reverse pass (HubnessReduction.fit, hubness_reduction/base.py:37-42) and
// forward pass (base.py:92-94) share one contraction: 2 n m d flop instead of 4 n m d.
//
// tau_col must be an UPPER bound of the column's final `cap`-th best key.  The host obtains
// it from an ordinary search of the columns against a strided SAMPLE of the rows (the cap-th
// best within any subset bounds the cap-th best overall), so the expected number of rows
// emitted per column is cap * n / n_sample, independent of the data distribution.
//   row key (F):    key_y[col] - 2 <x,y>      (as in knn_tc.cu)
//   column key (R): key_x[row] - 2 <x,y>      emitted when  < tau_col[col]
//
// Same CTA-pair tcgen05 pipeline as knn_tc2.cu, with a second set of epilogue warps:
//   warps 0-3  row epilogue (F): tcgen05.ld, key finish, selection (select.cuh)
//   warps 4-7  column epilogue (R): tcgen05.ld of the SAME accumulator (warp w+4 shares the
//              TMEM lane quadrant of warp w), threshold test, atomic append
//   warp  8    TMA producer, warp 9 MMA issuer (CTA 0), warps 10-11 idle
// tmem_empty therefore collects 8 warps x 2 CTAs.
#include "dual_common.cuh"

namespace kb2 {

constexpr int F_THREADS = 384;
constexpr int F_BN = 256;        // index (column) rows per tile of the CTA pair
constexpr int F_HALF = 128;

template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F_THREADS, 1)
knn_fused_kernel(const __grid_constant__ CUtensorMap map_q_hi,
                 const __grid_constant__ CUtensorMap map_q_lo,
                 const __grid_constant__ CUtensorMap map_y_hi,
                 const __grid_constant__ CUtensorMap map_y_lo, const TcParams P,
                 const FusedParams FP) {
    using Cfg = StageCfg<F_HALF, BK>;
    constexpr int BN = F_BN;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    // per epilogue warp a BN-float tile: warps 0-3 key_y, warps 4-7 tau_col
    float *tile_s = reinterpret_cast<float *>(stage_base + (size_t)P.stages * Cfg::STAGE_BYTES);
    // per column-epilogue warp: a queue of pending emits (key, column, owner lane)
    float *emit_key = tile_s + 8 * BN;                                   // [4][EMIT_Q]
    int *emit_col = reinterpret_cast<int *>(emit_key + 4 * EMIT_Q);      // [4][EMIT_Q]
    unsigned char *emit_lane = reinterpret_cast<unsigned char *>(emit_col + 4 * EMIT_Q);   // [4][EMIT_Q]
    RowLists L;
    L.cap = P.cap;
    L.B = P.buf_slots;
    L.stride = lists_stride(P.cap, P.buf_slots);
    L.ent = reinterpret_cast<ent_t *>(tile_s + 8 * BN + EMIT_WORDS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(L.ent + (size_t)BM * L.stride);
    uint64_t *full_bar = bars;
    uint64_t *empty_bar = bars + MAX_STAGES;
    uint64_t *tmem_full = bars + 2 * MAX_STAGES;
    uint64_t *tmem_empty = bars + 2 * MAX_STAGES + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);
    {
        uint32_t dyn_size;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
        if (reinterpret_cast<unsigned char *>(tmem_slot + 4) > smem_raw + dyn_size) {
            if (threadIdx.x == 0)
                printf("kiez_b200: knn_fused shared-memory carve-up exceeds the launch size\n");
            __trap();
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t q_pairs = (P.q_tiles + 1) / 2;
    const int64_t num_units = q_pairs * P.splits;
    const int64_t pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&map_q_hi);
        tma_prefetch_desc(&map_q_lo);
        tma_prefetch_desc(&map_y_hi);
        tma_prefetch_desc(&map_y_lo);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 16);   // 8 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t stage_u32 = smem_u32(stage_base);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    const uint32_t tmem_empty_leader = smem_u32(tmem_empty) & PEER_BIT_MASK;

    if (warp == 8) {
        // ------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        int wave = 0;
        for (int64_t u = pair_id; u < num_units; u += num_pairs, ++wave) {
            if (P.wave_sync && wave > 0) {
                // pairs taking part in this wave: those that still have a unit
                const int64_t first = (int64_t)wave * num_pairs;
                const unsigned int expected = (unsigned int)min(num_pairs, num_units - first);
                if (rank == 0 && lane == 0) wave_barrier(wave, expected);
                __syncwarp();
                // the peer CTA's producer follows through the shared full/empty barriers
            }
            const int64_t qt = 2 * (u % q_pairs) + rank;
            const int split = (int)(u / q_pairs);
            const int64_t y_begin = (int64_t)split * P.per_split;
            const int64_t y_end = min(P.ny, y_begin + P.per_split);
            const int q_row0 = (int)(qt * BM);
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                const int y_row0 = (int)c0 + (int)rank * F_HALF;
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        const uint32_t st = stage_u32 + (uint32_t)stage * Cfg::STAGE_BYTES;
                        const uint32_t fb = (full_u32 + (uint32_t)stage * 8) & PEER_BIT_MASK;
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        tma_load_2d_pair(&map_q_hi, st, fb, kc * BK, q_row0);
                        tma_load_2d_pair(&map_q_lo, st + Cfg::A_BYTES, fb, kc * BK, q_row0);
                        tma_load_2d_pair(&map_y_hi, st + 2 * Cfg::A_BYTES, fb, kc * BK, y_row0);
                        tma_load_2d_pair(&map_y_lo, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, fb,
                                         kc * BK, y_row0);
                    }
                    __syncwarp();
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t u = pair_id; u < num_units; u += num_pairs) {
                const int split = (int)(u / q_pairs);
                const int64_t y_begin = (int64_t)split * P.per_split;
                const int64_t y_end = min(P.ny, y_begin + P.per_split);
                for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = stage_u32 + (uint32_t)stage * Cfg::STAGE_BYTES;
                            const uint64_t d_qhi = make_smem_desc<Cfg>(sa);
                            const uint64_t d_qlo = make_smem_desc<Cfg>(sa + Cfg::A_BYTES);
                            const uint64_t d_yhi = make_smem_desc<Cfg>(sa + 2 * Cfg::A_BYTES);
                            const uint64_t d_ylo = make_smem_desc<Cfg>(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_tf32_pair(tmem_d, d_qlo + 2 * k, d_yhi + 2 * k, idesc, (kc | k) != 0);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_tf32_pair(tmem_d, d_qhi + 2 * k, d_ylo + 2 * k, idesc, 1u);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_tf32_pair(tmem_d, d_qhi + 2 * k, d_yhi + 2 * k, idesc, 1u);
                            umma_commit_pair(empty_u32 + (uint32_t)stage * 8, 0x3);
                            if (kc == P.kchunks - 1)
                                umma_commit_pair(smem_u32(&tmem_full[acc]), 0x3);
                        }
                        __syncwarp();
                        if (++stage == P.stages) { stage = 0; phase ^= 1; }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp < 4) {
        // ------------------------------------------------------ row epilogue (F), both CTAs
        const int lrow = warp * 32 + lane;
        float *yk = tile_s + warp * BN;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t u = pair_id; u < num_units; u += num_pairs) {
            const int64_t qt = 2 * (u % q_pairs) + rank;
            const int split = (int)(u / q_pairs);
            const int64_t y_begin = (int64_t)split * P.per_split;
            const int64_t y_end = min(P.ny, y_begin + P.per_split);
            const int64_t grow = qt * BM + lrow;
            lists_reset(L, warp * 32, 32, lane);
            float tau = (grow < P.nq) ? INFINITY : -INFINITY;
            int cnt = 0;
            float ykreg[BN / 32];
            load_ykey<BN>(P.y_key, y_begin, y_end, lane, ykreg);
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                __syncwarp();
#pragma unroll
                for (int t = 0; t < BN / 32; ++t) yk[t * 32 + lane] = ykreg[t];
                __syncwarp();
                load_ykey<BN>(P.y_key, c0 + BN, y_end, lane, ykreg);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
                epilogue_tile<BN, false>(L, lrow, yk, taddr, c0, tau, cnt, lane);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + (uint32_t)acc * 8);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            lists_flush(L, lrow, tau, cnt, lane);
            write_lists(L, P, warp, lane, qt * BM, split);
        }
    } else if (warp < 8) {
        // ------------------------------------------------------ column epilogue (R), both CTAs
        const int quad = warp - 4;                         // TMEM lane quadrant
        const int lrow = quad * 32 + lane;
        float *tk = tile_s + warp * BN;                    // this warp's tile of NEGATED tau_col
        EmitQueue Q;
        Q.key = emit_key + quad * EMIT_Q;
        Q.col = emit_col + quad * EMIT_Q;
        Q.lane = emit_lane + quad * EMIT_Q;
        emit_queue_init(Q);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t u = pair_id; u < num_units; u += num_pairs) {
            const int64_t qt = 2 * (u % q_pairs) + rank;
            const int split = (int)(u / q_pairs);
            const int64_t y_begin = (int64_t)split * P.per_split;
            const int64_t y_end = min(P.ny, y_begin + P.per_split);
            const int64_t grow = qt * BM + lrow;
            // rows that do not exist never emit: -xk = -inf
            const float xk = (grow < P.nq) ? __ldg(FP.x_key + grow) : INFINITY;
            const int64_t row_base = qt * BM + quad * 32;  // row of lane 0 of this warp
            float treg[BN / 32];
            load_taucol<BN>(FP.tau_col, y_begin, y_end, lane, treg);
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                __syncwarp();
#pragma unroll
                for (int t = 0; t < BN / 32; ++t) tk[t * 32 + lane] = -treg[t];
                __syncwarp();
                load_taucol<BN>(FP.tau_col, c0 + BN, y_end, lane, treg);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
                column_tile<BN>(FP, Q, tk, taddr, c0, xk, row_base, lane);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + (uint32_t)acc * 8);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            while (Q.n > 0) emit_flush(FP, Q, row_base, lane);   // rows change with the unit
        }
        emit_retire(FP, Q);                                      // the last drain's entries
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(TMEM_COLS)
                     : "memory");
    }
}

// Sort the first n entries of a column (in the warp's shared-memory slice e[0..P2), padded with
// EMPTY_ENTRY) ascending: bitonic network over the smallest power of two holding them.
__device__ __forceinline__ void col_sort(ent_t *e, int n, int lane) {
    int Pn = 32;
    while (Pn < n) Pn <<= 1;
    for (int size = 2; size <= Pn; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < (Pn >> 1); t += 32) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const ent_t a = e[lo], b = e[hi];
                if ((a > b) == up) { e[lo] = b; e[hi] = a; }
            }
        }
    }
    __syncwarp();
}

constexpr unsigned int COL_OVERFLOWED = 0x40000000u;   // sticky col_cnt value of a lost column

// Per column: keep the `cap` emitted rows with the smallest column keys (ties: lower row).
// One warp per column, bitonic sort of the packed entries in shared memory.
//   COMPACT = false (after the last row segment): cand_idx / overflow / col_tau out.
//   COMPACT = true (between row segments): the best cap entries go back to the head of the
//     column's buffer, col_cnt = cap, and tau_col tightens to the cap-th best key so far -- the
//     next segment emits ~cap * (its rows / rows seen so far) rows per column instead of
//     ~cap * (its rows / sample rows).  A column that overflowed keeps a sticky count.
constexpr int CS_WARPS = 4;
template <bool COMPACT>
__global__ void __launch_bounds__(CS_WARPS * 32)
col_select_kernel(ent_t *__restrict__ col_buf, unsigned int *__restrict__ col_cnt,
                  int64_t ny, int col_cap, int P2, int cap, int32_t *__restrict__ cand_idx,
                  int32_t *__restrict__ overflow, float *__restrict__ tau_col,
                  float *__restrict__ col_tau) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ent_t *e = reinterpret_cast<ent_t *>(smem_raw) + (size_t)warp * P2;
    const int64_t col = (int64_t)blockIdx.x * CS_WARPS + warp;
    if (col >= ny) return;
    const unsigned int total = col_cnt[col];
    const int n = (int)min(total, (unsigned int)col_cap);
    if constexpr (COMPACT) {
        if (total > (unsigned int)col_cap) {
            if (lane == 0) col_cnt[col] = COL_OVERFLOWED;
            return;
        }
        if (n < cap) return;                               // everything is kept, tau_col stands
    } else {
        if (lane == 0) overflow[col] = total > (unsigned int)col_cap;
    }
    for (int i = lane; i < P2; i += 32) e[i] = (i < n) ? col_buf[(size_t)col * col_cap + i] : EMPTY_ENTRY;
    col_sort(e, n, lane);
    if constexpr (COMPACT) {
        for (int p = lane; p < cap; p += 32) col_buf[(size_t)col * col_cap + p] = e[p];
        if (lane == 0) {
            col_cnt[col] = (unsigned int)cap;
            tau_col[col] = entry_key(e[cap - 1]);
        }
    } else {
        for (int p = lane; p < cap; p += 32) cand_idx[col * cap + p] = (p < n) ? entry_col(e[p]) : -1;
        // every row that was not kept has column key >= col_tau: emitted rows beyond the cap-th
        // best by the sort, rows never emitted by the threshold test (thresholds only tighten)
        if (col_tau && lane == 0) col_tau[col] = (n >= cap) ? entry_key(e[cap - 1]) : tau_col[col];
    }
}

// Multi-GPU exchanges of the row-sharded dual-direction pass (kiez_b200/distributed.py): every
// rank sees only its own rows, so the per-column state that must agree across ranks -- the
// thresholds between row segments, the best rows at the end -- travels as the HEAD of each
// column buffer: its first min(count, cap) entries (after kb2_col_compact: the best cap so far),
// padded with EMPTY_ENTRY.  out_ent: packed entries with row ids made global (+ row_offset);
// out_key: their fp32 keys (+inf padded).  Either may be NULL.
__global__ void __launch_bounds__(256)
col_heads_kernel(const ent_t *__restrict__ col_buf, const unsigned int *__restrict__ col_cnt,
                 int64_t ny, int col_cap, int cap, int64_t row_offset, ent_t *__restrict__ out_ent,
                 float *__restrict__ out_key) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ny * cap) return;
    const int64_t col = idx / cap;
    const int p = (int)(idx - col * cap);
    const unsigned int total = col_cnt[col];
    const int n = (int)min(min(total, (unsigned int)col_cap), (unsigned int)cap);
    ent_t e = EMPTY_ENTRY;
    if (p < n) e = col_buf[(size_t)col * col_cap + p] + (ent_t)row_offset;   // rows stay below 2^31
    if (out_ent) out_ent[idx] = e;
    if (out_key) out_key[idx] = entry_key(e);
}

// tau[col] = min(tau[col], kth smallest of the nparts x L keys gathered for the column): the
// kth best key over the union of what every rank has seen bounds the column's final kth best
// key from above.  One warp per column, bitonic sort over lanes x registers.
template <int R>
__global__ void __launch_bounds__(128)
kth_key_kernel(const float *__restrict__ keys, int nparts, int64_t part_stride, int64_t ny, int L,
               int kth, float *__restrict__ tau) {
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (col >= ny) return;
    const int total = nparts * L;
    ent_t x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = r * 32 + lane;
        float v = INFINITY;
        if (i < total) {
            const int part = i / L, j = i - part * L;
            v = keys[(size_t)part * part_stride + (size_t)col * L + j];
        }
        x[r] = pack_entry(v, 0);
    }
    warp_sort_entries<R>(x, lane);
    ent_t kth_ent = EMPTY_ENTRY;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const ent_t w = __shfl_sync(FULL_MASK, x[r], (kth - 1) & 31);
        if (r == ((kth - 1) >> 5)) kth_ent = w;
    }
    if (lane == 0) tau[col] = fminf(tau[col], entry_key(kth_ent));
}

static size_t fused_stage_bytes(int bk) { return (size_t)(2 * BM + 2 * F_HALF) * bk * 4; }

template <int BK>
static int launch_fused_cfg(const TcParams &P0, const FusedParams &FP, const float *q_hi,
                            const float *q_lo, const float *y_hi, const float *y_lo, int dpad,
                            int stages, int sm_count, int max_smem, cudaStream_t stream) {
    TcParams P = P0;
    P.stages = stages;
    P.kchunks = dpad / BK;
    P.per_split = ceil_div64(ceil_div64(P.ny, P.splits), F_BN) * F_BN;
    CUtensorMap mq_hi, mq_lo, my_hi, my_lo;
    if (make_map(&mq_hi, q_hi, P.nq, dpad, BM, BK)) return 1;
    if (make_map(&mq_lo, q_lo, P.nq, dpad, BM, BK)) return 1;
    if (make_map(&my_hi, y_hi, P.ny, dpad, F_HALF, BK)) return 1;
    if (make_map(&my_lo, y_lo, P.ny, dpad, F_HALF, BK)) return 1;
    const size_t need = stages * fused_stage_bytes(BK) + tc_fixed_smem(F_BN, P.cap, P.buf_slots, 8) +
                        EMIT_WORDS * sizeof(float);
    const size_t smem = min((size_t)max_smem, need + 1024);
    KB2_CUDA(cudaFuncSetAttribute(knn_fused_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t units = ((P.q_tiles + 1) / 2) * P.splits;
    const unsigned pairs = (unsigned)min((int64_t)(sm_count / 2), units);
    if (prepare_wave_sync(P, units, pairs, false, stream)) return 1;
    knn_fused_kernel<BK><<<2 * pairs, F_THREADS, smem, stream>>>(mq_hi, mq_lo, my_hi, my_lo, P, FP);
    KB2_LAUNCH_CHECK();
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_knn_fused(const float *x_hi, const float *x_lo, const float *x_key, int64_t nx,
                             const float *y_hi, const float *y_lo, const float *y_key, int64_t ny,
                             int dpad, int cap, int splits, const float *tau_col,
                             uint32_t *col_cnt, uint64_t *col_buf, int col_cap, int64_t row_id_base,
                             int32_t *cand_idx, void *stream) {
    KB2_CHECK(nx > 0 && ny > 0, "knn_fused: bad shape nx=%lld ny=%lld", (long long)nx, (long long)ny);
    KB2_CHECK(nx < (1LL << 31) - 256 && ny < (1LL << 31) - 256, "knn_fused: more than 2^31 rows");
    KB2_CHECK(dpad > 0 && dpad % 32 == 0, "knn_fused: dpad=%d must be a multiple of 32", dpad);
    KB2_CHECK(cap > 0 && cap <= 64, "knn_fused: cap=%d outside (0, 64]", cap);
    KB2_CHECK(splits >= 1 && (int64_t)splits * cap <= 2048, "knn_fused: splits*cap exceeds 2048");
    KB2_CHECK(col_cap >= cap && col_cap <= 4096, "knn_fused: col_cap=%d outside [cap, 4096]", col_cap);
    KB2_CHECK(row_id_base >= 0 && row_id_base + nx < (1LL << 31), "knn_fused: row ids exceed 2^31");
    int dev = 0, sm_count = 0, max_smem = 0;
    KB2_CUDA(cudaGetDevice(&dev));
    KB2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    KB2_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB2_CHECK(sm_count >= 2, "knn_fused: needs CTA pairs");
    TcParams P;
    P.nq = nx; P.ny = ny; P.kchunks = 0; P.cap = cap; P.splits = splits; P.stages = 0;
    P.per_split = 0; P.wave_sync = 0; P.q_tiles = ceil_div64(nx, BM); P.y_key = y_key; P.cand_idx = cand_idx;
    P.cand_key = nullptr;
    FusedParams FP;
    FP.x_key = x_key; FP.tau_col = tau_col; FP.col_cnt = col_cnt;
    FP.col_buf = reinterpret_cast<ent_t *>(col_buf); FP.col_cap = col_cap;
    FP.row_id_base = (int)row_id_base;
    auto stages_for = [&](int bk, int slots) {
        const size_t fixed = tc_fixed_smem(F_BN, cap, slots, 8) + EMIT_WORDS * sizeof(float);
        if (fixed >= (size_t)max_smem) return 0;
        return (int)min((size_t)MAX_STAGES, ((size_t)max_smem - fixed) / fused_stage_bytes(bk));
    };
    P.buf_slots = lists_buffer_slots(cap);
    while (P.buf_slots > LISTS_MIN_SLOTS && stages_for(32, P.buf_slots) < 2) P.buf_slots -= LISTS_GROUP;
    const int bk = stages_for(32, P.buf_slots) >= 2 ? 32 : 16;
    const int stages = stages_for(bk, P.buf_slots);
    KB2_CHECK(stages >= 2, "knn_fused: candidate lists of %d entries leave no room for 2 stages", cap);
    cudaStream_t st = (cudaStream_t)stream;
    if (bk == 32)
        return launch_fused_cfg<32>(P, FP, x_hi, x_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, st);
    return launch_fused_cfg<16>(P, FP, x_hi, x_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, st);
}

static int col_select_launch(bool compact, uint64_t *col_buf, uint32_t *col_cnt, int64_t ny,
                             int col_cap, int cap, int32_t *cand_idx, int32_t *overflow,
                             float *tau_col, float *col_tau, cudaStream_t stream) {
    const int P2 = max(32, next_pow2(col_cap));   // the sort network starts at 32 entries
    const size_t smem = (size_t)CS_WARPS * P2 * sizeof(ent_t);
    const unsigned grid = (unsigned)ceil_div64(ny, CS_WARPS);
    ent_t *buf = reinterpret_cast<ent_t *>(col_buf);
    if (compact) {
        KB2_CUDA(cudaFuncSetAttribute(col_select_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        col_select_kernel<true><<<grid, CS_WARPS * 32, smem, stream>>>(
            buf, col_cnt, ny, col_cap, P2, cap, cand_idx, overflow, tau_col, col_tau);
    } else {
        KB2_CUDA(cudaFuncSetAttribute(col_select_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        col_select_kernel<false><<<grid, CS_WARPS * 32, smem, stream>>>(
            buf, col_cnt, ny, col_cap, P2, cap, cand_idx, overflow, tau_col, col_tau);
    }
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_col_select(const uint64_t *col_buf, const uint32_t *col_cnt, int64_t ny,
                              int col_cap, int cap, int32_t *cand_idx, int32_t *overflow,
                              const float *tau_col, float *col_tau, void *stream) {
    KB2_CHECK(ny >= 0 && cap > 0 && col_cap >= cap && col_cap <= 4096, "col_select: bad arguments");
    KB2_CHECK(cand_idx && overflow, "col_select: cand_idx and overflow are required");
    KB2_CHECK(!col_tau || tau_col, "col_select: col_tau needs the tau_col the pass ran with");
    if (ny == 0) return 0;
    return col_select_launch(false, const_cast<uint64_t *>(col_buf), const_cast<uint32_t *>(col_cnt),
                             ny, col_cap, cap, cand_idx, overflow, const_cast<float *>(tau_col),
                             col_tau, (cudaStream_t)stream);
}

extern "C" int kb2_col_compact(uint64_t *col_buf, uint32_t *col_cnt, int64_t ny, int col_cap,
                               int cap, float *tau_col, void *stream) {
    KB2_CHECK(ny >= 0 && cap > 0 && col_cap >= cap && col_cap <= 4096, "col_compact: bad arguments");
    KB2_CHECK(col_buf && col_cnt && tau_col, "col_compact: NULL argument");
    if (ny == 0) return 0;
    return col_select_launch(true, col_buf, col_cnt, ny, col_cap, cap, nullptr, nullptr, tau_col,
                             nullptr, (cudaStream_t)stream);
}

extern "C" int kb2_col_heads(const uint64_t *col_buf, const uint32_t *col_cnt, int64_t ny,
                             int col_cap, int cap, int64_t row_offset, uint64_t *out_entries,
                             float *out_keys, void *stream) {
    KB2_CHECK(ny >= 0 && cap > 0 && col_cap >= cap && col_cap <= 4096, "col_heads: bad arguments");
    KB2_CHECK(col_buf && col_cnt && (out_entries || out_keys), "col_heads: NULL argument");
    KB2_CHECK(row_offset >= 0 && row_offset < (1LL << 31), "col_heads: row_offset out of range");
    if (ny == 0) return 0;
    const int64_t total = ny * cap;
    col_heads_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const ent_t *>(col_buf), col_cnt, ny, col_cap, cap, row_offset,
        reinterpret_cast<ent_t *>(out_entries), out_keys);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_kth_key(const float *keys, int nparts, int64_t part_stride, int64_t ny, int L,
                           int kth, float *tau, void *stream) {
    KB2_CHECK(keys && tau && nparts >= 1 && L >= 1 && ny >= 0, "kth_key: bad arguments");
    const int total = nparts * L;
    KB2_CHECK(total <= 1024, "kth_key: %d keys per column exceed 1024", total);
    KB2_CHECK(kth >= 1 && kth <= total, "kth_key: kth=%d outside [1, %d]", kth, total);
    if (ny == 0) return 0;
    const unsigned grid = (unsigned)ceil_div64(ny, 4);
    cudaStream_t st = (cudaStream_t)stream;
#define KB2_KTH(R) kth_key_kernel<R><<<grid, 128, 0, st>>>(keys, nparts, part_stride, ny, L, kth, tau)
    if (total <= 32) KB2_KTH(1);
    else if (total <= 64) KB2_KTH(2);
    else if (total <= 128) KB2_KTH(4);
    else if (total <= 256) KB2_KTH(8);
    else if (total <= 512) KB2_KTH(16);
    else KB2_KTH(32);
#undef KB2_KTH
    KB2_LAUNCH_CHECK();
    return 0;
}
