// Screening candidate search: ONE tcgen05.mma.kind::tf32 product per K step (the `hi` halves
// of the 3xTF32 split only), query tile RESIDENT in shared memory (all of it, or its first
// `resident` K chunks when long candidate lists or wide rows need the space), CTA pairs, index
// ranges sized for L2 and chained per query tile.  One-direction (rows only) and dual-direction
// (rows + per-column emits, as knn_fused.cu) forms of the same kernel.
//
// Why it is still exact: the kernel only proposes candidates.  The exact finish
// (refine.cu, kb2_refine_topk_checked) recomputes their distances in float64 and PROVES, per
// query row, that no row outside the list can be among the k nearest:
//     every non-candidate has screen key >= tau (the list's cap-th best screen key), and
//     |screen key - exact key| <= E  (E from the TF32 rounding bound 2^-10 ||q|| ||y||),
//     so  exact k-th best distance  <  tau - E   =>   the top k is complete.
// Rows that fail the proof (a fraction ~1e-4 at C4) are searched again by the 3xTF32 kernel.
//
// Why it is fast: 3x fewer MMAs per distance, and the operand stream that holds the 3xTF32
// kernels at ~80 % tensor utilisation (64 KB per 1536-cycle stage = 42.7 B/clk/SM, the chip's
// L2 throughput cap) shrinks to the index half-tile alone: the 128 x dpad query tile (hi
// only: 128 KB at d = 256) stays in shared memory for the whole work unit, leaving
// 16 KB per 512-cycle stage = 32 B/clk/SM.
//
// L2 blocking: work unit = (query-tile pair, index range of `per_step` rows ~ 24 MB of
// operands), ordered range-major, so that all CTA pairs sweep the same L2-resident range;
// the candidate lists of a query tile are carried from range to range through global memory
// (`chained`: unit (qt, s) seeds its lists from the output of (qt, s-1), handshake through
// one flag per (query tile, epilogue warp)); the result is ONE list per row, as if the whole
// index had been swept in one unit.  Without chaining (few query tiles: parallelism comes from
// the ranges) every range writes its own list, like the `splits` of knn_tc2.cu.
//
// Protocol per unit: producer (warp PROD of both CTAs) waits q_empty, TMA-loads the resident
// K chunks of the query tile (bytes credited to CTA 0's q_full), then streams index half-tiles
// through the stage ring exactly like knn_tc2.cu; the MMA issuer (CTA 0) waits q_full once per
// unit and commits q_empty (multicast) after the unit's last MMA.
//
// Partial residency (cap > 32 at d = 256, cap > 64, dpad > 256): the ring holds uniform 16 KB
// slots, one TMA box each.  K chunk kc of an index tile takes one slot for the index half-tile
// and, when kc >= resident, a second one for the query tile's chunk kc (re-read from L2 for
// every index tile).  Per index tile the operand stream is (2 kchunks - resident) slots instead
// of kchunks: e.g. d = 256, cap = 56, resident = 5 -> 11 x 16 KB per 4096 tensor cycles =
// 43 B/clk/SM, where streaming the whole query tile (the 3xTF32 pair kernel with one product)
// would need 64 B/clk/SM against the ~43 B/clk/SM the L2 delivers.
#include "dual_common.cuh"

namespace kb2 {

constexpr int S_BN = 256;        // index rows per tile of the CTA pair
constexpr int S_HALF = 128;      // ... of which each CTA stages 128
constexpr int S_BK = 32;         // K chunk: 128-byte swizzle rows
constexpr int S_MAX_DPAD = 1024; // the resident part of the query tile adapts; eps_acc of the proof grows with dpad

struct ScreenParams {
    int64_t nq, ny;
    int kchunks, cap, buf_slots, stages;
    int resident;         // K chunks of the query tile kept in shared memory (<= kchunks)
    int rank_max;         // lists > 32: buffered entries up to which a merge goes by rank (select.cuh)
    int steps;            // index ranges
    int chained;          // 1: ranges of a query tile run in order and carry its lists
    int64_t per_step;     // index rows per range (multiple of S_BN)
    int64_t q_tiles;
    const float *y_key;
    int32_t *cand_idx;    // [nq][(chained ? 1 : steps) * cap]
    float *cand_key;      // same shape (required: it carries the lists and the proof's tau)
    int *chain_flag;      // [q_tiles * 4] zeroed before the launch (chained only)
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool DUAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DUAL ? 384 : 192, 1)
knn_screen_kernel(const __grid_constant__ CUtensorMap map_q,
                  const __grid_constant__ CUtensorMap map_y, const ScreenParams P,
                  const FusedParams FP) {
    using Cfg = StageCfg<S_HALF, S_BK>;           // A_BYTES = B_BYTES = 16 KB
    constexpr int BN = S_BN;
    constexpr int BK = S_BK;
    constexpr int EPI_WARPS = DUAL ? 8 : 4;
    constexpr int PROD_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *q_base = smem;                                         // [resident][128 x 32 fp32]
    unsigned char *stage_base = q_base + (size_t)P.resident * Cfg::A_BYTES; // [stages][128 x 32 fp32]
    // per epilogue warp a BN-float tile: warps 0-3 key_y, (DUAL) warps 4-7 tau_col
    float *tile_s = reinterpret_cast<float *>(stage_base + (size_t)P.stages * Cfg::B_BYTES);
    float *emit_key = tile_s + EPI_WARPS * BN;                            // DUAL: [4][EMIT_Q]
    int *emit_col = reinterpret_cast<int *>(emit_key + 4 * EMIT_Q);
    unsigned char *emit_lane = reinterpret_cast<unsigned char *>(emit_col + 4 * EMIT_Q);
    RowLists L;
    L.cap = P.cap;
    L.B = P.buf_slots;
    L.stride = lists_stride(P.cap, P.buf_slots);
    L.rank_max = P.rank_max;
    L.ent = reinterpret_cast<ent_t *>(tile_s + EPI_WARPS * BN + (DUAL ? EMIT_WORDS : 0));
    uint64_t *bars = reinterpret_cast<uint64_t *>(L.ent + (size_t)BM * L.stride);
    uint64_t *full_bar = bars;                         // [stages]   used in CTA 0
    uint64_t *empty_bar = bars + MAX_STAGES;           // [stages]   both CTAs
    uint64_t *tmem_full = bars + 2 * MAX_STAGES;       // [2]        both CTAs
    uint64_t *tmem_empty = bars + 2 * MAX_STAGES + 2;  // [2]        used in CTA 0
    uint64_t *q_full = bars + 2 * MAX_STAGES + 4;      //            used in CTA 0
    uint64_t *q_empty = bars + 2 * MAX_STAGES + 5;     //            both CTAs
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 6);
    {
        uint32_t dyn_size;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
        if (reinterpret_cast<unsigned char *>(tmem_slot + 4) > smem_raw + dyn_size) {
            if (threadIdx.x == 0)
                printf("kiez_b200: knn_screen shared-memory carve-up exceeds the launch size\n");
            __trap();
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs)
    const int64_t q_pairs = (P.q_tiles + 1) / 2;
    const int64_t num_units = q_pairs * P.steps;
    const int64_t pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == PROD_WARP && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_y);
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * EPI_WARPS);   // epilogue warps x 2 CTAs
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                 // barriers of both CTAs are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t q_u32 = smem_u32(q_base);
    const uint32_t stage_u32 = smem_u32(stage_base);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    const uint32_t tmem_empty_leader = smem_u32(tmem_empty) & PEER_BIT_MASK;

    if (warp >= EPI_WARPS) {
    if (warp == PROD_WARP) {
        // ------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0, qphase = 0;
        const uint32_t qf = smem_u32(q_full) & PEER_BIT_MASK;
        for (int64_t u = pair_id; u < num_units; u += num_pairs) {
            const int64_t qt = 2 * (u % q_pairs) + rank;
            const int step = (int)(u / q_pairs);
            const int64_t y_begin = (int64_t)step * P.per_step;
            const int64_t y_end = min(P.ny, y_begin + P.per_step);
            const int q_row0 = (int)(qt * BM);          // may lie past nq: TMA zero-fills
            // the resident part of this unit's query tile: its region is free once the previous
            // unit's MMAs retired
            if (P.resident > 0) {
                mbar_wait(q_empty, qphase ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(q_full, 2u * (uint32_t)P.resident * Cfg::A_BYTES);
                    for (int kc = 0; kc < P.resident; ++kc)
                        tma_load_2d_pair(&map_q, q_u32 + (uint32_t)kc * Cfg::A_BYTES, qf, kc * BK, q_row0);
                }
                __syncwarp();
                qphase ^= 1;
            }
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                const int y_row0 = (int)c0 + (int)rank * S_HALF;
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    // one slot for the index half-tile, one more for a streamed query chunk
                    const int loads = kc < P.resident ? 1 : 2;
                    for (int l = 0; l < loads; ++l) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one()) {
                            const uint32_t st = stage_u32 + (uint32_t)stage * Cfg::B_BYTES;
                            const uint32_t fb = (full_u32 + (uint32_t)stage * 8) & PEER_BIT_MASK;
                            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::B_BYTES);
                            if (l == 0) tma_load_2d_pair(&map_y, st, fb, kc * BK, y_row0);
                            else tma_load_2d_pair(&map_q, st, fb, kc * BK, q_row0);
                        }
                        __syncwarp();
                        if (++stage == P.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0, qphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int64_t u = pair_id; u < num_units; u += num_pairs) {
                const int step = (int)(u / q_pairs);
                const int64_t y_begin = (int64_t)step * P.per_step;
                const int64_t y_end = min(P.ny, y_begin + P.per_step);
                if (P.resident > 0) {
                    mbar_wait(q_full, qphase);
                    tc_fence_after();
                    qphase ^= 1;
                }
                for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                    const bool last_tile = c0 + BN >= y_end;
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                    for (int kc = 0; kc < P.kchunks; ++kc) {
                        const bool streamed = kc >= P.resident;
                        const int y_stage = stage;
                        mbar_wait(&full_bar[stage], phase);
                        if (++stage == P.stages) { stage = 0; phase ^= 1; }
                        const int q_stage = stage;
                        if (streamed) {
                            mbar_wait(&full_bar[stage], phase);
                            if (++stage == P.stages) { stage = 0; phase ^= 1; }
                        }
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t d_q = make_smem_desc<Cfg>(
                                streamed ? stage_u32 + (uint32_t)q_stage * Cfg::B_BYTES
                                         : q_u32 + (uint32_t)kc * Cfg::A_BYTES);
                            const uint64_t d_y = make_smem_desc<Cfg>(stage_u32 + (uint32_t)y_stage * Cfg::B_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_tf32_pair(tmem_d, d_q + 2 * k, d_y + 2 * k, idesc, (kc | k) != 0);
                            umma_commit_pair(empty_u32 + (uint32_t)y_stage * 8, 0x3);
                            if (streamed) umma_commit_pair(empty_u32 + (uint32_t)q_stage * 8, 0x3);
                            if (kc == P.kchunks - 1) {
                                umma_commit_pair(smem_u32(&tmem_full[acc]), 0x3);
                                if (last_tile && P.resident > 0) umma_commit_pair(smem_u32(q_empty), 0x3);
                            }
                        }
                        __syncwarp();
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    }
    } else {
    if (warp < 4) {
        // ------------------------------------------------------ row epilogue, both CTAs
        const int lrow = warp * 32 + lane;
        float *yk = tile_s + warp * BN;
        const int out_ld = (P.chained ? 1 : P.steps) * P.cap;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t u = pair_id; u < num_units; u += num_pairs) {
            const int64_t qt = 2 * (u % q_pairs) + rank;
            const int step = (int)(u / q_pairs);
            const int64_t y_begin = (int64_t)step * P.per_step;
            const int64_t y_end = min(P.ny, y_begin + P.per_step);
            const int64_t row0 = qt * BM + warp * 32;       // first row of this warp
            const int64_t grow = qt * BM + lrow;
            float tau = (grow < P.nq) ? INFINITY : -INFINITY;
            int cnt = 0;
            if (P.chained && step > 0 && row0 < P.nq) {
                // seed from the lists the previous range of this query tile left in global memory
                int *flag = P.chain_flag + qt * 4 + warp;
                if (lane == 0) {
                    long long t0 = 0;
                    for (uint32_t spin = 1; ld_acquire(flag) < step; ++spin) {
                        __nanosleep(64);
                        if ((spin & 1023u) == 0) {
                            const long long now = clock64();
                            if (t0 == 0) t0 = now;
                            if (now - t0 > 8000000000LL) {
                                printf("kiez_b200: knn_screen chain wait timed out (block %d)\n", blockIdx.x);
                                __trap();
                            }
                        }
                    }
                }
                __syncwarp();
                // 16 rows per batch: their 32 loads are issued before the first store.  (Written
                // row by row the compiler kept load -> generic store -> load: one L2 round trip
                // per row and 32-entry slice, 25-50 k cycles at the start of every unit while the
                // MMA issuer waited for the first accumulator to be consumed.)
                constexpr int SEED_ROWS = 16;
                for (int p = lane; p < P.cap; p += 32) {
                    for (int r0 = 0; r0 < 32; r0 += SEED_ROWS) {
                        float sk[SEED_ROWS];
                        int si[SEED_ROWS];
#pragma unroll
                        for (int j = 0; j < SEED_ROWS; ++j) {
                            const int64_t gr = row0 + r0 + j;
                            const bool ok = gr < P.nq;
                            sk[j] = ok ? __ldcg(P.cand_key + gr * P.cap + p) : INFINITY;
                            si[j] = ok ? __ldcg(P.cand_idx + gr * P.cap + p) : -1;
                        }
#pragma unroll
                        for (int j = 0; j < SEED_ROWS; ++j)      // pack(+inf, -1) == EMPTY_ENTRY
                            L.ent[(size_t)(warp * 32 + r0 + j) * L.stride + p] = pack_entry(sk[j], si[j]);
                    }
                }
                __syncwarp();
                if (grow < P.nq) tau = entry_key(L.ent[(size_t)lrow * L.stride + P.cap - 1]);
            } else {
                lists_reset(L, warp * 32, 32, lane);
            }
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
                epilogue_tile<BN, !DUAL>(L, lrow, yk, taddr, c0, tau, cnt, lane);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + (uint32_t)acc * 8);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            lists_flush(L, lrow, tau, cnt, lane);
            // write the 32 lists of this warp (sorted ascending, +inf / -1 padded)
            for (int r = 0; r < 32; ++r) {
                const int64_t gr = row0 + r;
                if (gr >= P.nq) break;
                const ent_t *e = L.ent + (size_t)(warp * 32 + r) * L.stride;
                for (int p = lane; p < P.cap; p += 32) {
                    const int64_t o = gr * out_ld + (P.chained ? 0 : (int64_t)step * P.cap) + p;
                    P.cand_idx[o] = entry_col(e[p]);
                    P.cand_key[o] = entry_key(e[p]);
                }
            }
            if (P.chained && row0 < P.nq) {
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release(P.chain_flag + qt * 4 + warp, step + 1);
            }
            __syncwarp();
        }
    } else if (DUAL) {
        // ------------------------------------------------------ column epilogue, both CTAs
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
            const int step = (int)(u / q_pairs);
            const int64_t y_begin = (int64_t)step * P.per_step;
            const int64_t y_end = min(P.ny, y_begin + P.per_step);
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
    }

    // no CTA may exit (or free TMEM) while its peer can still signal its barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(TMEM_COLS)
                     : "memory");
    }
}

// Shared memory next to the resident query tile and the stages.
static size_t screen_fixed_smem(int cap, int slots, bool dual) {
    return (size_t)(dual ? 8 : 4) * S_BN * sizeof(float) + (dual ? EMIT_WORDS * sizeof(float) : 0) +
           lists_bytes(BM, cap, slots) + (2 * MAX_STAGES + 6) * 8 + 16;
}

// Ring slots the shape gets (0: the screen kernel does not take it), choosing the append-buffer
// size and how many K chunks of the query tile stay resident.  Shared memory left by the lists
// is cut into 16 KB units; at least SCREEN_MIN_STAGES of them form the ring, the rest hold
// query chunks.  A fully resident tile gets the remaining units as extra ring slots.
//
// `ny` (index rows one list sweeps; 0 = long index): a row accepts ~cap ln(ny / cap) entries
// whatever the key precision, and every `slots - 4` of them cost one warp-wide list merge.  Over
// a long index (C4: 1 M rows) that is noise next to the 4096 tensor cycles of an index tile and
// the shared memory is better spent on residency; over a short one with long lists (C3: cap 112,
// 100 k rows; C2: cap 56, 15 k rows) the merges ARE the kernel (C3 ran 39 k cycles per tile with
// 12-slot buffers), so the append buffers grow -- at the price of resident query chunks -- until
// the estimated merge work per tile drops below a quarter of the tile's tensor time.
constexpr int SCREEN_MIN_STAGES = 4;
static int screen_config(int dpad, int cap, bool dual, int max_smem, int *slots_out,
                         int *resident_out = nullptr, int64_t ny = 0) {
    if (dpad <= 0 || dpad % S_BK != 0 || dpad > S_MAX_DPAD || cap <= 0 || cap > 128) return 0;
    const int kchunks = dpad / S_BK;
    const size_t unit = (size_t)S_HALF * S_BK * 4;
    auto units_for = [&](int slots) {
        const size_t fixed = screen_fixed_smem(cap, slots, dual) + 1024;
        if (fixed >= (size_t)max_smem) return 0;
        return (int)(((size_t)max_smem - fixed) / unit);
    };
    // longer append buffers mean fewer list merges, but every 16 KB they cost is a resident
    // query chunk less: shrink them while that buys residency (or the minimum ring)
    int slots = lists_buffer_slots(cap);
    for (int cand = slots - LISTS_GROUP; cand >= LISTS_MIN_SLOTS; cand -= LISTS_GROUP)
        if (min(kchunks + SCREEN_MIN_STAGES, units_for(cand)) >
            min(kchunks + SCREEN_MIN_STAGES, units_for(slots)))
            slots = cand;
    if (ny > 0) {
        const int regs = cap <= 16 ? 1 : cap <= 32 ? 2 : cap <= 64 ? 4 : 8;      // merge cost class
        const int max_slots = cap <= 16 ? 16 : cap <= 32 ? 32 : 64;
        const double appends = cap * log(fmax(2.0, (double)ny / cap));            // per row
        const double merge_instr = 150.0 + 90.0 * regs;                           // warp-wide, per merge
        const double tiles = fmax(1.0, (double)ny / S_BN);
        auto merge_load = [&](int sl) {      // warp instructions per index tile spent merging
            return 32.0 * appends / (sl - LISTS_GROUP) * merge_instr / tiles;
        };
        while (merge_load(slots) > 1024.0 && slots + LISTS_GROUP <= max_slots &&
               units_for(slots + LISTS_GROUP) >= SCREEN_MIN_STAGES)
            slots += LISTS_GROUP;
    }
    if (const char *env = getenv("KB2_SCREEN_SLOTS")) {                           // tuning / A-B only
        const int sl = atoi(env) / LISTS_GROUP * LISTS_GROUP;
        if (sl >= LISTS_MIN_SLOTS && sl <= 64 && units_for(sl) >= SCREEN_MIN_STAGES) slots = sl;
    }
    const int units = units_for(slots);
    int resident = min(kchunks, units - SCREEN_MIN_STAGES);
    if (resident < 0) resident = 0;
    if (const char *env = getenv("KB2_SCREEN_RESIDENT")) {
        const int r = atoi(env);
        if (r >= 0 && r < resident) resident = r;
    }
    const int stages = min(MAX_STAGES, units - resident);
    if (slots_out) *slots_out = slots;
    if (resident_out) *resident_out = resident;
    // a streamed chunk occupies two slots: never run with fewer than 3
    return stages >= 3 ? stages : 0;
}

template <bool DUAL>
static int launch_screen(ScreenParams P, const FusedParams &FP, const float *q_hi, const float *y_hi,
                         int dpad, int sm_count, int max_smem, cudaStream_t stream) {
    CUtensorMap mq, my;
    if (make_map(&mq, q_hi, P.nq, dpad, BM, S_BK)) return 1;
    if (make_map(&my, y_hi, P.ny, dpad, S_HALF, S_BK)) return 1;
    const size_t need = (size_t)(P.resident + P.stages) * S_HALF * S_BK * 4 +
                        screen_fixed_smem(P.cap, P.buf_slots, DUAL);
    const size_t smem = min((size_t)max_smem, need + 1024);
    KB2_CUDA(cudaFuncSetAttribute(knn_screen_kernel<DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t units = ((P.q_tiles + 1) / 2) * P.steps;
    const unsigned pairs = (unsigned)min((int64_t)(sm_count / 2), units);
    if (P.chained)
        KB2_CUDA(cudaMemsetAsync(P.chain_flag, 0, (size_t)P.q_tiles * 4 * sizeof(int), stream));
    knn_screen_kernel<DUAL><<<2 * pairs, DUAL ? 384 : 192, smem, stream>>>(mq, my, P, FP);
    KB2_LAUNCH_CHECK();
    return 0;
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_screen_stages(int dpad, int cap, int dual) {
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return 0;
    return screen_config(dpad, cap, dual != 0, max_smem, nullptr);
}

extern "C" int kb2_screen_config(int dpad, int cap, int dual, int max_smem, int64_t ny, int *slots,
                                 int *resident) {
    if (max_smem <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
            return 0;
    }
    return screen_config(dpad, cap, dual != 0, max_smem, slots, resident, ny);
}

extern "C" int kb2_screen_plan(int64_t nq, int64_t ny, int dpad, int cap, int sm_count,
                               int *steps, int *chained) {
    KB2_CHECK(steps && chained, "screen_plan: NULL output");
    KB2_CHECK(nq > 0 && ny > 0 && dpad > 0 && cap > 0, "screen_plan: bad shape");
    if (sm_count <= 0) sm_count = 148;
    const int64_t q_tiles = (nq + BM - 1) / BM;
    if (q_tiles >= 2 * (int64_t)sm_count) {
        // enough query tiles to fill the GPU: ranges exist for L2 only and are chained
        double mb = 24.0;
        if (const char *env = getenv("KB2_SCREEN_RANGE_MB")) mb = atof(env);
        if (!(mb >= 0.25)) mb = 0.25;
        int64_t rows = (int64_t)(mb * 1048576.0 / ((double)dpad * 4.0));
        rows = (rows + S_BN - 1) / S_BN * S_BN;
        if (rows < 4 * S_BN) rows = 4 * S_BN;
        int64_t s = (ny + rows - 1) / rows;
        if (s > 32768) s = 32768;
        *steps = (int)(s < 1 ? 1 : s);
        *chained = 1;
    } else {
        *steps = kb2_suggest_splits(nq, ny, cap, sm_count);
        *chained = 0;
    }
    return 0;
}

extern "C" int kb2_knn_screen(const float *q_hi, const float *q_key, int64_t nq, const float *y_hi,
                              const float *y_key, int64_t ny, int dpad, int cap, int steps,
                              int chained, int32_t *cand_idx, float *cand_key, int32_t *chain_flag,
                              const float *tau_col, uint32_t *col_cnt, uint64_t *col_buf,
                              int col_cap, int64_t row_id_base, void *stream) {
    KB2_CHECK(nq > 0 && ny > 0, "knn_screen: bad shape nq=%lld ny=%lld", (long long)nq, (long long)ny);
    KB2_CHECK(nq < (1LL << 31) - 256 && ny < (1LL << 31) - 256, "knn_screen: more than 2^31 rows");
    KB2_CHECK(steps >= 1 && (chained || (int64_t)steps * cap <= 2048),
              "knn_screen: steps=%d with cap=%d exceeds 2048 candidates per row", steps, cap);
    KB2_CHECK(cand_idx && cand_key, "knn_screen: cand_idx and cand_key are required");
    KB2_CHECK(!chained || chain_flag, "knn_screen: chained ranges need chain_flag");
    const bool dual = tau_col != nullptr;
    KB2_CHECK(!dual || (q_key && col_cnt && col_buf && col_cap >= cap && col_cap <= 4096),
              "knn_screen: the dual-direction form needs q_key, col_cnt, col_buf and col_cap in [cap, 4096]");
    KB2_CHECK(row_id_base >= 0 && row_id_base + nq < (1LL << 31), "knn_screen: row ids exceed 2^31");
    int dev = 0, sm_count = 0, max_smem = 0;
    KB2_CUDA(cudaGetDevice(&dev));
    KB2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    KB2_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB2_CHECK(sm_count >= 2, "knn_screen: needs CTA pairs");
    ScreenParams P;
    P.nq = nq; P.ny = ny; P.kchunks = dpad / S_BK; P.cap = cap; P.steps = steps;
    P.chained = chained ? 1 : 0;
    P.per_step = ceil_div64(ceil_div64(ny, steps), S_BN) * S_BN;
    P.q_tiles = ceil_div64(nq, BM); P.y_key = y_key; P.cand_idx = cand_idx; P.cand_key = cand_key;
    P.chain_flag = chain_flag;
    // rows one candidate list sweeps: the whole index when the ranges are chained, else one range
    P.stages = screen_config(dpad, cap, dual, max_smem, &P.buf_slots, &P.resident,
                             chained ? ny : P.per_step);
    KB2_CHECK(P.stages > 0, "knn_screen: dpad=%d cap=%d is not supported (dpad <= %d, multiple of "
              "%d, cap <= 128; see kb2_screen_stages)", dpad, cap, S_MAX_DPAD, S_BK);
    P.rank_max = LISTS_RANK_MAX_CNT;
    if (const char *env = getenv("KB2_RANK_MAX_CNT")) {                           // tuning / A-B only
        const int r = atoi(env);
        if (r >= 0 && r <= 64) P.rank_max = r;
    }
    if (const char *env = getenv("KB2_SCREEN_STAGES")) {
        const int s = atoi(env);
        if (s >= 3 && s <= P.stages) P.stages = s;
    }
    FusedParams FP;
    FP.x_key = q_key; FP.tau_col = tau_col; FP.col_cnt = col_cnt;
    FP.col_buf = reinterpret_cast<ent_t *>(col_buf); FP.col_cap = col_cap;
    FP.row_id_base = (int)row_id_base;
    cudaStream_t st = (cudaStream_t)stream;
    if (dual) return launch_screen<true>(P, FP, q_hi, y_hi, dpad, sm_count, max_smem, st);
    return launch_screen<false>(P, FP, q_hi, y_hi, dpad, sm_count, max_smem, st);
}
