// Candidate search, CTA-pair form (tcgen05 cta_group::2): two CTAs on the two SMs of a TPC
// share one 256-row index tile and each own a 128-row query tile.  One
// tcgen05.mma.cta_group::2 computes M = 256 (128 query rows from each CTA's shared memory)
// x N = 256 (128 index rows from each CTA's shared memory); each CTA's TMEM receives its own
// 128 rows x 256 columns, so the epilogue (key finish + fused top-`cap` selection) is the
// same per-CTA code as in knn_tc.cu.
//
// Why: in the single-CTA form every SM stages the whole index tile, and operand reads
// (12 KB per 128-cycle MMA) plus TMA writes exceed the 128 B/clk shared-memory bandwidth,
// capping the tensor pipe near 70 % busy.  Here each SM stages and reads only half of
// the index tile: 8 KB of operand reads per MMA and 64 KB (not 96 KB) of TMA writes per
// K chunk, which also leaves room for a third pipeline stage.
//
// Protocol (per K chunk / stage s, per accumulator buffer b):
//   producers (warp 4 of BOTH CTAs): wait own empty[s]; TMA {q_hi, q_lo} (own 128 query rows)
//       and {y_hi, y_lo} (own half of the index tile) with the transaction bytes credited to
//       CTA 0's full[s]; CTA 0 arms full[s] with the bytes of both CTAs.
//   MMA issuer (warp 5 of CTA 0 only): wait tmem_empty[b] (8 arrivals: 4 epilogue warps x 2
//       CTAs) and full[s]; 3 products x BK/8 steps; tcgen05.commit multicast -> empty[s] of
//       both CTAs, and after the last chunk -> tmem_full[b] of both CTAs.
//   epilogue (warps 0-3 of BOTH CTAs): wait own tmem_full[b]; tcgen05.ld own TMEM; arrive on
//       CTA 0's tmem_empty[b].
#include "tc_common.cuh"

namespace kb2 {

constexpr int PAIR_BN = 256;     // index rows per tile of the CTA pair
constexpr int HALF_BN = 128;     // ... of which each CTA stages 128

template <int BK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
knn_tc2_kernel(const __grid_constant__ CUtensorMap map_q_hi,
               const __grid_constant__ CUtensorMap map_q_lo,
               const __grid_constant__ CUtensorMap map_y_hi,
               const __grid_constant__ CUtensorMap map_y_lo, const TcParams P) {
    using Cfg = StageCfg<HALF_BN, BK>;
    constexpr int BN = PAIR_BN;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    float *ykey_s = reinterpret_cast<float *>(stage_base + (size_t)P.stages * Cfg::STAGE_BYTES);
    RowLists L;
    L.cap = P.cap;
    L.B = P.buf_slots;
    L.stride = lists_stride(P.cap, P.buf_slots);
    L.ent = reinterpret_cast<ent_t *>(ykey_s + 4 * BN);
    uint64_t *bars = reinterpret_cast<uint64_t *>(L.ent + (size_t)BM * L.stride);
    uint64_t *full_bar = bars;                         // [stages]   used in CTA 0
    uint64_t *empty_bar = bars + MAX_STAGES;           // [stages]   both CTAs
    uint64_t *tmem_full = bars + 2 * MAX_STAGES;       // [2]        both CTAs
    uint64_t *tmem_empty = bars + 2 * MAX_STAGES + 2;  // [2]        used in CTA 0
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);
    {
        uint32_t dyn_size;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
        if (reinterpret_cast<unsigned char *>(tmem_slot + 4) > smem_raw + dyn_size) {
            if (threadIdx.x == 0)
                printf("kiez_b200: knn_tc2 shared-memory carve-up exceeds the launch size\n");
            __trap();
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();           // 0 = leader (issues the MMAs)
    const int64_t q_pairs = (P.q_tiles + 1) / 2;
    const int64_t num_units = q_pairs * P.splits;       // units of the pair
    const int64_t pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 4 && lane == 0) {
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
            mbar_init(&tmem_empty[b], 8);   // 4 epilogue warps x 2 CTAs
        }
        fence_barrier_init();
    }
    if (warp == 5) {
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
    const uint32_t stage_u32 = smem_u32(stage_base);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);

    if (warp == 4) {
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
            const int q_row0 = (int)(qt * BM);          // may lie past nq: TMA zero-fills
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                const int y_row0 = (int)c0 + (int)rank * HALF_BN;
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
    } else if (warp == 5) {
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
    } else {
        // ------------------------------------------------------ epilogue (warps 0-3, both CTAs)
        const int lrow = warp * 32 + lane;
        float *yk = ykey_s + warp * BN;
        const uint32_t tmem_empty_leader = smem_u32(tmem_empty) & PEER_BIT_MASK;
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
                epilogue_tile<BN>(L, lrow, yk, taddr, c0, tau, cnt, lane);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + (uint32_t)acc * 8);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            lists_flush(L, lrow, tau, cnt, lane);
            write_lists(L, P, warp, lane, qt * BM, split);
        }
    }

    // no CTA may exit (or free TMEM) while its peer can still signal its barriers
    tc_fence_before();
    cluster_sync_all();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(TMEM_COLS)
                     : "memory");
    }
}

static size_t tc2_stage_bytes(int bk) { return (size_t)(2 * BM + 2 * HALF_BN) * bk * 4; }

template <int BK>
static int launch_tc2_cfg(const TcParams &P0, const float *q_hi, const float *q_lo,
                          const float *y_hi, const float *y_lo, int dpad, int stages, int sm_count,
                          int max_smem, cudaStream_t stream) {
    TcParams P = P0;
    P.stages = stages;
    P.kchunks = dpad / BK;
    P.per_split = ceil_div64(ceil_div64(P.ny, P.splits), PAIR_BN) * PAIR_BN;
    CUtensorMap mq_hi, mq_lo, my_hi, my_lo;
    if (make_map(&mq_hi, q_hi, P.nq, dpad, BM, BK)) return 1;
    if (make_map(&mq_lo, q_lo, P.nq, dpad, BM, BK)) return 1;
    if (make_map(&my_hi, y_hi, P.ny, dpad, HALF_BN, BK)) return 1;
    if (make_map(&my_lo, y_lo, P.ny, dpad, HALF_BN, BK)) return 1;
    const size_t need = stages * tc2_stage_bytes(BK) + tc_fixed_smem(PAIR_BN, P.cap, P.buf_slots);
    const size_t smem = min((size_t)max_smem, need + 1024);
    KB2_CUDA(cudaFuncSetAttribute(knn_tc2_kernel<BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t units = ((P.q_tiles + 1) / 2) * P.splits;
    const unsigned pairs = (unsigned)min((int64_t)(sm_count / 2), units);
    if (prepare_wave_sync(P, units, pairs, true, stream)) return 1;
    knn_tc2_kernel<BK><<<2 * pairs, TC_THREADS, smem, stream>>>(mq_hi, mq_lo, my_hi, my_lo, P);
    KB2_LAUNCH_CHECK();
    return 0;
}

// Returns -1 when the pair kernel does not take the shape (caller falls back to knn_tc).
int launch_knn_tc2(const float *q_hi, const float *q_lo, int64_t nq, const float *y_hi,
                   const float *y_lo, const float *y_key, int64_t ny, int dpad, int cap, int splits,
                   int32_t *cand_idx, float *cand_key, cudaStream_t stream) {
    int dev = 0, sm_count = 0, max_smem = 0;
    KB2_CUDA(cudaGetDevice(&dev));
    KB2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    KB2_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    TcParams P;
    P.nq = nq; P.ny = ny; P.kchunks = 0; P.cap = cap; P.splits = splits; P.stages = 0;
    P.per_split = 0; P.wave_sync = 0;
    P.q_tiles = ceil_div64(nq, BM); P.y_key = y_key; P.cand_idx = cand_idx; P.cand_key = cand_key;
    auto stages_for = [&](int bk, int slots) {
        const size_t fixed = tc_fixed_smem(PAIR_BN, cap, slots);
        if (fixed >= (size_t)max_smem) return 0;
        return (int)min((size_t)MAX_STAGES, ((size_t)max_smem - fixed) / tc2_stage_bytes(bk));
    };
    P.buf_slots = lists_buffer_slots(cap);
    while (P.buf_slots > LISTS_MIN_SLOTS && stages_for(16, P.buf_slots) < 3) P.buf_slots -= LISTS_GROUP;
    int bk = 0;
    if (const char *env = getenv("KB2_TC2_BK")) {
        bk = atoi(env);
        if (bk != 32 && bk != 16) {
            set_error("KB2_TC2_BK=%s: expected 32 or 16", env);
            return 1;
        }
    } else {
        bk = stages_for(32, P.buf_slots) >= 2 ? 32 : 16;   // measured: 2 x 64 KB beats 4 x 32 KB
    }
    const int stages = stages_for(bk, P.buf_slots);
    if (stages < 2 || sm_count < 2) return -1;
    if (bk == 32) return launch_tc2_cfg<32>(P, q_hi, q_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, stream);
    return launch_tc2_cfg<16>(P, q_hi, q_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, stream);
}

}  // namespace kb2
