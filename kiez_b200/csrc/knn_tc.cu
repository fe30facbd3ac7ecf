// Candidate search on the 5th-gen tensor cores (sm_100a): the source x target
// contraction of kiez's exact kNN as 3xTF32 tcgen05.mma tiles fed by TMA, with the
// distance finish and the running top-`cap` selection fused into the epilogue, so
// the nq x ny matrix never leaves TMEM.
//
//   key[row][col] = y_key[col] - 2 * <q_row, y_col>,
//   <q, y> ~= q_lo.y_hi + q_hi.y_lo + q_hi.y_hi        (hi/lo: rn-TF32 split, prep.cu)
//
// This file: ONE CTA per (128-row query tile, index split) work unit, persistent over
// the units.  CTA = 6 warps:
//   warps 0-3  epilogue: tcgen05.ld their 32 TMEM lanes (one query row per thread),
//              key finish + threshold test + buffered append / merge by rank (select.cuh)
//   warp  4    TMA producer: one K chunk (BK features = one swizzle row) of
//              {q_hi, q_lo, y_hi, y_lo} per pipeline stage
//   warp  5    MMA issuer: 3 products x BK/8 UMMA_K steps per stage into one of two
//              TMEM accumulator buffers (128 lanes x BN fp32 columns each)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers
// (MMA <-> epilogue).  In this single-CTA form the tensor pipe tops out near 70 %
// busy: per 128-cycle MMA the SM reads 12 KB of operands from shared memory while TMA
// writes 8 KB of the next stage -- more than the 128 B/clk the banks deliver.  The
// CTA-pair kernel (knn_tc2.cu) halves the index-tile traffic per SM and is the default;
// this one serves shapes the pair kernel does not take and as its cross-check.
// Replaces the ArgKmin chunk loop of scikit-learn that
// kiez/neighbors/exact/sklearn_nearest_neighbors.py:96-101 calls.
#include "tc_common.cuh"

namespace kb2 {

template <int BN, int BK>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi,
              const __grid_constant__ CUtensorMap map_q_lo,
              const __grid_constant__ CUtensorMap map_y_hi,
              const __grid_constant__ CUtensorMap map_y_lo, const TcParams P) {
    using Cfg = StageCfg<BN, BK>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [stages x {q_hi, q_lo, y_hi, y_lo}] | y_key tile per epilogue warp [4][BN] |
    //        candidate lists | barriers
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
    uint64_t *full_bar = bars;                         // [stages]
    uint64_t *empty_bar = bars + MAX_STAGES;           // [stages]
    uint64_t *tmem_full = bars + 2 * MAX_STAGES;       // [2]
    uint64_t *tmem_empty = bars + 2 * MAX_STAGES + 2;  // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_STAGES + 4);
    {
        uint32_t dyn_size;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
        if (reinterpret_cast<unsigned char *>(tmem_slot + 4) > smem_raw + dyn_size) {
            if (threadIdx.x == 0)
                printf("kiez_b200: knn_tc shared-memory carve-up exceeds the launch size\n");
            __trap();
        }
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_units = P.q_tiles * P.splits;

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
            mbar_init(&tmem_empty[b], 4);   // one elected lane per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t stage_u32 = smem_u32(stage_base);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);

    if (warp == 4) {
        // ------------------------------------------------------ TMA producer
        // The whole warp runs the loop (warp-uniform values stay in uniform registers);
        // one elected lane issues the copies.
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int64_t qt = u % P.q_tiles;
            const int split = (int)(u / P.q_tiles);
            const int64_t y_begin = (int64_t)split * P.per_split;
            const int64_t y_end = min(P.ny, y_begin + P.per_split);
            const int q_row0 = (int)(qt * BM);
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                for (int kc = 0; kc < P.kchunks; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        const uint32_t st = stage_u32 + (uint32_t)stage * Cfg::STAGE_BYTES;
                        const uint32_t fb = full_u32 + (uint32_t)stage * 8;
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(&map_q_hi, st, fb, kc * BK, q_row0);
                        tma_load_2d(&map_q_lo, st + Cfg::A_BYTES, fb, kc * BK, q_row0);
                        tma_load_2d(&map_y_hi, st + 2 * Cfg::A_BYTES, fb, kc * BK, (int)c0);
                        tma_load_2d(&map_y_lo, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, fb, kc * BK,
                                    (int)c0);
                    }
                    __syncwarp();
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int split = (int)(u / P.q_tiles);
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
                        // small terms first, then hi*hi; K advances 32 B (UMMA_K tf32) per step
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_tf32(tmem_d, d_qlo + 2 * k, d_yhi + 2 * k, idesc, (kc | k) != 0);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_tf32(tmem_d, d_qhi + 2 * k, d_ylo + 2 * k, idesc, 1u);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_tf32(tmem_d, d_qhi + 2 * k, d_yhi + 2 * k, idesc, 1u);
                        umma_commit(empty_u32 + (uint32_t)stage * 8);   // slot free once these retire
                        if (kc == P.kchunks - 1)
                            umma_commit(smem_u32(&tmem_full[acc]));     // accumulator ready
                    }
                    __syncwarp();
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------ epilogue (warps 0-3)
        const int lrow = warp * 32 + lane;                 // TMEM lane == row within the tile
        float *yk = ykey_s + warp * BN;                    // this warp's private copy
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int64_t qt = u % P.q_tiles;
            const int split = (int)(u / P.q_tiles);
            const int64_t y_begin = (int64_t)split * P.per_split;
            const int64_t y_end = min(P.ny, y_begin + P.per_split);
            const int64_t grow = qt * BM + lrow;
            lists_reset(L, warp * 32, 32, lane);
            float tau = (grow < P.nq) ? INFINITY : -INFINITY;
            int cnt = 0;
            float ykreg[BN / 32];                          // next tile's terms, one tile ahead
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
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            lists_flush(L, lrow, tau, cnt, lane);
            write_lists(L, P, warp, lane, qt * BM, split);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"(TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------- host side
static size_t tc_stage_bytes(int bn, int bk) { return (size_t)(2 * BM + 2 * bn) * bk * 4; }

template <int BN, int BK>
static int launch_tc_cfg(const TcParams &P0, const float *q_hi, const float *q_lo,
                         const float *y_hi, const float *y_lo, int dpad, int stages, int sm_count,
                         int max_smem, cudaStream_t stream) {
    TcParams P = P0;
    P.stages = stages;
    P.kchunks = dpad / BK;
    P.per_split = ceil_div64(ceil_div64(P.ny, P.splits), BN) * BN;
    CUtensorMap mq_hi, mq_lo, my_hi, my_lo;
    if (make_map(&mq_hi, q_hi, P.nq, dpad, BM, BK)) return 1;
    if (make_map(&mq_lo, q_lo, P.nq, dpad, BM, BK)) return 1;
    if (make_map(&my_hi, y_hi, P.ny, dpad, BN, BK)) return 1;
    if (make_map(&my_lo, y_lo, P.ny, dpad, BN, BK)) return 1;
    // The kernel rounds its base up to 1024 B (the base is 1024-aligned in practice: there is
    // no static shared memory; the kernel traps if the carve-up overflows the launch size).
    const size_t need = stages * tc_stage_bytes(BN, BK) + tc_fixed_smem(BN, P.cap, P.buf_slots);
    const size_t smem = min((size_t)max_smem, need + 1024);
    KB2_CUDA(cudaFuncSetAttribute(knn_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t units = P.q_tiles * P.splits;
    const unsigned grid = (unsigned)min((int64_t)sm_count, units);
    knn_tc_kernel<BN, BK><<<grid, TC_THREADS, smem, stream>>>(mq_hi, mq_lo, my_hi, my_lo, P);
    KB2_LAUNCH_CHECK();
    return 0;
}

int launch_knn_tc(const float *q_hi, const float *q_lo, int64_t nq, const float *y_hi,
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
    auto stages_for = [&](int bn, int bk, int slots) {
        const size_t fixed = tc_fixed_smem(bn, cap, slots);
        if (fixed >= (size_t)max_smem) return 0;
        return (int)min((size_t)MAX_STAGES, ((size_t)max_smem - fixed) / tc_stage_bytes(bn, bk));
    };
    // Append-buffer slots per row: more slots = fewer merges; shrink towards the minimum
    // when the lists would otherwise squeeze the operand pipeline below 3 stages.
    P.buf_slots = lists_buffer_slots(cap);
    while (P.buf_slots > LISTS_MIN_SLOTS && stages_for(128, 16, P.buf_slots) < 3)
        P.buf_slots -= LISTS_GROUP;
    // Tile choice (measured): the widest index tile with 128-byte rows while two
    // stages fit, else 128-row tiles; KB2_TC_CONFIG=<BN>x<BK> overrides (tuning).
    int bn = 0, bk = 0;
    if (const char *env = getenv("KB2_TC_CONFIG")) {
        if (sscanf(env, "%dx%d", &bn, &bk) != 2 || (bn != 256 && bn != 128) || (bk != 32 && bk != 16)) {
            set_error("KB2_TC_CONFIG=%s: expected 256x32, 256x16, 128x32 or 128x16", env);
            return 1;
        }
    } else if (stages_for(256, 32, P.buf_slots) >= 2) {
        bn = 256; bk = 32;
    } else if (stages_for(128, 32, P.buf_slots) >= 2) {
        bn = 128; bk = 32;
    } else {
        bn = 128; bk = 16;
    }
    const int stages = stages_for(bn, bk, P.buf_slots);
    KB2_CHECK(stages >= 1, "knn_tc: candidate lists of %d entries do not fit in shared memory", cap);
    if (bn == 256 && bk == 32) return launch_tc_cfg<256, 32>(P, q_hi, q_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, stream);
    if (bn == 256 && bk == 16) return launch_tc_cfg<256, 16>(P, q_hi, q_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, stream);
    if (bn == 128 && bk == 32) return launch_tc_cfg<128, 32>(P, q_hi, q_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, stream);
    return launch_tc_cfg<128, 16>(P, q_hi, q_lo, y_hi, y_lo, dpad, stages, sm_count, max_smem, stream);
}

}  // namespace kb2
