// Candidate search on the 5th-gen tensor cores (sm_100a): the source x target
// contraction of kiez's exact kNN as 3xTF32 tcgen05.mma tiles fed by TMA, with the
// distance finish and the running top-`cap` selection fused into the epilogue, so
// the nq x ny matrix never leaves TMEM.
//
//   key[row][col] = y_key[col] - 2 * <q_row, y_col>,
//   <q, y> ~= q_hi.y_hi + q_hi.y_lo + q_lo.y_hi        (hi/lo: rn-TF32 split, prep.cu)
//
// CTA = 6 warps, persistent over work units (query tile of 128 rows, index split):
//   warps 0-3  epilogue: tcgen05.ld their 32 TMEM lanes (one query row per thread),
//              key finish + threshold test + cooperative sorted insert (select.cuh)
//   warp  4    TMA producer: one K-chunk (32 features = one 128 B swizzle row) of
//              {q_hi, q_lo, y_hi, y_lo} per pipeline stage
//   warp  5    MMA issuer: 3 products x 4 UMMA_K=8 steps per stage into one of two
//              TMEM accumulator buffers (128 lanes x BN fp32 columns each)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers
// (MMA <-> epilogue).  Replaces the ArgKmin chunk loop of scikit-learn that
// kiez/neighbors/exact/sklearn_nearest_neighbors.py:96-101 calls.
#include <cuda.h>

#include "select.cuh"

namespace kb2 {

constexpr int BM = 128;          // query rows per tile  (UMMA M, TMEM lanes)
constexpr int BK = 32;           // fp32 features per stage row = 128 B (SWIZZLE_128B span)
constexpr int UMMA_K = 8;        // tf32: 32 B of K per instruction
constexpr int TC_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr uint32_t TMEM_COLS = 512;

template <int BN>
struct TileCfg {
    static constexpr int A_BYTES = BM * BK * 4;                 // 16 KB per operand half
    static constexpr int B_BYTES = BN * BK * 4;                 // 32 KB (BN=256) / 16 KB
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int NUM_ACC = 2;                           // TMEM accumulator buffers
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails loudly) instead of
// hanging the GPU.  ~4 s at 2 GHz; a healthy wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            printf("kiez_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x,
                   threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, void *dst, uint64_t *bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate, single CTA.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread retire
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor
// layout): start>>4 | LBO(unused for swizzled K-major, 1)<<16 | SBO=1024 B (8 rows x
// 128 B)<<32 | version 1<<46 | layout_type SWIZZLE_128B (2)<<61.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1)<<4,
// a_format/b_format TF32 (2)<<7/<<10, A and B K-major (0), N>>3 <<17, M>>4 <<24.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
    int64_t nq, ny;
    int kchunks;          // dpad / BK
    int cap, buf_slots, splits, stages;
    int exclude_self;
    int64_t self_offset;
    int64_t per_split;    // index rows per split (multiple of BN)
    int64_t q_tiles;
    const float *y_key;
    int32_t *cand_idx;
    float *cand_key;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi,
              const __grid_constant__ CUtensorMap map_q_lo,
              const __grid_constant__ CUtensorMap map_y_hi,
              const __grid_constant__ CUtensorMap map_y_lo, const TcParams P) {
    using Cfg = TileCfg<BN>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [stages x {q_hi, q_lo, y_hi, y_lo}] | y_key tiles [2][BN] | lists | barriers
    unsigned char *smem = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *stage_base = smem;
    float *ykey_s = reinterpret_cast<float *>(stage_base + (size_t)P.stages * Cfg::STAGE_BYTES);
    RowLists L;
    L.cap = P.cap;
    L.B = P.buf_slots;
    L.stride = lists_stride(P.cap, P.buf_slots);
    L.ent = reinterpret_cast<ent_t *>(ykey_s + 2 * BN);
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

    if (warp == 4) {
        // ------------------------------------------------------ TMA producer
        if (lane == 0) {
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
                        unsigned char *st = stage_base + (size_t)stage * Cfg::STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        tma_load_2d(&map_q_hi, st, &full_bar[stage], kc * BK, q_row0);
                        tma_load_2d(&map_q_lo, st + Cfg::A_BYTES, &full_bar[stage], kc * BK, q_row0);
                        tma_load_2d(&map_y_hi, st + 2 * Cfg::A_BYTES, &full_bar[stage], kc * BK,
                                    (int)c0);
                        tma_load_2d(&map_y_lo, st + 2 * Cfg::A_BYTES + Cfg::B_BYTES,
                                    &full_bar[stage], kc * BK, (int)c0);
                        if (++stage == P.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ------------------------------------------------------ MMA issuer
        if (lane == 0) {
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
                        const uint32_t sa = smem_u32(stage_base + (size_t)stage * Cfg::STAGE_BYTES);
                        const uint64_t d_qhi = make_smem_desc(sa);
                        const uint64_t d_qlo = make_smem_desc(sa + Cfg::A_BYTES);
                        const uint64_t d_yhi = make_smem_desc(sa + 2 * Cfg::A_BYTES);
                        const uint64_t d_ylo = make_smem_desc(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                        // small terms first, then hi*hi
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
                            umma_tf32(tmem_d, d_qlo + adv, d_yhi + adv, idesc, (kc | k) != 0);
                        }
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
                            umma_tf32(tmem_d, d_qhi + adv, d_ylo + adv, idesc, 1u);
                        }
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
                            umma_tf32(tmem_d, d_qhi + adv, d_yhi + adv, idesc, 1u);
                        }
                        umma_commit(&empty_bar[stage]);   // smem slot free once these retire
                        if (++stage == P.stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[acc]);          // accumulator ready for the epilogue
                    if (++acc == Cfg::NUM_ACC) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------ epilogue (warps 0-3)
        const int lrow = warp * 32 + lane;                 // TMEM lane == row within the tile
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
            for (int64_t c0 = y_begin; c0 < y_end; c0 += BN) {
                // stage this tile's selection terms (double-buffered by accumulator slot)
                float *yk = ykey_s + acc * BN;
                for (int j = threadIdx.x; j < BN; j += 128) {
                    const int64_t col = c0 + j;
                    yk[j] = (col < y_end) ? __ldg(P.y_key + col) : INFINITY;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr + ch * 32, r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = fmaf(-2.f, __uint_as_float(r[j]), yk[ch * 32 + j]);   // +inf if masked
                    if (P.exclude_self) {
                        const int64_t selfcol = grow - P.self_offset - (c0 + ch * 32);
                        if (selfcol >= 0 && selfcol < 32) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j == (int)selfcol) v[j] = INFINITY;
                        }
                    }
                    select_chunk<32>(L, lrow, v, (int)(c0 + ch * 32), tau, cnt, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                if (++acc == Cfg::NUM_ACC) { acc = 0; acc_phase ^= 1; }
            }
            lists_flush(L, lrow, tau, cnt, lane);
            for (int r = 0; r < 32; ++r) {
                const int lr = warp * 32 + r;
                const int64_t gr = qt * BM + lr;
                if (gr >= P.nq) break;
                for (int p = lane; p < P.cap; p += 32) {
                    const int64_t o = gr * ((int64_t)P.splits * P.cap) + (int64_t)split * P.cap + p;
                    const ent_t e = L.ent[(size_t)lr * L.stride + p];
                    P.cand_idx[o] = entry_col(e);
                    if (P.cand_key) P.cand_key[o] = entry_key(e);
                }
            }
            __syncwarp();
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// rows x dpad fp32, row-major; box = BK features x box_rows rows, 128 B swizzle.
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int dpad, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    KB2_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dpad * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KB2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// Bytes the kernel carves out of dynamic shared memory, excluding the slack for
// rounding the base up to 1024 B (the base is 1024-aligned in practice: there is no
// static shared memory in this kernel; the kernel traps if the carve-up overflows).
static size_t tc_smem_bytes(int bn, int stages, int cap, int buf_slots) {
    const size_t stage = (size_t)(2 * BM + 2 * bn) * BK * 4;
    return stages * stage + 2 * bn * sizeof(float) + lists_bytes(BM, cap, buf_slots) +
           (2 * MAX_STAGES + 4) * 8 + 16;
}

template <int BN>
static int launch_tc_bn(const TcParams &P0, const float *q_hi, const float *q_lo, const float *y_hi,
                        const float *y_lo, int dpad, int stages, int sm_count, int max_smem,
                        cudaStream_t stream) {
    TcParams P = P0;
    P.stages = stages;
    P.per_split = ceil_div64(ceil_div64(P.ny, P.splits), BN) * BN;
    CUtensorMap mq_hi, mq_lo, my_hi, my_lo;
    if (make_map(&mq_hi, q_hi, P.nq, dpad, BM)) return 1;
    if (make_map(&mq_lo, q_lo, P.nq, dpad, BM)) return 1;
    if (make_map(&my_hi, y_hi, P.ny, dpad, BN)) return 1;
    if (make_map(&my_lo, y_lo, P.ny, dpad, BN)) return 1;
    const size_t need = tc_smem_bytes(BN, stages, P.cap, P.buf_slots);
    const size_t smem = min((size_t)max_smem, need + 1024);
    KB2_CUDA(cudaFuncSetAttribute(knn_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const int64_t units = P.q_tiles * P.splits;
    const unsigned grid = (unsigned)min((int64_t)sm_count, units);
    knn_tc_kernel<BN><<<grid, TC_THREADS, smem, stream>>>(mq_hi, mq_lo, my_hi, my_lo, P);
    KB2_LAUNCH_CHECK();
    return 0;
}

int launch_knn_tc(const float *q_hi, const float *q_lo, int64_t nq, const float *y_hi,
                  const float *y_lo, const float *y_key, int64_t ny, int dpad, int cap, int splits,
                  int exclude_self, int64_t self_offset, int32_t *cand_idx, float *cand_key,
                  cudaStream_t stream) {
    int dev = 0, sm_count = 0, max_smem = 0;
    KB2_CUDA(cudaGetDevice(&dev));
    KB2_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    KB2_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    TcParams P;
    P.nq = nq; P.ny = ny; P.kchunks = dpad / BK; P.cap = cap; P.splits = splits; P.stages = 0;
    P.exclude_self = exclude_self; P.self_offset = self_offset; P.per_split = 0; P.buf_slots = 16;
    P.q_tiles = ceil_div64(nq, BM); P.y_key = y_key; P.cand_idx = cand_idx; P.cand_key = cand_key;
    // Append-buffer slots per row: more slots = fewer merges; shrink towards 16 when the
    // lists would otherwise squeeze the operand pipeline below 2 stages.
    auto stages_with = [&](int bn, int slots) {
        const size_t fixed = tc_smem_bytes(bn, 0, cap, slots);
        const size_t stage = (size_t)(2 * BM + 2 * bn) * BK * 4;
        if (fixed >= (size_t)max_smem) return 0;
        return (int)min((size_t)MAX_STAGES, ((size_t)max_smem - fixed) / stage);
    };
    P.buf_slots = lists_buffer_slots(cap);
    while (P.buf_slots > 16 && stages_with(128, P.buf_slots) < 2) P.buf_slots -= 8;
    // widest tile whose pipeline still has >= 2 stages next to the candidate lists
    auto stages_for = [&](int bn) {
        const size_t fixed = tc_smem_bytes(bn, 0, cap, P.buf_slots);
        const size_t stage = (size_t)(2 * BM + 2 * bn) * BK * 4;
        if (fixed >= (size_t)max_smem) return 0;
        return (int)min((size_t)MAX_STAGES, ((size_t)max_smem - fixed) / stage);
    };
    const int s256 = stages_for(256), s128 = stages_for(128);
    if (s256 >= 2) return launch_tc_bn<256>(P, q_hi, q_lo, y_hi, y_lo, dpad, s256, sm_count, max_smem, stream);
    if (s128 >= 1) return launch_tc_bn<128>(P, q_hi, q_lo, y_hi, y_lo, dpad, s128, sm_count, max_smem, stream);
    set_error("knn_tc: candidate lists of %d entries do not fit in shared memory", cap);
    return 1;
}

}  // namespace kb2
