// Shared pieces of the two tensor-core candidate-search kernels (knn_tc.cu: one CTA per
// tile; knn_tc2.cu: CTA pairs with tcgen05 cta_group::2): PTX wrappers for mbarrier / TMA /
// tcgen05, descriptor builders, the per-tile epilogue (TMEM -> key finish -> selection)
// and the host-side tensor-map encoder.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "select.cuh"

namespace kb2 {

constexpr int BM = 128;          // query rows per CTA tile (UMMA M per CTA, TMEM lanes)
constexpr int UMMA_K = 8;        // tf32: 32 B of K per instruction
constexpr int TC_THREADS = 192;  // warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer
constexpr int MAX_STAGES = 8;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of CTA 0 of a pair

// A pipeline-stage row is BK*4 bytes = one swizzle span: BK=32 -> SWIZZLE_128B, BK=16 ->
// SWIZZLE_64B.  ROWS_B = index rows of the y tiles held by ONE CTA.
template <int ROWS_B, int BK>
struct StageCfg {
    static constexpr int ROW_BYTES = BK * 4;
    static constexpr int A_BYTES = BM * ROW_BYTES;              // q_hi or q_lo tile
    static constexpr int B_BYTES = ROWS_B * ROW_BYTES;          // y_hi or y_lo tile (this CTA's part)
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr uint32_t SBO = 8 * ROW_BYTES;              // 8-row core-matrix group pitch
    static constexpr uint32_t LAYOUT = (BK == 32) ? 2u : 4u;    // UMMA::LayoutType SWIZZLE_128B / _64B
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
// arrive on the barrier at a shared::cluster address (own CTA or the pair's CTA 0)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
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
// Bounded wait: a protocol bug traps (the launch fails loudly) instead of hanging the
// GPU.  try_wait suspends in hardware up to its time limit, so the loop turns slowly;
// the clock is consulted only every 4096 turns.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    for (uint32_t spin = 1;; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spin & 4095u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 8000000000LL) {
                printf("kiez_b200: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x,
                       threadIdx.x);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 2-D tile load into this CTA's shared memory, completion on this CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint32_t dst, uint32_t bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// Same, issued by either CTA of a pair: data lands in the issuing CTA's shared memory, the
// transaction bytes are credited to the mbarrier at `bar` (a shared::cluster address,
// here always CTA 0's, where the MMA issuer waits).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *map, uint32_t dst,
                                                 uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// CTA-pair form: M = 256 (128 rows from each CTA's A tile), N = 256 (128 rows from each
// CTA's B tile); each CTA's TMEM receives its own 128 rows x 256 columns.  Leader only.
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread retire
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(bar)
                 : "memory");
}
// pair form: arrives on the barrier at the same offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
        " [%0], %1;" ::"r"(bar),
        "h"(cta_mask)
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

// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 | LBO (unused for swizzled K-major, 1)<<16 | SBO (8 rows x row bytes)>>4 <<32 |
// version 1<<46 | layout_type<<61.  Only the low word depends on the address.
template <class Cfg>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    constexpr uint64_t hi = ((uint64_t)(Cfg::SBO >> 4) << 32) | ((uint64_t)1 << 46) |
                            ((uint64_t)Cfg::LAYOUT << 61) | ((uint64_t)1 << 16);
    return hi | (uint64_t)((saddr & 0x3FFFF) >> 4);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1)<<4,
// a_format/b_format TF32 (2)<<7/<<10, A and B K-major (0), N>>3 <<17, M>>4 <<24.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// Wave barrier of the persistent CTA-pair kernels: before a pair's producer starts the
// loads of its w-th work unit it waits until every pair that has a w-th unit has finished
// loading its (w-1)-th.  All pairs then sweep the index tiles of a wave in step, which is
// what lets them share each tile through L2 instead of each streaming it from HBM (pairs
// that drift apart by more than the L2 capacity, ~4 % of a 1M-row sweep, stop sharing).
// Safe to spin: the grid never exceeds one CTA per SM, so all pairs are co-resident.
constexpr int MAX_WAVES = 8192;
__device__ unsigned int g_wave_arrivals[MAX_WAVES];

__device__ __forceinline__ void wave_barrier(int wave, unsigned int expected) {
    // called by ONE thread per CTA pair
    __threadfence();
    atomicAdd(&g_wave_arrivals[wave], 1u);
    const long long t0 = clock64();
    while (atomicAdd(&g_wave_arrivals[wave], 0u) < expected) {
        __nanosleep(200);
        if (clock64() - t0 > 8000000000LL) {
            printf("kiez_b200: wave barrier timed out (block %d wave %d)\n", blockIdx.x, wave);
            __trap();
        }
    }
}

struct TcParams {
    int wave_sync;        // 1: producers meet at every work-unit boundary (wave_barrier)
    int64_t nq, ny;
    int kchunks;          // dpad / BK
    int cap, buf_slots, splits, stages;
    int64_t per_split;    // index rows per split (multiple of the index tile)
    int64_t q_tiles;      // 128-row query tiles
    const float *y_key;
    int32_t *cand_idx;
    float *cand_key;
};

// ---------------------------------------------------------------- epilogue pieces
// Selection terms y_key[c0 .. c0+BN) -> registers, lane l holds columns t*32 + l.
template <int BN>
__device__ __forceinline__ void load_ykey(const float *__restrict__ y_key, int64_t c0,
                                          int64_t y_end, int lane, float (&ykreg)[BN / 32]) {
#pragma unroll
    for (int t = 0; t < BN / 32; ++t) {
        const int64_t col = c0 + t * 32 + lane;
        ykreg[t] = (col < y_end) ? __ldg(y_key + col) : INFINITY;   // +inf masks the column
    }
}

// Key finish of four accumulator elements against four per-column terms read from shared memory
// (a warp-uniform address: one broadcast LDS.128):  v[i] = term[i] - 2 acc[i]  as two packed
// FFMA2 (fma.rn.f32x2: same rounding as fmaf, half the issue slots of the epilogue's hottest
// loop).  The explicit shared-space load matters too: through the generic `const float *` the
// compiler emitted generic LD.E.128 for these tiles.
__device__ __forceinline__ void key_finish4(uint32_t term_saddr, uint32_t a0, uint32_t a1, uint32_t a2,
                                            uint32_t a3, float &v0, float &v1, float &v2, float &v3) {
    float t0, t1, t2, t3;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t0), "=f"(t1), "=f"(t2), "=f"(t3)
                 : "r"(term_saddr)
                 : "memory");
    asm("{\n\t.reg .b64 a, t, m, r;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 t, {%4, %5};\n\t"
        "mov.b64 m, {%6, %6};\n\t"
        "fma.rn.f32x2 r, a, m, t;\n\t"
        "mov.b64 {%0, %1}, r;\n\t}"
        : "=f"(v0), "=f"(v1)
        : "r"(a0), "r"(a1), "f"(t0), "f"(t1), "f"(-2.f));
    asm("{\n\t.reg .b64 a, t, m, r;\n\t"
        "mov.b64 a, {%2, %3};\n\t"
        "mov.b64 t, {%4, %5};\n\t"
        "mov.b64 m, {%6, %6};\n\t"
        "fma.rn.f32x2 r, a, m, t;\n\t"
        "mov.b64 {%0, %1}, r;\n\t}"
        : "=f"(v2), "=f"(v3)
        : "r"(a2), "r"(a3), "f"(t2), "f"(t3), "f"(-2.f));
}

// One accumulator tile (this warp's 32 TMEM lanes x BN columns): key finish + selection.
// `yk` = this warp's private shared copy of the tile's selection terms.
template <int BN, bool DOUBLE_BUFFER = true>
__device__ __forceinline__ void epilogue_tile(const RowLists &L, int lrow, const float *yk,
                                              uint32_t taddr, int64_t c0, float &tau, int &cnt,
                                              int lane) {
    constexpr int NCH = BN / 32;
    uint32_t ra[32];
    const uint32_t yk_saddr = smem_u32(yk);
    auto process = [&](const uint32_t (&r)[32], int ch) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            key_finish4(yk_saddr + (uint32_t)(ch * 32 + j) * 4u, r[j], r[j + 1], r[j + 2], r[j + 3],
                        v[j], v[j + 1], v[j + 2], v[j + 3]);
        select_chunk<32>(L, lrow, v, (int)(c0 + ch * 32), tau, cnt, lane);
    };
    if constexpr (!DOUBLE_BUFFER) {
        // two epilogue warps per scheduler hide the TMEM latency; saves 32 registers
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
            tmem_ld_32x32b_x32(taddr + ch * 32, ra);
            tmem_ld_wait();
            process(ra, ch);
        }
        return;
    }
    uint32_t rb[32];
    tmem_ld_32x32b_x32(taddr, ra);
#pragma unroll 1
    for (int ch = 0; ch < NCH; ch += 2) {
        tmem_ld_wait();
        tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, rb);   // in flight while ra is processed
        process(ra, ch);
        tmem_ld_wait();
        if (ch + 2 < NCH) tmem_ld_32x32b_x32(taddr + (ch + 2) * 32, ra);
        process(rb, ch + 1);
    }
}

// Write the 32 finished lists of this warp: [row][split*cap + p], coalesced per row.
__device__ __forceinline__ void write_lists(const RowLists &L, const TcParams &P, int warp, int lane,
                                            int64_t row0, int split) {
    for (int r = 0; r < 32; ++r) {
        const int lr = warp * 32 + r;
        const int64_t gr = row0 + lr;
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

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
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

// rows x dpad fp32, row-major; box = bk features x box_rows rows, swizzle span = bk*4 bytes.
static inline int make_map(CUtensorMap *map, const float *base, int64_t rows, int dpad,
                           int box_rows, int bk) {
    EncodeTiledFn enc = get_encode_fn();
    KB2_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dpad * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KB2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// Arms the wave barrier for one launch (stream-ordered reset of the arrival counters).
// The counters are one per-device array: launches of the pair kernels on different streams
// of the same device must not overlap (the host classes use one stream).
// Measured at C4 (profiles/r01_ab_experiments.md block G): the one-direction pair kernel gains
// (DRAM reads halve, clocks under the power cap rise), the dual-direction kernel loses (its
// column epilogue makes unit times uneven) -> `default_on` differs; KB2_WAVE_SYNC=0/1 overrides.
static inline int prepare_wave_sync(TcParams &P, int64_t units, unsigned pairs, bool default_on,
                                    cudaStream_t stream) {
    const int64_t waves = (units + pairs - 1) / pairs;
    const char *env = getenv("KB2_WAVE_SYNC");
    const bool on = env ? env[0] != '0' : default_on;
    P.wave_sync = (waves > 1 && waves <= MAX_WAVES && on) ? 1 : 0;
    if (P.wave_sync) {
        void *addr = nullptr;
        KB2_CUDA(cudaGetSymbolAddress(&addr, g_wave_arrivals));
        KB2_CUDA(cudaMemsetAsync(addr, 0, (size_t)waves * sizeof(unsigned int), stream));
    }
    return 0;
}

// Shared-memory bytes next to the operand stages: per-epilogue-warp y_key tile, the
// candidate lists, the barriers.
static inline size_t tc_fixed_smem(int bn, int cap, int buf_slots, int ykey_copies = 4) {
    return (size_t)ykey_copies * bn * sizeof(float) + lists_bytes(BM, cap, buf_slots) +
           (2 * MAX_STAGES + 4) * 8 + 16;
}

}  // namespace kb2
