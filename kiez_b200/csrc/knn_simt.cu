// Candidate search on the FP32 FMA pipe: same inputs, outputs and selection code
// as the tcgen05 kernel (knn_tc.cu), but the dot products are plain FFMA tiles.
// It exists as an independent cross-check of the tensor-core path and as a
// debugging aid; kb2_knn_candidates(KB2_KNN_AUTO) never picks it.
#include "select.cuh"

namespace kb2 {

constexpr int SIMT_ROWS = 128;   // query rows per CTA = threads per CTA
constexpr int SIMT_COLS = 32;    // index rows per tile
constexpr int SIMT_K = 32;       // features per smem chunk

// grid = (q_tiles, splits); thread t owns query row q0 + t.
__global__ void __launch_bounds__(SIMT_ROWS)
knn_simt_kernel(const float *__restrict__ q_hi, const float *__restrict__ q_lo, int64_t nq,
                const float *__restrict__ y_hi, const float *__restrict__ y_lo,
                const float *__restrict__ y_key, int64_t ny, int dpad, int cap, int splits,
                int32_t *__restrict__ cand_idx, float *__restrict__ cand_key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *As = reinterpret_cast<float *>(smem_raw);            // [SIMT_K][SIMT_ROWS]
    float *Bs = As + SIMT_K * SIMT_ROWS;                        // [SIMT_K][SIMT_COLS]
    RowLists L;
    L.cap = cap;
    L.B = lists_buffer_slots(cap);
    L.stride = lists_stride(cap, L.B);
    L.ent = reinterpret_cast<ent_t *>(Bs + SIMT_K * SIMT_COLS);  // [SIMT_ROWS][stride]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * SIMT_ROWS;
    const int64_t row = q0 + tid;
    const int split = blockIdx.y;
    const int64_t per = ceil_div64(ceil_div64(ny, splits), SIMT_COLS) * SIMT_COLS;
    const int64_t y_begin = (int64_t)split * per;
    const int64_t y_end = min(ny, y_begin + per);

    lists_reset(L, warp * 32, 32, lane);
    float tau = (row < nq) ? INFINITY : -INFINITY;
    int cnt = 0;

    for (int64_t c0 = y_begin; c0 < y_end; c0 += SIMT_COLS) {
        float acc[SIMT_COLS];
#pragma unroll
        for (int j = 0; j < SIMT_COLS; ++j) acc[j] = 0.f;
        for (int k0 = 0; k0 < dpad; k0 += SIMT_K) {
            __syncthreads();
            // A chunk: thread t loads its own row (x = hi + lo restores the fp32 value)
            if (row < nq) {
                const float4 *ph = reinterpret_cast<const float4 *>(q_hi + row * dpad + k0);
                const float4 *pl = reinterpret_cast<const float4 *>(q_lo + row * dpad + k0);
#pragma unroll
                for (int v = 0; v < SIMT_K / 4; ++v) {
                    const float4 h = ph[v], l = pl[v];
                    As[(4 * v + 0) * SIMT_ROWS + tid] = h.x + l.x;
                    As[(4 * v + 1) * SIMT_ROWS + tid] = h.y + l.y;
                    As[(4 * v + 2) * SIMT_ROWS + tid] = h.z + l.z;
                    As[(4 * v + 3) * SIMT_ROWS + tid] = h.w + l.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < SIMT_K; ++v) As[v * SIMT_ROWS + tid] = 0.f;
            }
            // B chunk: 32 cols x 32 k; thread t loads 8 consecutive k of column t/4
            {
                const int col = tid >> 2, kk0 = (tid & 3) * 8;
                const int64_t yc = c0 + col;
                float vals[8];
                if (yc < y_end) {
                    const float4 *ph = reinterpret_cast<const float4 *>(y_hi + yc * dpad + k0 + kk0);
                    const float4 *pl = reinterpret_cast<const float4 *>(y_lo + yc * dpad + k0 + kk0);
                    const float4 h0 = ph[0], h1 = ph[1], l0 = pl[0], l1 = pl[1];
                    vals[0] = h0.x + l0.x; vals[1] = h0.y + l0.y; vals[2] = h0.z + l0.z; vals[3] = h0.w + l0.w;
                    vals[4] = h1.x + l1.x; vals[5] = h1.y + l1.y; vals[6] = h1.z + l1.z; vals[7] = h1.w + l1.w;
                } else {
#pragma unroll
                    for (int v = 0; v < 8; ++v) vals[v] = 0.f;
                }
#pragma unroll
                for (int v = 0; v < 8; ++v) Bs[(kk0 + v) * SIMT_COLS + col] = vals[v];
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < SIMT_K; ++kk) {
                const float a = As[kk * SIMT_ROWS + tid];
                const float4 *b4 = reinterpret_cast<const float4 *>(Bs + kk * SIMT_COLS);
#pragma unroll
                for (int v = 0; v < SIMT_COLS / 4; ++v) {
                    const float4 b = b4[v];
                    acc[4 * v + 0] = fmaf(a, b.x, acc[4 * v + 0]);
                    acc[4 * v + 1] = fmaf(a, b.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(a, b.z, acc[4 * v + 2]);
                    acc[4 * v + 3] = fmaf(a, b.w, acc[4 * v + 3]);
                }
            }
        }
        // keys + masking, then the shared selection
        float v[SIMT_COLS];
#pragma unroll
        for (int j = 0; j < SIMT_COLS; ++j) {
            const int64_t col = c0 + j;
            float key = INFINITY;
            if (col < y_end) key = fmaf(-2.f, acc[j], __ldg(y_key + col));
            v[j] = key;
        }
        select_chunk<SIMT_COLS>(L, tid, v, (int)c0, tau, cnt, lane);
    }
    lists_flush(L, tid, tau, cnt, lane);
    // write this split's lists: warp-cooperative, coalesced per row
    for (int r = 0; r < 32; ++r) {
        const int lr = warp * 32 + r;
        const int64_t grow = q0 + lr;
        if (grow >= nq) break;
        for (int p = lane; p < cap; p += 32) {
            const int64_t o = grow * ((int64_t)splits * cap) + (int64_t)split * cap + p;
            const ent_t e = L.ent[(size_t)lr * L.stride + p];
            cand_idx[o] = entry_col(e);
            if (cand_key) cand_key[o] = entry_key(e);
        }
    }
}

int launch_knn_simt(const float *q_hi, const float *q_lo, int64_t nq, const float *y_hi,
                    const float *y_lo, const float *y_key, int64_t ny, int dpad, int cap,
                    int splits, int32_t *cand_idx, float *cand_key, cudaStream_t stream) {
    const size_t smem = (size_t)(SIMT_K * SIMT_ROWS + SIMT_K * SIMT_COLS) * sizeof(float) +
                        lists_bytes(SIMT_ROWS, cap, lists_buffer_slots(cap));
    KB2_CUDA(cudaFuncSetAttribute(knn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    dim3 grid((unsigned)ceil_div64(nq, SIMT_ROWS), (unsigned)splits);
    knn_simt_kernel<<<grid, SIMT_ROWS, smem, stream>>>(q_hi, q_lo, nq, y_hi, y_lo, y_key, ny, dpad,
                                                       cap, splits, cand_idx, cand_key);
    KB2_LAUNCH_CHECK();
    return 0;
}

}  // namespace kb2
