// "Index build": 3xTF32 operand split + selection term + exact norms.
// Replaces SklearnNN._fit (kiez/neighbors/exact/sklearn_nearest_neighbors.py:83-94).
#include "common.cuh"

namespace kb2 {

__device__ __forceinline__ float to_tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// One warp per row.  HBM-bound: reads d*4 B, writes 2*dpad*4 + 12 B per row.
__global__ void __launch_bounds__(256)
prepare_rows_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ldx,
                    const float *__restrict__ center, int normalize, float *__restrict__ hi,
                    float *__restrict__ lo, int dpad, float *__restrict__ key_term,
                    double *__restrict__ sqnorm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *xr = x + row * ldx;
    double raw2 = 0.0;   // ||x||^2 of the raw row, fp64
    double cen2 = 0.0;   // ||x - center||^2, fp64
    for (int j = lane; j < d; j += 32) {
        const float v = xr[j];
        raw2 += (double)v * (double)v;
        const float w = center ? v - center[j] : v;
        cen2 += (double)w * (double)w;
    }
    raw2 = warp_sum(raw2);
    cen2 = warp_sum(cen2);
    float scale = 1.0f;
    if (normalize) scale = cen2 > 0.0 ? (float)(1.0 / sqrt(cen2)) : 1.0f;
    float *hr = hi + row * dpad;
    float *lr = lo + row * dpad;
    for (int j = lane; j < dpad; j += 32) {
        float h = 0.f, l = 0.f;
        if (j < d) {
            float w = xr[j];
            if (center) w -= center[j];
            w *= scale;
            h = to_tf32_rn(w);
            l = to_tf32_rn(w - h);
        }
        hr[j] = h;
        lr[j] = l;
    }
    if (lane == 0) {
        key_term[row] = normalize ? 0.0f : (float)cen2;
        if (sqnorm) sqnorm[row] = raw2;
    }
}

// max(0, max_i x[i]) of non-negative selection terms: the bit patterns of non-negative floats
// order like unsigned integers.
__global__ void __launch_bounds__(256)
max_f32_kernel(const float *__restrict__ x, int64_t n, unsigned int *__restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// ||lo||^2 per row, rounded UP, and its maximum: lo = rn_tf32(delta) with delta = w - hi the exact
// TF32 rounding error of the centred row w, so |delta_i| <= |lo_i| (1 + 2^-10) and
// ||delta||^2 <= ||lo||^2 (1 + 2^-9); the fp64 sum is rounded to fp32 with another 2^-20 of slack.
__global__ void __launch_bounds__(256)
split_error_kernel(const float *__restrict__ lo, int64_t n, int dpad, float *__restrict__ err_term) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float4 *lr = reinterpret_cast<const float4 *>(lo + row * dpad);
    double acc = 0.0;
    for (int j = lane; j < dpad / 4; j += 32) {
        const float4 v = lr[j];
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        err_term[row] = __double2float_ru(acc * (1.0 + 0x1p-9) * (1.0 + 0x1p-20));
    }
}

}  // namespace kb2

extern "C" int kb2_split_error_terms(const float *lo, int64_t n, int dpad, float *err_term,
                                     float *err_max, void *stream) {
    KB2_CHECK(n >= 0 && dpad > 0 && dpad % 32 == 0 && err_term && err_max,
              "split_error_terms: bad arguments");
    KB2_CUDA(cudaMemsetAsync(err_max, 0, sizeof(float), (cudaStream_t)stream));
    if (n == 0) return 0;
    const int warps = 8;
    kb2::split_error_kernel<<<(unsigned)kb2::ceil_div64(n, warps), warps * 32, 0, (cudaStream_t)stream>>>(
        lo, n, dpad, err_term);
    KB2_LAUNCH_CHECK();
    const int64_t blocks = kb2::ceil_div64(n, 256 * 8);
    kb2::max_f32_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, (cudaStream_t)stream>>>(
        err_term, n, reinterpret_cast<unsigned int *>(err_max));
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_max_f32(const float *x, int64_t n, float *out, void *stream) {
    KB2_CHECK(n >= 0 && out, "max_f32: bad arguments");
    KB2_CUDA(cudaMemsetAsync(out, 0, sizeof(float), (cudaStream_t)stream));
    if (n == 0) return 0;
    const int64_t blocks = kb2::ceil_div64(n, 256 * 8);
    kb2::max_f32_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, (cudaStream_t)stream>>>(
        x, n, reinterpret_cast<unsigned int *>(out));
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_padded_dim(int d) { return ((d + 31) / 32) * 32; }

extern "C" int kb2_prepare_rows(const float *x, int64_t n, int d, int64_t ldx,
                                const float *center, int metric, float *hi, float *lo,
                                int dpad, float *key_term, double *sqnorm, void *stream) {
    KB2_CHECK(n >= 0 && d > 0 && ldx >= d, "prepare_rows: bad shape n=%lld d=%d ldx=%lld",
              (long long)n, d, (long long)ldx);
    KB2_CHECK(dpad == kb2_padded_dim(d), "prepare_rows: dpad=%d, expected %d", dpad,
              kb2_padded_dim(d));
    KB2_CHECK(metric >= 0 && metric <= 2, "prepare_rows: unknown metric %d", metric);
    if (n == 0) return 0;
    const int warps = 8;
    const int64_t blocks = kb2::ceil_div64(n, warps);
    kb2::prepare_rows_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(
        x, n, d, ldx, center, metric == KB2_METRIC_COSINE, hi, lo, dpad, key_term, sqnorm);
    KB2_LAUNCH_CHECK();
    return 0;
}
