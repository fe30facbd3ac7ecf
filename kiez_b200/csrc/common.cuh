// Shared helpers for the kiez_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/kiez_b200.h"

namespace kb2 {

void set_error(const char *fmt, ...);

#define KB2_CHECK(cond, ...)                                                            \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            kb2::set_error(__VA_ARGS__);                                                \
            return 1;                                                                   \
        }                                                                               \
    } while (0)

#define KB2_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            kb2::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                           __FILE__, __LINE__);                                         \
            return 2;                                                                   \
        }                                                                               \
    } while (0)

#define KB2_LAUNCH_CHECK() KB2_CUDA(cudaGetLastError())

constexpr unsigned FULL_MASK = 0xffffffffu;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Total order used by every sort in this library: ascending key, NaN last
// (numpy argsort/argpartition order), ties by the secondary integer.
__device__ __forceinline__ bool pair_less(double ka, int64_t ta, double kb, int64_t tb) {
    const bool na = isnan(ka), nb = isnan(kb);
    if (na || nb) return (!na && nb) || (na && nb && ta < tb);
    return ka < kb || (ka == kb && ta < tb);
}

// Bitonic sort of P (power of two) (key, tie, payload) triples held in shared
// memory, executed cooperatively by ONE warp.  tie doubles as the payload when
// payload == nullptr.
__device__ __forceinline__ void warp_bitonic_sort(double *key, int64_t *tie, int64_t *payload,
                                                  int P, int lane) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int t = lane; t < (P >> 1); t += 32) {
                const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` clear
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const double ka = key[lo], kb = key[hi];
                const int64_t ta = tie[lo], tb = tie[hi];
                const bool swap = up ? pair_less(kb, tb, ka, ta) : pair_less(ka, ta, kb, tb);
                if (swap) {
                    key[lo] = kb; key[hi] = ka;
                    tie[lo] = tb; tie[hi] = ta;
                    if (payload) { const int64_t pa = payload[lo]; payload[lo] = payload[hi]; payload[hi] = pa; }
                }
            }
        }
    }
    __syncwarp();
}

}  // namespace kb2
