// kiez.analysis.hubness_score on device (kiez/analysis/estimation.py:272-351) and
// kiez.evaluate.hits (kiez/evaluate/eval_metrics.py:23-61).  Integer work is exact
// (int64); scalar measures are fp64 block reductions.  HBM-bound: one pass over the
// (n, k) ids for the histogram, a handful of passes over the histogram.
#include "common.cuh"

namespace kb2 {

__global__ void index_range_kernel(const int64_t *__restrict__ ind, int64_t n, int64_t ld, int k,
                                   int64_t *__restrict__ out) {
    int64_t mn = INT64_MAX, mx = INT64_MIN;
    const int64_t total = n * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = ind[(i / k) * ld + (i % k)];
        mn = min(mn, v);
        mx = max(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(FULL_MASK, mn, o));
        mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(reinterpret_cast<long long *>(out), (long long)mn);
        atomicMax(reinterpret_cast<long long *>(out + 1), (long long)mx);
    }
}
__global__ void index_range_init_kernel(int64_t *out) {
    out[0] = INT64_MAX;
    out[1] = INT64_MIN;
}

// bincount of the first k columns; negatives dropped (estimation.py:286-295)
__global__ void k_occurrence_kernel(const int64_t *__restrict__ ind, int64_t n, int64_t ld, int k,
                                    int64_t nbins, unsigned long long *__restrict__ hist) {
    const int64_t total = n * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = ind[(i / k) * ld + (i % k)];
        if (v >= 0 && v < nbins) atomicAdd(hist + v, 1ULL);
    }
}

constexpr int MOM_THREADS = 256;
constexpr int N_MOM = 10;

__global__ void __launch_bounds__(MOM_THREADS)
hub_moments_kernel(const int64_t *__restrict__ hist, int64_t nbins, double mean, double hub_thresh,
                   double *__restrict__ out) {
    double acc[N_MOM];
#pragma unroll
    for (int i = 0; i < N_MOM; ++i) acc[i] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)hist[i];
        const double e = x - mean;
        acc[0] += x;
        acc[1] += e * e;
        acc[2] += e * e * e;
        acc[3] += fabs(e);
        acc[4] += sqrt(x);
        acc[5] = fmax(acc[5], x);
        acc[6] += (x == 0.0);
        const bool hub = x >= hub_thresh;
        acc[7] += hub;
        acc[8] += hub ? x : 0.0;
        acc[9] += x * x;
    }
    __shared__ double red[N_MOM][MOM_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N_MOM; ++i) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(FULL_MASK, v, o);
            v = (i == 5) ? fmax(v, w) : v + w;
        }
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < N_MOM) {
        const int i = threadIdx.x;
        double v = red[i][0];
        for (int w = 1; w < MOM_THREADS / 32; ++w) v = (i == 5) ? fmax(v, red[i][w]) : v + red[i][w];
        if (i == 5) {
            // max via CAS on the fp64 bit pattern (all values >= 0)
            unsigned long long *a = reinterpret_cast<unsigned long long *>(out + 5);
            atomicMax(a, (unsigned long long)__double_as_longlong(v));
        } else {
            atomicAdd(out + i, v);
        }
    }
}

// ascending stream compaction of matching bins: per-1024-block counts, a serial
// scan of the (few) block counts, then an ordered scatter.
__device__ __forceinline__ bool bin_match(int64_t x, int mode, double thresh) {
    return mode == 0 ? (x == 0) : ((double)x >= thresh);
}
__global__ void __launch_bounds__(1024)
compact_count_kernel(const int64_t *__restrict__ hist, int64_t nbins, int mode, double thresh,
                     int64_t *__restrict__ block_counts) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int hit = (i < nbins) && bin_match(hist[i], mode, thresh);
    const int total = __syncthreads_count(hit);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}
__global__ void compact_scan_kernel(int64_t *block_counts, int64_t nblocks, int64_t *out_count) {
    // single thread: nblocks = nbins/1024 (<= ~10^4 for 10M bins)
    int64_t run = 0;
    for (int64_t b = 0; b < nblocks; ++b) {
        const int64_t c = block_counts[b];
        block_counts[b] = run;
        run += c;
    }
    block_counts[nblocks] = run;
    *out_count = run;
}
__global__ void __launch_bounds__(1024)
compact_scatter_kernel(const int64_t *__restrict__ hist, int64_t nbins, int mode, double thresh,
                       const int64_t *__restrict__ block_offsets, int64_t *__restrict__ out_ids) {
    __shared__ int warp_counts[32];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hit = (i < nbins) && bin_match(hist[i], mode, thresh);
    const unsigned bal = __ballot_sync(FULL_MASK, hit);
    if (lane == 0) warp_counts[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_counts[w];
    if (hit) out_ids[block_offsets[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1))] = i;
}

// Gini numerator sum_ij |x_i - x_j| = 2 * sum_v v * H[v] * (2 P[v] + H[v] - n), with H the
// histogram of occurrence values and P its exclusive prefix (ranks of value v in sorted order).
__global__ void value_hist_kernel(const int64_t *__restrict__ hist, int64_t nbins, int64_t max_value,
                                  unsigned long long *__restrict__ vh) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = hist[i];
        if (v >= 0 && v <= max_value) atomicAdd(vh + v, 1ULL);
    }
}
__global__ void __launch_bounds__(1024)
gini_from_value_hist_kernel(const int64_t *__restrict__ vh, int64_t nvals, int64_t nbins,
                            int64_t *__restrict__ out) {
    // single block: chunked inclusive scan over the value histogram
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    __shared__ long long acc_s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    long long local = 0;
    __syncthreads();
    for (int64_t base = 0; base < nvals; base += 1024) {
        const int64_t v = base + threadIdx.x;
        const long long h = (v < nvals) ? (long long)vh[v] : 0;
        long long incl = h;
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        long long before = carry_s;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        const long long excl = before + incl - h;            // P[v]
        local += (long long)v * h * (2 * excl + h - (long long)nbins);
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + incl;
        __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL_MASK, local, o);
    if (lane == 0) acc_s[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 32; ++w) t += acc_s[w];
        *out = 2 * t;
    }
}

__global__ void hits_kernel(const int64_t *__restrict__ ind, int64_t n, int64_t ld, int k,
                            const int64_t *__restrict__ gold, const int32_t *__restrict__ ks, int nks,
                            unsigned long long *__restrict__ counts) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = gold[r];
        if (g < 0) continue;
        int pos = k;
        for (int j = 0; j < k; ++j)
            if (ind[r * ld + j] == g) { pos = j; break; }
        for (int i = 0; i < nks; ++i)
            if (pos < ks[i]) atomicAdd(counts + i, 1ULL);
    }
}

static inline unsigned grid_for(int64_t work, int threads) {
    const int64_t b = ceil_div64(work, threads);
    return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace kb2

using namespace kb2;

extern "C" int kb2_index_range(const int64_t *ind, int64_t n, int64_t ld, int k, int64_t *out,
                               void *stream) {
    KB2_CHECK(n >= 0 && k > 0 && ld >= k, "index_range: bad shape");
    index_range_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out);
    if (n > 0)
        index_range_kernel<<<grid_for(n * k, 256), 256, 0, (cudaStream_t)stream>>>(ind, n, ld, k, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_k_occurrence(const int64_t *ind, int64_t n, int64_t ld, int k, int64_t nbins,
                                int64_t *hist, void *stream) {
    KB2_CHECK(n >= 0 && k > 0 && ld >= k && nbins > 0, "k_occurrence: bad shape");
    KB2_CUDA(cudaMemsetAsync(hist, 0, (size_t)nbins * sizeof(int64_t), (cudaStream_t)stream));
    if (n > 0)
        k_occurrence_kernel<<<grid_for(n * k, 256), 256, 0, (cudaStream_t)stream>>>(
            ind, n, ld, k, nbins, reinterpret_cast<unsigned long long *>(hist));
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_hub_moments(const int64_t *hist, int64_t nbins, double mean, double hub_thresh,
                               double *out, void *stream) {
    KB2_CHECK(nbins > 0, "hub_moments: bad shape");
    KB2_CUDA(cudaMemsetAsync(out, 0, N_MOM * sizeof(double), (cudaStream_t)stream));
    hub_moments_kernel<<<grid_for(nbins, MOM_THREADS), MOM_THREADS, 0, (cudaStream_t)stream>>>(
        hist, nbins, mean, hub_thresh, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_compact_ids(const int64_t *hist, int64_t nbins, int mode, double thresh,
                               int64_t *scratch, int64_t *out_ids, int64_t *out_count, void *stream) {
    KB2_CHECK(nbins > 0 && (mode == 0 || mode == 1), "compact_ids: bad arguments");
    const int64_t nblocks = ceil_div64(nbins, 1024);
    cudaStream_t st = (cudaStream_t)stream;
    compact_count_kernel<<<(unsigned)nblocks, 1024, 0, st>>>(hist, nbins, mode, thresh, scratch);
    compact_scan_kernel<<<1, 1, 0, st>>>(scratch, nblocks, out_count);
    compact_scatter_kernel<<<(unsigned)nblocks, 1024, 0, st>>>(hist, nbins, mode, thresh, scratch,
                                                             out_ids);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_gini_numerator(const int64_t *hist, int64_t nbins, int64_t max_value,
                                  int64_t *scratch, int64_t *out, void *stream) {
    KB2_CHECK(nbins > 0 && max_value >= 0, "gini_numerator: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nvals = max_value + 1;
    KB2_CUDA(cudaMemsetAsync(scratch, 0, (size_t)nvals * sizeof(int64_t), st));
    value_hist_kernel<<<grid_for(nbins, 256), 256, 0, st>>>(
        hist, nbins, max_value, reinterpret_cast<unsigned long long *>(scratch));
    gini_from_value_hist_kernel<<<1, 1024, 0, st>>>(scratch, nvals, nbins, out);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_hits(const int64_t *ind, int64_t n, int64_t ld, int k, const int64_t *gold,
                        const int32_t *ks, int nks, int64_t *counts, void *stream) {
    KB2_CHECK(n >= 0 && k > 0 && ld >= k && nks > 0, "hits: bad shape");
    if (n == 0) return 0;
    hits_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        ind, n, ld, k, gold, ks, nks, reinterpret_cast<unsigned long long *>(counts));
    KB2_LAUNCH_CHECK();
    return 0;
}
