// Exact float64 finish of the candidate lists, and the generic row-wise top-k
// (final hubness sort + multi-GPU merge).
#include "common.cuh"

namespace kb2 {

constexpr int REFINE_WARPS = 4;

// Completeness proof for candidate lists that come from the 1xTF32 screen (knn_screen.cu).
// Every index row NOT in the list has screen key >= tau (the list's cap-th best screen key;
// the minimum over the lists when the index was searched in independent ranges), and
// |screen key - exact key| <= E, where
//   exact key  = d^2(q, y) - ||q-c||^2 = ||y-c||^2 - 2 <q-c, y-c>   (euclidean metrics, c = centring vector)
//              = 2 (cosine distance - 1)                             (cosine: rows are normalised)
//   screen key = fp32(y_key - 2 acc),  acc = tcgen05 fp32 accumulation of <hi(q), hi(y)>.
// With w = the fp32 centred (normalised) row, hi = rn_tf32(w), delta = w - hi (known exactly at
// prepare time: kb2_split_error_terms gives ||delta||^2 per row, rounded up, and its maximum):
//   <w_q, w_y> - <hi_q, hi_y> = <delta_q, w_y> + <hi_q, delta_y>
//   |.| <= ||delta_q|| ||w_y||_max + ||w_q|| (1 + 2^-11) ||delta_y||_max            (Cauchy-Schwarz per term)
// plus eps_acc ||w_q|| ||w_y||_max for the accumulation (TF32 products are exact in fp32; dpad
// additions of unknown order, truncation allowed: eps_acc = dpad 2^-22 + 2^-21, which also covers
// the fp32 rounding of the centring), so
//   E = 2 (dq ym + qn (1 + 2^-11) dym + eps_acc qn ym) + 2^-21 (ym^2 + 2 qn ym)     (last: fp32 key terms)
// -- about 2.4x tighter than the operand-independent 2^-10 ||q|| ||y|| bound, since the RMS TF32
// rounding error of a row is ~0.2 x 2^-10 of its norm.
// So if the exact k-th best distance satisfies  key_k < tau - E  no outside row can belong to
// the k nearest and the result is exact; otherwise unverified[row] = 1 and the caller
// searches that row again with the 3xTF32 kernel.
struct CheckParams {
    const float *tau;          // first tau of row 0
    int64_t tau_row_stride;    // elements between rows
    int tau_step, tau_count;   // tau_count values per row, tau_step elements apart
    const float *q_key;        // [nq] fp32 ||w_q||^2 (unused for cosine: 1)
    const float *y_key_max;    // device scalar: max ||w_y||^2 over the index (unused for cosine)
    const float *q_err;        // [nq] fp32 ||delta_q||^2, rounded up
    const float *y_err_max;    // device scalar: max ||delta_y||^2 over the index
    double eps_acc;            // accumulation bound relative to ||w_q|| ||w_y||
    int32_t *unverified;       // [nq] out
};

// One warp per query row: gather each candidate's raw fp32 row, accumulate the
// distance in fp64, then bitonic-sort (dist, id) in shared memory and write the
// best k.  HBM/L2-bound on the row gathers (ncand * d * 4 B per query).
template <typename T, bool VEC4, bool CHECK>
__global__ void __launch_bounds__(REFINE_WARPS * 32)
refine_topk_kernel(const T *__restrict__ q, int64_t nq, int64_t ldq,
                   const T *__restrict__ y, int64_t ny, int64_t ldy, int d,
                   const double *__restrict__ q_sqnorm, const double *__restrict__ y_sqnorm,
                   const int32_t *__restrict__ cand_idx, int ncand, int P, int metric,
                   int64_t index_base, int exclude_self, int64_t self_offset, int k,
                   double *__restrict__ out_dist, int64_t *__restrict__ out_ind,
                   const CheckParams CP) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *key = reinterpret_cast<double *>(smem_raw) + (size_t)warp * P;
    int64_t *tie = reinterpret_cast<int64_t *>(reinterpret_cast<double *>(smem_raw) +
                                               (size_t)REFINE_WARPS * P) + (size_t)warp * P;
    const int64_t row = (int64_t)blockIdx.x * REFINE_WARPS + warp;
    if (row >= nq) return;
    const T *qr = q + row * ldq;
    const int32_t *cr = cand_idx + row * (int64_t)ncand;
    for (int j = 0; j < P; ++j) {
        int32_t id = (j < ncand) ? cr[j] : -1;
        if (exclude_self && (int64_t)id + self_offset == row) id = -1;   // the query's own row
        double dist = INFINITY;
        if (id >= 0 && id < ny) {
            const T *yr = y + (int64_t)id * ldy;
            double acc = 0.0;
            if (metric == KB2_METRIC_COSINE) {
                if constexpr (VEC4) {
                    for (int t = lane * 4; t < d; t += 128) {
                        const float4 a = *reinterpret_cast<const float4 *>(qr + t);
                        const float4 b = *reinterpret_cast<const float4 *>(yr + t);
                        acc += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z +
                               (double)a.w * b.w;
                    }
                } else {
                    for (int t = lane; t < d; t += 32) acc += (double)qr[t] * (double)yr[t];
                }
                acc = warp_sum(acc);
                double nq2 = q_sqnorm[row], ny2 = y_sqnorm[id];
                const double nx = nq2 > 0.0 ? sqrt(nq2) : 1.0;
                const double nyv = ny2 > 0.0 ? sqrt(ny2) : 1.0;
                dist = 1.0 - acc / (nx * nyv);
                dist = fmin(fmax(dist, 0.0), 2.0);
            } else {
                if constexpr (VEC4) {
                    for (int t = lane * 4; t < d; t += 128) {
                        const float4 a = *reinterpret_cast<const float4 *>(qr + t);
                        const float4 b = *reinterpret_cast<const float4 *>(yr + t);
                        const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y,
                                     dz = (double)a.z - (double)b.z, dw = (double)a.w - (double)b.w;
                        acc += dx * dx + dy * dy + dz * dz + dw * dw;
                    }
                } else {
                    for (int t = lane; t < d; t += 32) {
                        const double dx = (double)qr[t] - (double)yr[t];
                        acc += dx * dx;
                    }
                }
                acc = warp_sum(acc);
                dist = (metric == KB2_METRIC_EUCLIDEAN) ? sqrt(acc) : acc;
            }
        }
        if (lane == 0) {
            key[j] = dist;
            tie[j] = (id >= 0 && id < ny) ? (int64_t)id + index_base : INT64_MAX;
        }
    }
    warp_bitonic_sort(key, tie, nullptr, P, lane);
    for (int j = lane; j < k; j += 32) {
        out_dist[row * k + j] = key[j];
        out_ind[row * k + j] = (tie[j] == INT64_MAX) ? -1 : tie[j];
    }
    if constexpr (CHECK) {
        if (lane == 0) {
            float tau = INFINITY;
            for (int j = 0; j < CP.tau_count; ++j)
                tau = fminf(tau, CP.tau[row * CP.tau_row_stride + (int64_t)j * CP.tau_step]);
            const double kd = key[k - 1];          // exact k-th best distance (+inf: fewer than k)
            bool ok;
            if (tau == INFINITY) {
                ok = true;                         // lists not full: every index row is a candidate
            } else if (!(kd < INFINITY)) {
                ok = false;
            } else {
                const bool cosine = metric == KB2_METRIC_COSINE;
                const double up = 1.000001;            // slack of the fp32 norms and square roots
                const double qn2 = cosine ? 1.0 : (double)CP.q_key[row];
                const double ym2 = cosine ? 1.0 : (double)*CP.y_key_max;
                const double qn = sqrt(qn2) * up, ym = sqrt(ym2) * up;
                const double dq = sqrt((double)CP.q_err[row]) * up;
                const double dym = sqrt((double)*CP.y_err_max) * up;
                const double E = 2.0 * (dq * ym + qn * (1.0 + 0x1p-11) * dym + CP.eps_acc * qn * ym) +
                                 4.76837158203125e-07 * (ym2 + 2.0 * qn * ym);
                if (cosine) {
                    ok = 2.0 * (kd - 1.0) + 1e-9 < (double)tau - E - 1e-6;
                } else {
                    const double d2 = (metric == KB2_METRIC_EUCLIDEAN) ? kd * kd : kd;
                    ok = d2 * (1.0 + 1e-12) < (double)tau - E + qn2 * (1.0 - 2.4e-7);
                }
            }
            CP.unverified[row] = ok ? 0 : 1;
        }
    }
}

// Row-wise top-k of (dist, ind): one warp per row, bitonic sort in smem keyed by
// (dist, input position) so that ties resolve like a stable argsort.
__global__ void __launch_bounds__(REFINE_WARPS * 32)
topk_rows_kernel(const double *__restrict__ dist, const int64_t *__restrict__ ind, int64_t n, int c,
                 int nparts, int64_t part_stride, int P, int k, double *__restrict__ out_dist,
                 int64_t *__restrict__ out_ind) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *key = reinterpret_cast<double *>(smem_raw) + (size_t)warp * P;
    int64_t *tie = reinterpret_cast<int64_t *>(reinterpret_cast<double *>(smem_raw) +
                                               (size_t)REFINE_WARPS * P) + (size_t)warp * P;
    int64_t *pay = reinterpret_cast<int64_t *>(reinterpret_cast<double *>(smem_raw) +
                                               (size_t)2 * REFINE_WARPS * P) + (size_t)warp * P;
    const int64_t row = (int64_t)blockIdx.x * REFINE_WARPS + warp;
    if (row >= n) return;
    const int total = c * nparts;
    for (int j = lane; j < P; j += 32) {
        if (j < total) {
            const int part = j / c, w = j - part * c;
            const int64_t o = (int64_t)part * part_stride + row * c + w;
            key[j] = dist[o];
            pay[j] = ind[o];
            tie[j] = j;
        } else {
            key[j] = __longlong_as_double(0x7ff8000000000000LL);   // NaN pad sorts last
            pay[j] = -1;
            tie[j] = INT64_MAX;
        }
    }
    warp_bitonic_sort(key, tie, pay, P, lane);
    for (int j = lane; j < k; j += 32) {
        out_dist[row * k + j] = key[j];
        out_ind[row * k + j] = pay[j];
    }
}

int launch_topk_rows_rg(const double *, const int64_t *, int64_t, int, int, int64_t, int, double *,
                        int64_t *, cudaStream_t);   // rescale.cu, register-resident path

}  // namespace kb2

template <typename T, bool VEC4, bool CHECK>
static int launch_refine(const void *q, int64_t nq, int64_t ldq, const void *y, int64_t ny,
                         int64_t ldy, int d, const double *q_sqnorm, const double *y_sqnorm,
                         const int32_t *cand_idx, int ncand, int P, int metric, int64_t index_base,
                         int exclude_self, int64_t self_offset, int k, double *out_dist,
                         int64_t *out_ind, const kb2::CheckParams &CP, cudaStream_t st) {
    using namespace kb2;
    const size_t smem = (size_t)REFINE_WARPS * P * 16;
    KB2_CUDA(cudaFuncSetAttribute(refine_topk_kernel<T, VEC4, CHECK>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    refine_topk_kernel<T, VEC4, CHECK><<<(unsigned)ceil_div64(nq, REFINE_WARPS), REFINE_WARPS * 32, smem, st>>>(
        static_cast<const T *>(q), nq, ldq, static_cast<const T *>(y), ny, ldy, d, q_sqnorm, y_sqnorm,
        cand_idx, ncand, P, metric, index_base, exclude_self, self_offset, k, out_dist, out_ind, CP);
    KB2_LAUNCH_CHECK();
    return 0;
}

template <bool CHECK>
static int refine_dispatch(const void *q, int64_t nq, int64_t ldq, const void *y, int64_t ny,
                           int64_t ldy, int d, int elem_size, const double *q_sqnorm,
                           const double *y_sqnorm, const int32_t *cand_idx, int ncand, int metric,
                           int64_t index_base, int exclude_self, int64_t self_offset, int k,
                           double *out_dist, int64_t *out_ind, const kb2::CheckParams &CP,
                           cudaStream_t st) {
    using namespace kb2;
    KB2_CHECK(nq >= 0 && ny > 0 && d > 0 && ldq >= d && ldy >= d, "refine_topk: bad shape");
    KB2_CHECK(elem_size == 4 || elem_size == 8, "refine_topk: elem_size must be 4 (fp32) or 8 (fp64)");
    KB2_CHECK(ncand > 0 && ncand <= 2048, "refine_topk: ncand=%d outside (0, 2048]", ncand);
    KB2_CHECK(k > 0 && k <= ncand, "refine_topk: k=%d must be in (0, ncand=%d]", k, ncand);
    KB2_CHECK(metric >= 0 && metric <= 2, "refine_topk: unknown metric %d", metric);
    KB2_CHECK(metric != KB2_METRIC_COSINE || (q_sqnorm && y_sqnorm),
              "refine_topk: cosine needs q_sqnorm and y_sqnorm");
    if (nq == 0) return 0;
    const int P = next_pow2(ncand);
    if (elem_size == 8)
        return launch_refine<double, false, CHECK>(q, nq, ldq, y, ny, ldy, d, q_sqnorm, y_sqnorm, cand_idx,
                                                   ncand, P, metric, index_base, exclude_self, self_offset, k, out_dist, out_ind, CP, st);
    const bool vec = (d % 4 == 0) && (ldq % 4 == 0) && (ldy % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(y)) % 16 == 0);
    if (vec)
        return launch_refine<float, true, CHECK>(q, nq, ldq, y, ny, ldy, d, q_sqnorm, y_sqnorm, cand_idx,
                                                 ncand, P, metric, index_base, exclude_self, self_offset, k, out_dist, out_ind, CP, st);
    return launch_refine<float, false, CHECK>(q, nq, ldq, y, ny, ldy, d, q_sqnorm, y_sqnorm, cand_idx, ncand,
                                              P, metric, index_base, exclude_self, self_offset, k, out_dist,
                                              out_ind, CP, st);
}

extern "C" int kb2_refine_topk(const void *q, int64_t nq, int64_t ldq, const void *y, int64_t ny,
                               int64_t ldy, int d, int elem_size, const double *q_sqnorm,
                               const double *y_sqnorm, const int32_t *cand_idx, int ncand,
                               int metric, int64_t index_base, int exclude_self,
                               int64_t self_offset, int k, double *out_dist, int64_t *out_ind,
                               void *stream) {
    kb2::CheckParams CP = {};
    return refine_dispatch<false>(q, nq, ldq, y, ny, ldy, d, elem_size, q_sqnorm, y_sqnorm, cand_idx,
                                  ncand, metric, index_base, exclude_self, self_offset, k, out_dist,
                                  out_ind, CP, (cudaStream_t)stream);
}

extern "C" int kb2_refine_topk_checked(const void *q, int64_t nq, int64_t ldq, const void *y,
                                       int64_t ny, int64_t ldy, int d, int elem_size,
                                       const double *q_sqnorm, const double *y_sqnorm,
                                       const int32_t *cand_idx, int ncand, int metric,
                                       int64_t index_base, int exclude_self, int64_t self_offset,
                                       int k, double *out_dist, int64_t *out_ind, const float *tau,
                                       int64_t tau_row_stride, int tau_step, int tau_count,
                                       const float *q_key, const float *y_key_max,
                                       const float *q_err, const float *y_err_max, double eps_acc,
                                       int32_t *unverified, void *stream) {
    using namespace kb2;
    KB2_CHECK(tau && unverified && tau_count >= 1, "refine_topk_checked: tau and unverified are required");
    KB2_CHECK(metric == KB2_METRIC_COSINE || (q_key && y_key_max),
              "refine_topk_checked: euclidean metrics need q_key and y_key_max");
    KB2_CHECK(q_err && y_err_max, "refine_topk_checked: q_err and y_err_max (kb2_split_error_terms) are required");
    KB2_CHECK(eps_acc > 0.0 && eps_acc < 1.0, "refine_topk_checked: eps_acc=%g out of range", eps_acc);
    CheckParams CP;
    CP.tau = tau; CP.tau_row_stride = tau_row_stride; CP.tau_step = tau_step; CP.tau_count = tau_count;
    CP.q_key = q_key; CP.y_key_max = y_key_max; CP.q_err = q_err; CP.y_err_max = y_err_max;
    CP.eps_acc = eps_acc; CP.unverified = unverified;
    return refine_dispatch<true>(q, nq, ldq, y, ny, ldy, d, elem_size, q_sqnorm, y_sqnorm, cand_idx,
                                 ncand, metric, index_base, exclude_self, self_offset, k, out_dist,
                                 out_ind, CP, (cudaStream_t)stream);
}

extern "C" int kb2_topk_rows(const double *dist, const int64_t *ind, int64_t n, int c, int nparts,
                             int64_t part_stride, int k, double *out_dist, int64_t *out_ind,
                             void *stream) {
    using namespace kb2;
    KB2_CHECK(n >= 0 && c > 0 && nparts > 0, "topk_rows: bad shape");
    const int total = c * nparts;
    KB2_CHECK(total <= 2048, "topk_rows: %d candidates per row exceed 2048", total);
    KB2_CHECK(k > 0 && k <= total, "topk_rows: k=%d must be in (0, %d]", k, total);
    if (n == 0) return 0;
    if (total <= 256)
        return launch_topk_rows_rg(dist, ind, n, c, nparts, part_stride, k, out_dist, out_ind,
                                   (cudaStream_t)stream);
    const int P = next_pow2(total);
    const size_t smem = (size_t)REFINE_WARPS * P * 24;
    KB2_CUDA(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    topk_rows_kernel<<<(unsigned)ceil_div64(n, REFINE_WARPS), REFINE_WARPS * 32, smem,
                       (cudaStream_t)stream>>>(dist, ind, n, c, nparts, part_stride, P, k, out_dist,
                                               out_ind);
    KB2_LAUNCH_CHECK();
    return 0;
}
