// Hubness-reduction rescaling of the candidate distances, fused with the final
// top-k sort.  All kernels are float64 and HBM-bound: they stream the (n, c) candidate
// arrays once, gather 1-2 per-target statistics per element and write (n, k).  Rows of
// <= 16 values: one thread per row (rows_small_kernel); <= 256: a row per lane group in
// registers (rows_rg_kernel); longer: one warp per row in shared memory.
//   CSLS        kiez/hubness_reduction/csls.py:85-96
//   LS / NICDM  kiez/hubness_reduction/local_scaling.py:129-151
//   MP Gaussian kiez/hubness_reduction/mutual_proximity.py:166-183 (numpy branch:
//               nanmean / nanstd ddof=0 / scipy.stats.norm.sf)
//   MP empiric  kiez/hubness_reduction/mutual_proximity.py:185-212
//   DisSimLocal kiez/hubness_reduction/dis_sim.py:95-108,139-181
#include "rowgroup.cuh"

namespace kb2 {

constexpr int RS_WARPS = 4;

struct RowBuf {
    double *key;
    int64_t *tie;
    int64_t *pay;
};
__device__ __forceinline__ RowBuf row_buf(unsigned char *smem_raw, int warp, int P) {
    RowBuf b;
    b.key = reinterpret_cast<double *>(smem_raw) + (size_t)warp * P;
    b.tie = reinterpret_cast<int64_t *>(reinterpret_cast<double *>(smem_raw) + (size_t)RS_WARPS * P) +
            (size_t)warp * P;
    b.pay = reinterpret_cast<int64_t *>(reinterpret_cast<double *>(smem_raw) + (size_t)2 * RS_WARPS * P) +
            (size_t)warp * P;
    return b;
}

// nan-skipping mean / population std of a row held in smem (numpy nanmean/nanstd)
__device__ __forceinline__ void row_mean_sd(const double *v, int c, int lane, double &mean,
                                            double &sd) {
    double s = 0.0, cnt = 0.0;
    for (int j = lane; j < c; j += 32) {
        const double x = v[j];
        if (!isnan(x)) { s += x; cnt += 1.0; }
    }
    s = warp_sum(s);
    cnt = warp_sum(cnt);
    mean = s / cnt;
    double q = 0.0;
    for (int j = lane; j < c; j += 32) {
        const double x = v[j];
        if (!isnan(x)) { const double e = x - mean; q += e * e; }
    }
    q = warp_sum(q);
    sd = sqrt(q / cnt);
}

__device__ __forceinline__ double norm_sf(double x, double mu, double sd) {
    if (!(sd > 0.0)) return __longlong_as_double(0x7ff8000000000000LL);   // scipy: scale <= 0 -> nan
    return 0.5 * erfc(((x - mu) / sd) * 0.70710678118654752440);
}

// write the row either unsorted (k == 0: the HubnessReduction.transform contract)
// or as its k best after a bitonic sort keyed by (value, position)
__device__ __forceinline__ void emit_row(RowBuf b, int c, int P, int k, int64_t row, int lane,
                                         double *out_dist, int64_t *out_ind) {
    if (k == 0) {
        __syncwarp();
        for (int j = lane; j < c; j += 32) {
            out_dist[row * c + j] = b.key[j];
            out_ind[row * c + j] = b.pay[j];
        }
        return;
    }
    for (int j = c + lane; j < P; j += 32) {
        b.key[j] = __longlong_as_double(0x7ff8000000000000LL);
        b.tie[j] = INT64_MAX;
        b.pay[j] = -1;
    }
    warp_bitonic_sort(b.key, b.tie, b.pay, P, lane);
    for (int j = lane; j < k; j += 32) {
        out_dist[row * k + j] = b.key[j];
        out_ind[row * k + j] = b.pay[j];
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
row_stats_kernel(const double *__restrict__ dist, int64_t n, int c, double *__restrict__ mean,
                 double *__restrict__ sd, double *__restrict__ last) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (row >= n) return;
    const double *r = dist + row * c;
    // plain mean (CSLS/NICDM use ndarray.mean) and nan-aware mean/std (MP uses nanmean/nanstd)
    double s = 0.0, sn = 0.0, cnt = 0.0;
    for (int j = lane; j < c; j += 32) {
        const double x = r[j];
        s += x;
        if (!isnan(x)) { sn += x; cnt += 1.0; }
    }
    s = warp_sum(s);
    sn = warp_sum(sn);
    cnt = warp_sum(cnt);
    const double mu = sn / cnt;
    double q = 0.0;
    for (int j = lane; j < c; j += 32) {
        const double x = r[j];
        if (!isnan(x)) { const double e = x - mu; q += e * e; }
    }
    q = warp_sum(q);
    if (lane == 0) {
        // `mean` serves CSLS/NICDM when sd == nullptr, MutualProximity otherwise
        if (mean) mean[row] = sd ? mu : s / (double)c;
        if (sd) sd[row] = sqrt(q / cnt);
        if (last) last[row] = r[c - 1];
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
rescale_topk_kernel(int mode, const double *__restrict__ dist, const int64_t *__restrict__ ind,
                    int64_t n, int c, const double *__restrict__ stat_a,
                    const double *__restrict__ stat_b, int64_t n_stats, int P, int k,
                    double *__restrict__ out_dist, int64_t *__restrict__ out_ind) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (row >= n) return;
    RowBuf b = row_buf(smem_raw, warp, P);
    const double *dr = dist + row * c;
    const int64_t *ir = ind + row * c;
    double s = 0.0;
    for (int j = lane; j < c; j += 32) {
        const double x = dr[j];
        b.key[j] = x;
        b.pay[j] = ir[j];
        b.tie[j] = j;
        s += x;
    }
    s = warp_sum(s);
    __syncwarp();
    const double mean = s / (double)c;
    const double last = b.key[c - 1];
    double mu = 0.0, sd = 0.0;
    if (mode == KB2_RESCALE_MP_GAUSS) row_mean_sd(b.key, c, lane, mu, sd);
    __syncwarp();
    for (int j = lane; j < c; j += 32) {
        const double x = b.key[j];
        const int64_t id = b.pay[j];
        const bool ok = id >= 0 && id < n_stats;
        const double a = ok ? stat_a[id] : __longlong_as_double(0x7ff8000000000000LL);
        double r;
        if (mode == KB2_RESCALE_CSLS) {
            r = 2.0 * x - mean - a;
        } else if (mode == KB2_RESCALE_LS) {
            r = 1.0 - exp(-1.0 * (x * x) / (last * a));
        } else if (mode == KB2_RESCALE_NICDM) {
            r = x / sqrt(mean * a);
        } else {
            const double sb = ok ? stat_b[id] : __longlong_as_double(0x7ff8000000000000LL);
            r = 1.0 - norm_sf(x, mu, sd) * norm_sf(x, a, sb);
        }
        b.key[j] = r;
    }
    emit_row(b, c, P, k, row, lane, out_dist, out_ind);
}

// MutualProximity empiric: for candidate j of query i (target id cj):
//   d_j[l] = rev_dist[cj][p] if ind[i][l] == rev_ind[cj][p] (last p wins) else rev_dist[cj][-1]+1e-6
//   out[j] = 1 - #{l : d[i][l] > d[i][j] and d_j[l] > d[i][j]} / c
__global__ void __launch_bounds__(RS_WARPS * 32)
mp_empiric_kernel(const double *__restrict__ dist, const int64_t *__restrict__ ind, int64_t n, int c,
                  const double *__restrict__ rev_dist, const int64_t *__restrict__ rev_ind, int64_t m,
                  int c_rev, int P, int k, double *__restrict__ out_dist,
                  int64_t *__restrict__ out_ind) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (row >= n) return;
    RowBuf b = row_buf(smem_raw, warp, P);
    // extra per-warp scratch after the three RowBuf planes: d copy [P] + result [P]
    double *dcopy = reinterpret_cast<double *>(smem_raw) + (size_t)3 * RS_WARPS * P + (size_t)warp * 2 * P;
    double *res = dcopy + P;
    for (int j = lane; j < c; j += 32) {
        dcopy[j] = dist[row * c + j];
        b.pay[j] = ind[row * c + j];
        b.tie[j] = j;
    }
    __syncwarp();
    for (int j = 0; j < c; ++j) {
        const int64_t cj = b.pay[j];
        const double dij = dcopy[j];
        int count = 0;
        if (cj >= 0 && cj < m) {
            const double *rd = rev_dist + cj * c_rev;
            const int64_t *ri = rev_ind + cj * c_rev;
            const double fill = rd[c_rev - 1] + 1e-6;
            for (int l = lane; l < c; l += 32) {
                const int64_t tl = b.pay[l];
                double dj = fill;
                for (int p = 0; p < c_rev; ++p)
                    if (ri[p] == tl) dj = rd[p];
                count += (dcopy[l] > dij) && (dj > dij);
            }
        }
        count = __reduce_add_sync(FULL_MASK, count);
        if (lane == 0) res[j] = 1.0 - (double)count / (double)c;
    }
    __syncwarp();
    for (int j = lane; j < c; j += 32) b.key[j] = res[j];
    emit_row(b, c, P, k, row, lane, out_dist, out_ind);
}

// DisSimLocal._fit: centroid of the reverse neighbours + squared distance to it.
// One warp per target row; lanes stride over features, fp64 accumulation.
template <typename T>
__global__ void __launch_bounds__(RS_WARPS * 32)
dsl_fit_kernel(const T *__restrict__ source, int64_t n_source, int64_t lds,
               const T *__restrict__ target, int64_t m, int64_t ldt, int d,
               const int64_t *__restrict__ rev_ind, int c_rev, double *__restrict__ centroids,
               double *__restrict__ dist_to_cent) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (row >= m) return;
    const int64_t *ri = rev_ind + row * c_rev;
    double acc2 = 0.0;
    for (int t = lane; t < d; t += 32) {
        double s = 0.0;
        for (int p = 0; p < c_rev; ++p) {
            const int64_t id = ri[p];
            s += (id >= 0 && id < n_source) ? (double)source[id * lds + t] : 0.0;
        }
        const double cen = s / (double)c_rev;
        if (centroids) centroids[row * d + t] = cen;
        const double e = (double)target[row * ldt + t] - cen;
        acc2 += e * e;
    }
    acc2 = warp_sum(acc2);
    if (lane == 0) dist_to_cent[row] = acc2;
}

__device__ __forceinline__ void atomic_min_double(double *addr, double v) {
    unsigned long long *a = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *a;
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}

// DisSimLocal.transform stage 1.  The oracle computes ||q-t||^2 in the expanded
// form clamped at 0 (sklearn euclidean_distances); the direct fp64 difference
// used here is the same number to ~1e-13 relative and never negative.
// One warp per query row; lane l owns features l, l+32, ... (U per 32*U-wide
// chunk, in registers) so every candidate row is read once, coalesced, and feeds
// both the pairwise distance and the local centroid.
template <typename T, int U>
__global__ void __launch_bounds__(RS_WARPS * 32)
dsl_transform_kernel(const T *__restrict__ query, int64_t n, int64_t ldq,
                     const T *__restrict__ target, int64_t m, int64_t ldt, int d,
                     const int64_t *__restrict__ ind, int c, const double *__restrict__ dist_to_cent,
                     double *__restrict__ raw, double *__restrict__ global_min) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (row >= n) return;
    double *sq = reinterpret_cast<double *>(smem_raw) + (size_t)warp * c;   // [c] per warp
    const T *qr = query + row * ldq;
    const int64_t *ir = ind + row * c;
    for (int j = lane; j < c; j += 32) sq[j] = 0.0;
    __syncwarp();
    double qc2 = 0.0;   // ||q - centroid||^2, this lane's features
    for (int f0 = 0; f0 < d; f0 += 32 * U) {
        T qv[U];
        double csum[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = f0 + lane + 32 * u;
            qv[u] = (t < d) ? qr[t] : T(0);
            csum[u] = 0.0;
        }
        for (int j = 0; j < c; ++j) {
            const int64_t id = ir[j];
            const bool ok = id >= 0 && id < m;
            const T *tr = target + (ok ? id : 0) * ldt;
            double part = 0.0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = f0 + lane + 32 * u;
                const T tv = (ok && t < d) ? tr[t] : T(0);
                csum[u] += (double)tv;
                const double e = (double)qv[u] - (double)tv;
                part += e * e;
            }
            part = warp_sum(part);
            if (lane == 0) sq[j] += part;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = f0 + lane + 32 * u;
            if (t < d) {
                const double e = (double)qv[u] - csum[u] / (double)c;
                qc2 += e * e;
            }
        }
    }
    qc2 = warp_sum(qc2);
    __syncwarp();
    double mn = INFINITY;
    for (int j = lane; j < c; j += 32) {
        const int64_t id = ir[j];
        const double dc = (id >= 0 && id < m) ? dist_to_cent[id] : 0.0;
        const double v = sq[j] - qc2 - dc;
        raw[row * c + j] = v;
        mn = fmin(mn, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL_MASK, mn, o));
    if (lane == 0) atomic_min_double(global_min, mn);
}

__global__ void __launch_bounds__(RS_WARPS * 32)
dsl_finish_kernel(const double *__restrict__ raw, const int64_t *__restrict__ ind, int64_t n, int c,
                  const double *__restrict__ global_min, int squared, int P, int k,
                  double *__restrict__ out_dist, int64_t *__restrict__ out_ind) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RS_WARPS + warp;
    if (row >= n) return;
    RowBuf b = row_buf(smem_raw, warp, P);
    const double mn = *global_min;
    const double shift = (mn < 0.0) ? -mn : 0.0;
    for (int j = lane; j < c; j += 32) {
        double v = raw[row * c + j] + shift;
        if (!squared) v = sqrt(v);
        b.key[j] = v;
        b.pay[j] = ind[row * c + j];
        b.tie[j] = j;
    }
    emit_row(b, c, P, k, row, lane, out_dist, out_ind);
}

// ---------------------------------------------------------------------------
// register-resident fast path (c * nparts <= 256): see rowgroup.cuh
// ---------------------------------------------------------------------------
constexpr int RG_OP_TOPK = 4;        // plain row-wise top-k (final sort / multi-GPU merge)
constexpr int RG_OP_DSL_FINISH = 5;  // DisSimLocal: shift by the global minimum, sqrt
constexpr int RG_WARPS = 8;

struct RgParams {
    int op;                  // KB2_RESCALE_* or RG_OP_*
    const double *dist;
    const int64_t *ind;
    int64_t n;
    int c;                   // values per row and part
    int nparts;              // row r = concat over parts p of dist[p*part_stride + r*c .. +c)
    int64_t part_stride;
    const double *stat_a, *stat_b;
    int64_t n_stats;
    const double *gmin;
    int squared;
    int k;                   // 0: write the unsorted transform
    double *out_dist;
    int64_t *out_ind;
};

// Order-preserving 64-bit integer image of a double: unsigned order == (value ascending, NaN
// last); -0.0 is folded onto +0.0 first so that equal values stay ties (broken by position).
__device__ __forceinline__ unsigned long long sortable_bits(double v) {
    if (isnan(v)) return ~0ull;                           // NaN last (numpy order), ties by position
    v += 0.0;                                             // -0.0 -> +0.0
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// The k smallest (value, position) pairs of a row held by a full warp (E entries per lane,
// element e = t * 32 + lane), written in order: k rounds of a warp-wide arg-min on the integer
// images -- three redux.sync (high word, low word, position) per round, the winner stores its
// entry and retires it.  For the k <= 16 of 17..256 candidates that kiez asks for this is a third
// of the instructions of the bitonic sort over lanes x registers (which made the c = 50 / 100
// rescale kernels instruction-bound at 0.10-0.13 of the HBM roofline).
constexpr int RG_SELECT_MAX_K = 16;
template <int E>
__device__ __forceinline__ void warp_select_topk(const double (&r)[E], const int64_t (&id)[E],
                                                 int total, bool row_ok, int lane, int k,
                                                 double *out_dist, int64_t *out_ind) {
    unsigned long long u[E];
    unsigned int pos[E];
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const int e = t * 32 + lane;
        const bool ok = row_ok && e < total;
        u[t] = ok ? sortable_bits(r[t]) : ~0ull;
        pos[t] = ok ? (unsigned int)e : 0xFFFFFFFFu;
    }
    for (int j = 0; j < k; ++j) {
        unsigned long long ub = u[0];
        unsigned int pb = pos[0];
#pragma unroll
        for (int t = 1; t < E; ++t) {
            const bool better = u[t] < ub || (u[t] == ub && pos[t] < pb);
            ub = better ? u[t] : ub;
            pb = better ? pos[t] : pb;
        }
        const unsigned int hi = (unsigned int)(ub >> 32), lo = (unsigned int)ub;
        const unsigned int mh = __reduce_min_sync(FULL_MASK, hi);
        const unsigned int ml = __reduce_min_sync(FULL_MASK, hi == mh ? lo : 0xFFFFFFFFu);
        const bool match = hi == mh && lo == ml;
        const unsigned int mp = __reduce_min_sync(FULL_MASK, match ? pb : 0xFFFFFFFFu);
        if (match && pb == mp && mp != 0xFFFFFFFFu) {      // the one lane that holds the winner
#pragma unroll
            for (int t = 0; t < E; ++t) {
                if (pos[t] == pb) {
                    out_dist[j] = r[t];
                    out_ind[j] = id[t];
                    u[t] = ~0ull;
                    pos[t] = 0xFFFFFFFFu;
                }
            }
        }
    }
}

template <int G, int E>
__global__ void __launch_bounds__(RG_WARPS * 32)
rows_rg_kernel(const RgParams p) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const int64_t row = ((int64_t)blockIdx.x * RG_WARPS + (threadIdx.x >> 5)) * (32 / G) + lane / G;
    const bool row_ok = row < p.n;
    const int total = p.c * p.nparts;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double x[E];
    int64_t off[E];
    int64_t id[E];
    double s = 0.0, sn = 0.0, cnt = 0.0;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const int e = t * G + gl;
        const bool ok = row_ok && e < total;
        const int part = ok ? e / p.c : 0;
        off[t] = (int64_t)part * p.part_stride + row * p.c + (e - part * p.c);
        x[t] = ok ? p.dist[off[t]] : qnan;
        id[t] = ok ? p.ind[off[t]] : -1;
        if (ok) {
            s += x[t];
            if (!isnan(x[t])) { sn += x[t]; cnt += 1.0; }
        }
    }
    double r[E];
    if (p.op <= KB2_RESCALE_MP_GAUSS) {
        const double mean = group_sum<G>(s) / (double)p.c;               // ndarray.mean
        double mu = 0.0, sd = 0.0, last = 0.0;
        if (p.op == KB2_RESCALE_LS) last = group_element<G, E>(x, p.c - 1, lane);
        if (p.op == KB2_RESCALE_MP_GAUSS) {                              // nanmean / nanstd(ddof=0)
            const double nn = group_sum<G>(cnt);
            mu = group_sum<G>(sn) / nn;
            double q = 0.0;
#pragma unroll
            for (int t = 0; t < E; ++t)
                if (!isnan(x[t])) { const double d = x[t] - mu; q += d * d; }
            sd = sqrt(group_sum<G>(q) / nn);
        }
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const bool ok = id[t] >= 0 && id[t] < p.n_stats;
            const double a = ok ? p.stat_a[id[t]] : qnan;
            if (p.op == KB2_RESCALE_CSLS) {
                r[t] = 2.0 * x[t] - mean - a;
            } else if (p.op == KB2_RESCALE_LS) {
                r[t] = 1.0 - exp(-1.0 * (x[t] * x[t]) / (last * a));
            } else if (p.op == KB2_RESCALE_NICDM) {
                r[t] = x[t] / sqrt(mean * a);
            } else {
                const double sb = ok ? p.stat_b[id[t]] : qnan;
                r[t] = 1.0 - norm_sf(x[t], mu, sd) * norm_sf(x[t], a, sb);
            }
        }
    } else if (p.op == RG_OP_DSL_FINISH) {
        const double mn = *p.gmin;
        const double shift = (mn < 0.0) ? -mn : 0.0;
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const double v = x[t] + shift;
            r[t] = p.squared ? v : sqrt(v);
        }
    } else {
#pragma unroll
        for (int t = 0; t < E; ++t) r[t] = x[t];
    }
    if (p.k == 0) {                                      // HubnessReduction.transform: unsorted
#pragma unroll
        for (int t = 0; t < E; ++t) {
            const int e = t * G + gl;
            if (row_ok && e < total) {
                p.out_dist[row * total + e] = r[t];
                p.out_ind[row * total + e] = id[t];
            }
        }
        return;
    }
    if constexpr (G == 32) {
        if (p.k <= RG_SELECT_MAX_K) {                    // warp-uniform
            warp_select_topk<E>(r, id, total, row_ok, lane, p.k, p.out_dist + row * p.k,
                                p.out_ind + row * p.k);
            return;
        }
    }
    int pos[E];
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const int e = t * G + gl;
        const bool ok = row_ok && e < total;
        pos[t] = ok ? e : RG_POS_PAD;
        if (!ok) r[t] = qnan;
    }
    group_sort<G, E>(r, pos, gl);
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const int e = t * G + gl;                        // rank after the sort
        if (row_ok && e < p.k) {
            const int src = pos[t];
            const int part = src / p.c;
            p.out_dist[row * p.k + e] = r[t];
            p.out_ind[row * p.k + e] =
                p.ind[(int64_t)part * p.part_stride + row * p.c + (src - part * p.c)];
        }
    }
}

template <int G, int E>
__global__ void __launch_bounds__(RG_WARPS * 32)
row_stats_rg_kernel(const double *__restrict__ dist, int64_t n, int c, double *__restrict__ mean,
                    double *__restrict__ sd, double *__restrict__ last) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const int64_t row = ((int64_t)blockIdx.x * RG_WARPS + (threadIdx.x >> 5)) * (32 / G) + lane / G;
    const bool row_ok = row < n;
    double x[E];
    double s = 0.0, sn = 0.0, cnt = 0.0;
#pragma unroll
    for (int t = 0; t < E; ++t) {
        const int e = t * G + gl;
        const bool ok = row_ok && e < c;
        x[t] = ok ? dist[row * c + e] : __longlong_as_double(0x7ff8000000000000LL);
        if (ok) {
            s += x[t];
            if (!isnan(x[t])) { sn += x[t]; cnt += 1.0; }
        }
    }
    s = group_sum<G>(s);
    const double nn = group_sum<G>(cnt);
    const double mu = group_sum<G>(sn) / nn;
    double q = 0.0;
#pragma unroll
    for (int t = 0; t < E; ++t)
        if (!isnan(x[t])) { const double d = x[t] - mu; q += d * d; }
    q = group_sum<G>(q);
    const double lst = group_element<G, E>(x, c - 1, lane);
    if (row_ok && gl == 0) {
        // `mean` serves CSLS/NICDM (plain mean) when sd == nullptr, MutualProximity otherwise
        if (mean) mean[row] = sd ? mu : s / (double)c;
        if (sd) sd[row] = sqrt(q / nn);
        if (last) last[row] = lst;
    }
}

// ---------------------------------------------------------------------------
// narrow rows (c * nparts <= 16, e.g. kiez's default n_candidates = 10): ONE THREAD per row.
// A 160-byte row does not feed a lane group: the 16-entry bitonic network of the row-group
// path (shuffles of (double, int) pairs, NaN-aware compares) made those kernels instruction
// bound at 0.14-0.20 of the HBM roofline.  Here a thread keeps its row in registers, orders it
// with a branch-free rank sort on order-preserving 64-bit keys (C^2 integer compares, no
// shuffles, no divergence) and scatters the k best straight to their final positions; a warp
// covers 32 consecutive rows, so its loads and stores touch one contiguous span.
// ---------------------------------------------------------------------------

// Block = 128 consecutive rows.  Global memory is only touched with coalesced accesses: the
// block's contiguous span of the (n, c) arrays is copied into shared memory (row stride padded
// to an odd number of 8-byte words: conflict-free per-row reads), every thread takes its row
// into registers, and the results go back through the same buffers.
constexpr int SMALL_ROWS = 128;
__host__ __device__ inline int small_stride(int width) { return width | 1; }

// 8-byte asynchronous copy global -> shared (LDGSTS): the staging loop below issues every load of
// the block's span back to back -- nothing waits in registers for a store, so one DRAM round trip
// covers the whole span (the register-staged loop of the first version took 2.5 of them and the
// kernel sat at half the HBM roofline, latency-bound at 30 % occupancy).
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// C = exact row width handled in registers (>= c * nparts), OP = KB2_RESCALE_* / RG_OP_* (a
// template parameter: the CSLS instance does not pay the registers of the erfc path).
template <int C, int OP>
__global__ void __launch_bounds__(SMALL_ROWS, (C <= 10 && OP != KB2_RESCALE_MP_GAUSS) ? 7 : 1)
rows_small_kernel(const RgParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int total = p.c * p.nparts;
    const int ld = small_stride(total);
    double *sd = reinterpret_cast<double *>(smem_raw);                       // [SMALL_ROWS][ld]
    int64_t *si = reinterpret_cast<int64_t *>(sd + (size_t)SMALL_ROWS * ld); // [SMALL_ROWS][ld]
    const int64_t row0 = (int64_t)blockIdx.x * SMALL_ROWS;
    const int rows_here = (int)min((int64_t)SMALL_ROWS, p.n - row0);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int part = 0; part < p.nparts; ++part) {
        const double *gd = p.dist + (int64_t)part * p.part_stride + row0 * p.c;
        const int64_t *gi = p.ind + (int64_t)part * p.part_stride + row0 * p.c;
        for (int e = threadIdx.x; e < rows_here * p.c; e += SMALL_ROWS) {
            const int r = e / p.c, j = e - r * p.c;
            cp_async8(sd + r * ld + part * p.c + j, gd + e);
            cp_async8(si + r * ld + part * p.c + j, gi + e);
        }
    }
    cp_async_wait_all();
    __syncthreads();
    const bool row_ok = threadIdx.x < rows_here;
    double x[C];
    int64_t id[C];
    double s = 0.0, sn = 0.0, cnt = 0.0;
#pragma unroll
    for (int e = 0; e < C; ++e) {
        x[e] = qnan;
        id[e] = -1;
        if (row_ok && e < total) {
            x[e] = sd[threadIdx.x * ld + e];
            id[e] = si[threadIdx.x * ld + e];
            s += x[e];
            if (OP == KB2_RESCALE_MP_GAUSS && !isnan(x[e])) { sn += x[e]; cnt += 1.0; }
        }
    }
    double r[C];
    if constexpr (OP <= KB2_RESCALE_MP_GAUSS) {
        const double mean = s / (double)p.c;                              // ndarray.mean
        double mu = 0.0, sd_row = 0.0, last = 0.0;
        if (OP == KB2_RESCALE_LS && row_ok) last = sd[threadIdx.x * ld + p.c - 1];
        if constexpr (OP == KB2_RESCALE_MP_GAUSS) {                       // nanmean / nanstd(ddof=0)
            mu = sn / cnt;
            double q = 0.0;
#pragma unroll
            for (int e = 0; e < C; ++e)
                if (!isnan(x[e])) { const double d = x[e] - mu; q += d * d; }
            sd_row = sqrt(q / cnt);
        }
#pragma unroll
        for (int e = 0; e < C; ++e) {
            const bool ok = id[e] >= 0 && id[e] < p.n_stats;
            const double a = ok ? __ldg(p.stat_a + id[e]) : qnan;
            if constexpr (OP == KB2_RESCALE_CSLS) {
                r[e] = 2.0 * x[e] - mean - a;
            } else if constexpr (OP == KB2_RESCALE_LS) {
                r[e] = 1.0 - exp(-1.0 * (x[e] * x[e]) / (last * a));
            } else if constexpr (OP == KB2_RESCALE_NICDM) {
                r[e] = x[e] / sqrt(mean * a);
            } else {
                const double sb = ok ? __ldg(p.stat_b + id[e]) : qnan;
                r[e] = 1.0 - norm_sf(x[e], mu, sd_row) * norm_sf(x[e], a, sb);
            }
        }
    } else if constexpr (OP == RG_OP_DSL_FINISH) {
        const double mn = *p.gmin;
        const double shift = (mn < 0.0) ? -mn : 0.0;
#pragma unroll
        for (int e = 0; e < C; ++e) {
            const double v = x[e] + shift;
            r[e] = p.squared ? v : sqrt(v);
        }
    } else {
#pragma unroll
        for (int e = 0; e < C; ++e) r[e] = x[e];
    }
    __syncthreads();                                     // every row is in registers: reuse the buffers
    const int width = p.k == 0 ? total : p.k;            // values written per row
    if (p.k == 0) {                                      // HubnessReduction.transform: unsorted
#pragma unroll
        for (int e = 0; e < C; ++e) {
            if (row_ok && e < total) {
                sd[threadIdx.x * ld + e] = r[e];
                si[threadIdx.x * ld + e] = id[e];
            }
        }
    } else {
        unsigned long long u[C];
#pragma unroll
        for (int e = 0; e < C; ++e) u[e] = (e < total) ? sortable_bits(r[e]) : ~0ull;
#pragma unroll
        for (int e = 0; e < C; ++e) {
            if (e >= total) continue;
            int rank = 0;                                // entries ordered before e: (value, position)
#pragma unroll
            for (int j = 0; j < C; ++j) {
                if (j == e) continue;
                rank += (j < e) ? (u[j] <= u[e]) : (u[j] < u[e]);
            }
            if (row_ok && rank < p.k) {
                sd[threadIdx.x * ld + rank] = r[e];
                si[threadIdx.x * ld + rank] = id[e];
            }
        }
    }
    __syncthreads();
    double *od = p.out_dist + row0 * width;
    int64_t *oi = p.out_ind + row0 * width;
#pragma unroll 4
    for (int e = threadIdx.x; e < rows_here * width; e += SMALL_ROWS) {
        const int rr = e / width, j = e - rr * width;
        od[e] = sd[rr * ld + j];
        oi[e] = si[rr * ld + j];
    }
}

template <int C>
__global__ void __launch_bounds__(SMALL_ROWS)
row_stats_small_kernel(const double *__restrict__ dist, int64_t n, int c, double *__restrict__ mean,
                       double *__restrict__ sd, double *__restrict__ last) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ld = small_stride(c);
    double *sx = reinterpret_cast<double *>(smem_raw);                       // [SMALL_ROWS][ld]
    const int64_t row0 = (int64_t)blockIdx.x * SMALL_ROWS;
    const int rows_here = (int)min((int64_t)SMALL_ROWS, n - row0);
    const double *gd = dist + row0 * c;
    for (int e = threadIdx.x; e < rows_here * c; e += SMALL_ROWS) {
        const int r = e / c;
        cp_async8(sx + r * ld + (e - r * c), gd + e);
    }
    cp_async_wait_all();
    __syncthreads();
    if ((int)threadIdx.x >= rows_here) return;
    const int64_t row = row0 + threadIdx.x;
    double x[C];
    double s = 0.0, sn = 0.0, cnt = 0.0;
#pragma unroll
    for (int e = 0; e < C; ++e) {
        x[e] = __longlong_as_double(0x7ff8000000000000LL);
        if (e < c) {
            x[e] = sx[threadIdx.x * ld + e];
            s += x[e];
            if (!isnan(x[e])) { sn += x[e]; cnt += 1.0; }
        }
    }
    const double lst = sx[threadIdx.x * ld + c - 1];
    const double mu = sn / cnt;
    double q = 0.0;
#pragma unroll
    for (int e = 0; e < C; ++e)
        if (!isnan(x[e])) { const double d = x[e] - mu; q += d * d; }
    // `mean` serves CSLS/NICDM (plain mean) when sd == nullptr, MutualProximity otherwise
    if (mean) mean[row] = sd ? mu : s / (double)c;
    if (sd) sd[row] = sqrt(q / cnt);
    if (last) last[row] = lst;
}

constexpr int SMALL_MAX_WIDTH = 16;
template <int OP>
static int launch_rows_small_op(const RgParams &p, cudaStream_t st) {
    const int total = p.c * p.nparts;
    const unsigned grid = (unsigned)ceil_div64(p.n, SMALL_ROWS);
    const size_t smem = (size_t)SMALL_ROWS * small_stride(total) * 16;
    if (total <= 8) rows_small_kernel<8, OP><<<grid, SMALL_ROWS, smem, st>>>(p);
    else if (total <= 10) rows_small_kernel<10, OP><<<grid, SMALL_ROWS, smem, st>>>(p);   // kiez's default
    else if (total <= 12) rows_small_kernel<12, OP><<<grid, SMALL_ROWS, smem, st>>>(p);
    else rows_small_kernel<16, OP><<<grid, SMALL_ROWS, smem, st>>>(p);
    KB2_LAUNCH_CHECK();
    return 0;
}
static int launch_rows_small(const RgParams &p, cudaStream_t st) {
    switch (p.op) {
        case KB2_RESCALE_CSLS: return launch_rows_small_op<KB2_RESCALE_CSLS>(p, st);
        case KB2_RESCALE_LS: return launch_rows_small_op<KB2_RESCALE_LS>(p, st);
        case KB2_RESCALE_NICDM: return launch_rows_small_op<KB2_RESCALE_NICDM>(p, st);
        case KB2_RESCALE_MP_GAUSS: return launch_rows_small_op<KB2_RESCALE_MP_GAUSS>(p, st);
        case RG_OP_DSL_FINISH: return launch_rows_small_op<RG_OP_DSL_FINISH>(p, st);
        default: return launch_rows_small_op<RG_OP_TOPK>(p, st);
    }
}

// dispatch on the row width: G lanes x E registers >= width
template <template <int, int> class Launcher, class... Args>
static int rg_dispatch(int width, Args... args) {
    if (width <= 8) return Launcher<8, 1>::run(args...);
    if (width <= 16) return Launcher<16, 1>::run(args...);
    if (width <= 32) return Launcher<32, 1>::run(args...);
    if (width <= 64) return Launcher<32, 2>::run(args...);
    if (width <= 128) return Launcher<32, 4>::run(args...);
    return Launcher<32, 8>::run(args...);
}
template <int G, int E>
struct RowsLauncher {
    static int run(const RgParams &p, cudaStream_t st) {
        if (p.c * p.nparts <= SMALL_MAX_WIDTH) return launch_rows_small(p, st);
        const int64_t rows_per_block = (int64_t)RG_WARPS * (32 / G);
        rows_rg_kernel<G, E><<<(unsigned)ceil_div64(p.n, rows_per_block), RG_WARPS * 32, 0, st>>>(p);
        KB2_LAUNCH_CHECK();
        return 0;
    }
};
template <int G, int E>
struct StatsLauncher {
    static int run(const double *dist, int64_t n, int c, double *mean, double *sd, double *last,
                   cudaStream_t st) {
        if (c <= SMALL_MAX_WIDTH) {
            const unsigned grid = (unsigned)ceil_div64(n, SMALL_ROWS);
            const size_t smem = (size_t)SMALL_ROWS * small_stride(c) * 8;
            if (c <= 8) row_stats_small_kernel<8><<<grid, SMALL_ROWS, smem, st>>>(dist, n, c, mean, sd, last);
            else if (c <= 10) row_stats_small_kernel<10><<<grid, SMALL_ROWS, smem, st>>>(dist, n, c, mean, sd, last);
            else if (c <= 12) row_stats_small_kernel<12><<<grid, SMALL_ROWS, smem, st>>>(dist, n, c, mean, sd, last);
            else row_stats_small_kernel<16><<<grid, SMALL_ROWS, smem, st>>>(dist, n, c, mean, sd, last);
            KB2_LAUNCH_CHECK();
            return 0;
        }
        const int64_t rows_per_block = (int64_t)RG_WARPS * (32 / G);
        row_stats_rg_kernel<G, E><<<(unsigned)ceil_div64(n, rows_per_block), RG_WARPS * 32, 0, st>>>(
            dist, n, c, mean, sd, last);
        KB2_LAUNCH_CHECK();
        return 0;
    }
};
constexpr int RG_MAX_WIDTH = 256;

int launch_topk_rows_rg(const double *dist, const int64_t *ind, int64_t n, int c, int nparts,
                        int64_t part_stride, int k, double *out_dist, int64_t *out_ind,
                        cudaStream_t st) {
    RgParams p{};
    p.op = RG_OP_TOPK; p.dist = dist; p.ind = ind; p.n = n; p.c = c; p.nparts = nparts;
    p.part_stride = part_stride; p.k = k; p.out_dist = out_dist; p.out_ind = out_ind;
    return rg_dispatch<RowsLauncher>(c * nparts, p, st);
}

}  // namespace kb2

using namespace kb2;

#define RS_GRID(n) (unsigned)ceil_div64((n), RS_WARPS), RS_WARPS * 32

extern "C" int kb2_row_stats(const double *dist, int64_t n, int c, double *mean, double *sd,
                             double *last, void *stream) {
    KB2_CHECK(n >= 0 && c > 0, "row_stats: bad shape");
    if (n == 0) return 0;
    if (c <= RG_MAX_WIDTH)
        return rg_dispatch<StatsLauncher>(c, dist, n, c, mean, sd, last, (cudaStream_t)stream);
    row_stats_kernel<<<RS_GRID(n), 0, (cudaStream_t)stream>>>(dist, n, c, mean, sd, last);
    KB2_LAUNCH_CHECK();
    return 0;
}

static int check_row_args(const char *what, int64_t n, int c, int k) {
    KB2_CHECK(n >= 0 && c > 0 && c <= 2048, "%s: c=%d outside (0, 2048]", what, c);
    KB2_CHECK(k >= 0 && k <= c, "%s: k=%d must be in [0, c=%d]", what, k, c);
    return 0;
}

extern "C" int kb2_rescale_topk(int mode, const double *dist, const int64_t *ind, int64_t n, int c,
                                const double *stat_a, const double *stat_b, int64_t n_stats, int k,
                                double *out_dist, int64_t *out_ind, void *stream) {
    if (check_row_args("rescale_topk", n, c, k)) return 1;
    KB2_CHECK(mode >= 0 && mode <= 3, "rescale_topk: unknown mode %d", mode);
    KB2_CHECK(stat_a != nullptr && (mode != KB2_RESCALE_MP_GAUSS || stat_b != nullptr),
              "rescale_topk: missing per-target statistics");
    if (n == 0) return 0;
    if (c <= RG_MAX_WIDTH) {
        RgParams p{};
        p.op = mode; p.dist = dist; p.ind = ind; p.n = n; p.c = c; p.nparts = 1; p.part_stride = 0;
        p.stat_a = stat_a; p.stat_b = stat_b; p.n_stats = n_stats; p.k = k;
        p.out_dist = out_dist; p.out_ind = out_ind;
        return rg_dispatch<RowsLauncher>(c, p, (cudaStream_t)stream);
    }
    const int P = next_pow2(c);
    const size_t smem = (size_t)RS_WARPS * P * 24;
    KB2_CUDA(cudaFuncSetAttribute(rescale_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    rescale_topk_kernel<<<RS_GRID(n), smem, (cudaStream_t)stream>>>(
        mode, dist, ind, n, c, stat_a, stat_b, n_stats, P, k, out_dist, out_ind);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_mp_empiric_topk(const double *dist, const int64_t *ind, int64_t n, int c,
                                   const double *rev_dist, const int64_t *rev_ind, int64_t m,
                                   int c_rev, int k, double *out_dist, int64_t *out_ind,
                                   void *stream) {
    if (check_row_args("mp_empiric_topk", n, c, k)) return 1;
    KB2_CHECK(m > 0 && c_rev > 0, "mp_empiric_topk: bad reverse shape");
    if (n == 0) return 0;
    const int P = next_pow2(c);
    const size_t smem = (size_t)RS_WARPS * P * (24 + 16);
    KB2_CUDA(cudaFuncSetAttribute(mp_empiric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    mp_empiric_kernel<<<RS_GRID(n), smem, (cudaStream_t)stream>>>(
        dist, ind, n, c, rev_dist, rev_ind, m, c_rev, P, k, out_dist, out_ind);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_dsl_fit(const void *source, int64_t n_source, int64_t lds, const void *target,
                           int64_t m, int64_t ldt, int d, int elem_size, const int64_t *rev_ind,
                           int c_rev, double *centroids, double *dist_to_cent, void *stream) {
    KB2_CHECK(m >= 0 && d > 0 && c_rev > 0 && lds >= d && ldt >= d, "dsl_fit: bad shape");
    KB2_CHECK(elem_size == 4 || elem_size == 8, "dsl_fit: elem_size must be 4 or 8");
    if (m == 0) return 0;
    if (elem_size == 4)
        dsl_fit_kernel<float><<<RS_GRID(m), 0, (cudaStream_t)stream>>>(
            (const float *)source, n_source, lds, (const float *)target, m, ldt, d, rev_ind, c_rev,
            centroids, dist_to_cent);
    else
        dsl_fit_kernel<double><<<RS_GRID(m), 0, (cudaStream_t)stream>>>(
            (const double *)source, n_source, lds, (const double *)target, m, ldt, d, rev_ind, c_rev,
            centroids, dist_to_cent);
    KB2_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int launch_dsl_transform(const void *query, int64_t n, int64_t ldq, const void *target,
                                int64_t m, int64_t ldt, int d, const int64_t *ind, int c,
                                const double *dist_to_cent, double *raw, double *global_min,
                                cudaStream_t st) {
    const size_t smem = (size_t)RS_WARPS * c * sizeof(double);
    if (d <= 128)
        dsl_transform_kernel<T, 4><<<RS_GRID(n), smem, st>>>((const T *)query, n, ldq, (const T *)target,
                                                            m, ldt, d, ind, c, dist_to_cent, raw,
                                                            global_min);
    else
        dsl_transform_kernel<T, 8><<<RS_GRID(n), smem, st>>>((const T *)query, n, ldq, (const T *)target,
                                                            m, ldt, d, ind, c, dist_to_cent, raw,
                                                            global_min);
    KB2_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb2_dsl_transform(const void *query, int64_t n, int64_t ldq, const void *target,
                                 int64_t m, int64_t ldt, int d, int elem_size, const int64_t *ind,
                                 int c, const double *dist_to_cent, double *raw, double *global_min,
                                 void *stream) {
    KB2_CHECK(n >= 0 && d > 0 && c > 0 && c <= 2048 && ldq >= d && ldt >= d, "dsl_transform: bad shape");
    KB2_CHECK(elem_size == 4 || elem_size == 8, "dsl_transform: elem_size must be 4 or 8");
    if (n == 0) return 0;
    if (elem_size == 4)
        return launch_dsl_transform<float>(query, n, ldq, target, m, ldt, d, ind, c, dist_to_cent, raw,
                                           global_min, (cudaStream_t)stream);
    return launch_dsl_transform<double>(query, n, ldq, target, m, ldt, d, ind, c, dist_to_cent, raw,
                                        global_min, (cudaStream_t)stream);
}

extern "C" int kb2_dsl_finish_topk(const double *raw, const int64_t *ind, int64_t n, int c,
                                   const double *global_min, int squared, int k, double *out_dist,
                                   int64_t *out_ind, void *stream) {
    if (check_row_args("dsl_finish_topk", n, c, k)) return 1;
    if (n == 0) return 0;
    if (c <= RG_MAX_WIDTH) {
        RgParams p{};
        p.op = RG_OP_DSL_FINISH; p.dist = raw; p.ind = ind; p.n = n; p.c = c; p.nparts = 1;
        p.gmin = global_min; p.squared = squared; p.k = k; p.out_dist = out_dist; p.out_ind = out_ind;
        return rg_dispatch<RowsLauncher>(c, p, (cudaStream_t)stream);
    }
    const int P = next_pow2(c);
    const size_t smem = (size_t)RS_WARPS * P * 24;
    KB2_CUDA(cudaFuncSetAttribute(dsl_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    dsl_finish_kernel<<<RS_GRID(n), smem, (cudaStream_t)stream>>>(raw, ind, n, c, global_min, squared,
                                                                P, k, out_dist, out_ind);
    KB2_LAUNCH_CHECK();
    return 0;
}
