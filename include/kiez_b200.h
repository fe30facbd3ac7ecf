/*
 * kiez_b200 -- C ABI of the B200 (sm_100a) exact-kNN + hubness-reduction hot path.
 *
 * The reference (dobraczka/kiez v0.5.0) is pure Python: it has no FFI.  Its
 * boundary for this path is the NNAlgorithm / HubnessReduction plugin interface
 * (kiez/neighbors/neighbor_algorithm_base.py:13-136,
 *  kiez/hubness_reduction/base.py:17-105).  The entry points below are what the
 * thin Python host classes in kiez_b200/ bind with ctypes; each one names the
 * reference code it replaces.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: extern "C"; raw DEVICE pointers + sizes; the caller owns every
 * buffer; no hidden allocation except where stated; everything is enqueued on
 * `stream` (a cudaStream_t passed as void*) without host synchronisation unless
 * stated; return value 0 = ok, non-zero = error with a thread-local message in
 * kb2_last_error().  Row-major, C-contiguous unless a leading dimension is given.
 * Distances cross the boundary as float64, neighbour ids as int64 -- the dtypes
 * the reference's SklearnNN path returns.
 */
#ifndef KIEZ_B200_H
#define KIEZ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB2_VERSION 4

/* metric codes (kiez SklearnNN metric names; minkowski/l2 are p=2 euclidean) */
#define KB2_METRIC_EUCLIDEAN   0
#define KB2_METRIC_SQEUCLIDEAN 1
#define KB2_METRIC_COSINE      2

/* rescale modes */
#define KB2_RESCALE_CSLS      0 /* kiez/hubness_reduction/csls.py:85-96            */
#define KB2_RESCALE_LS        1 /* local_scaling.py:135-140 (method ls/standard)   */
#define KB2_RESCALE_NICDM     2 /* local_scaling.py:142-147 (method nicdm)         */
#define KB2_RESCALE_MP_GAUSS  3 /* mutual_proximity.py:166-183, numpy branch       */

/* which candidate-search kernel kb2_knn_candidates runs */
#define KB2_KNN_AUTO  0 /* = KB2_KNN_TC */
#define KB2_KNN_TC    1 /* tcgen05/TMEM 3xTF32 tiles fed by TMA, CTA pairs (cta_group::2) */
#define KB2_KNN_SIMT  2 /* fp32 FFMA tiles; cross-check + debugging aid          */
#define KB2_KNN_TC1   3 /* tcgen05, one CTA per tile (cross-check of the pair kernel)    */

int kb2_version(void);
const char *kb2_last_error(void);

/* Sizing helpers of the kNN stage (no counterpart in the reference: scikit-learn sizes its own
 * chunks behind sklearn_nearest_neighbors.py:96-101).
 *   kb2_max_candidates: widest candidate list (per query row, per split) the search kernels support
 *   kb2_padded_dim:     padded feature count the prepared operands use for a raw feature count d
 *   kb2_suggest_splits: number of index splits for nq queries (fills the SMs when nq is small) */
int kb2_max_candidates(void);
int kb2_padded_dim(int d);
int kb2_suggest_splits(int64_t nq, int64_t ny, int cap, int sm_count);

/*
 * "Index build" -- replaces SklearnNN._fit / NearestNeighbors.fit
 * (kiez/neighbors/exact/sklearn_nearest_neighbors.py:83-94), which for brute
 * force only keeps the matrix.  Produces the operands the tensor-core search
 * reads: the 3xTF32 split  x = hi + lo  (hi = rn_tf32(x), lo = rn_tf32(x-hi)),
 * zero-padded to dpad columns, plus the per-row selection term
 * (||x||^2 for euclidean, 0 for cosine where rows are L2-normalised first).
 *   x        [n][ldx] fp32 (first d columns used)
 *   center   [d] fp32 or NULL: subtracted from every row first (distances are
 *            translation invariant; keeps the expanded form well conditioned)
 *   hi, lo   [n][dpad] fp32 out
 *   key_term [n] fp32 out
 *   sqnorm   [n] fp64 out or NULL: exact ||x||^2 of the *raw* rows (cosine finish)
 */
int kb2_prepare_rows(const float *x, int64_t n, int d, int64_t ldx, const float *center,
                     int metric, float *hi, float *lo, int dpad, float *key_term,
                     double *sqnorm, void *stream);

/*
 * Candidate search -- the contraction of ArgKmin / pairwise_distances_chunked
 * behind NearestNeighbors.kneighbors (sklearn_nearest_neighbors.py:96-101,
 * reached from neighbor_algorithm_base.py:116-136) without ever materialising
 * the nq x ny matrix.  For every query row it keeps the `cap` smallest
 * selection keys  key = y_key[col] - 2 <q,y_col>  per index split.
 *   q_hi,q_lo [nq][dpad], y_hi,y_lo [ny][dpad], y_key [ny]
 *   splits    >=1: the index rows are cut into `splits` contiguous ranges, each
 *             searched by its own CTA set (fills the GPU when nq is small)
 *   cand_idx  [nq][splits*cap] int32 out: LOCAL index row ids, -1 = empty slot
 *   cand_key  [nq][splits*cap] fp32 out or NULL (approximate keys; tests only)
 */
int kb2_knn_candidates(int impl, const float *q_hi, const float *q_lo, int64_t nq,
                       const float *y_hi, const float *y_lo, const float *y_key,
                       int64_t ny, int dpad, int cap, int splits, int32_t *cand_idx,
                       float *cand_key, void *stream);

/*
 * Dual-direction candidate search: one pass over the x (rows) x y (columns) tiles yields the
 * row-wise candidate lists (as kb2_knn_candidates with queries = x, index = y) AND, per
 * column, every row whose column key  x_key[row] - 2 <x,y>  is below tau_col[col], appended
 * to col_buf[col][..col_cap) as packed (order-preserving key bits << 32 | row) with the count
 * in col_cnt[col] (zeroed by the caller; counts above col_cap mean the column overflowed and
 * must be searched separately).  tau_col must bound the column's final cap-th best key from
 * above, e.g. the cap-th best key against a sample of the rows.  Serves kiez's reverse pass
 * (hubness_reduction/base.py:37-42) and forward pass (:92-94) from ONE contraction.
 * The rows may be passed in SEGMENTS (one call per contiguous row range, row_id_base = its
 * first row: added to the emitted row ids) with kb2_col_compact between the calls, which
 * tightens tau_col to the cap-th best key seen so far.
 */
int kb2_knn_fused(const float *x_hi, const float *x_lo, const float *x_key, int64_t nx,
                  const float *y_hi, const float *y_lo, const float *y_key, int64_t ny,
                  int dpad, int cap, int splits, const float *tau_col, uint32_t *col_cnt,
                  uint64_t *col_buf, int col_cap, int64_t row_id_base, int32_t *cand_idx,
                  void *stream);
/* Column side of the dual-direction pass = candidates of kiez's reverse kNN
 * (hubness_reduction/base.py:37-42 -> neighbor_algorithm_base.py:116-136).
 * Per column: the cap emitted rows with the smallest column keys -> cand_idx [ny][cap]
 * (-1 padded), overflow[col] = 1 if more than col_cap rows were emitted.  Optional (both or
 * neither): tau_col = the thresholds the pass ran with, col_tau [ny] out = a lower bound of
 * the column key of every row that was NOT kept (input of kb2_refine_topk_checked). */
int kb2_col_select(const uint64_t *col_buf, const uint32_t *col_cnt, int64_t ny, int col_cap,
                   int cap, int32_t *cand_idx, int32_t *overflow, const float *tau_col,
                   float *col_tau, void *stream);
/* Between two row segments of a dual-direction pass (same stage of the reverse kNN,
 * hubness_reduction/base.py:37-42): per column with >= cap emitted rows, the
 * best cap entries move to the head of its buffer, col_cnt = cap and tau_col = the cap-th best
 * key so far (thresholds only tighten, so the bound kb2_col_select reports stays valid).  A
 * column with more than col_cap emitted rows keeps a sticky overflow count. */
int kb2_col_compact(uint64_t *col_buf, uint32_t *col_cnt, int64_t ny, int col_cap, int cap,
                    float *tau_col, void *stream);
/* Multi-GPU form of the same stage (reverse kNN of hubness_reduction/base.py:37-42 with the
 * SOURCE rows sharded over the ranks): every rank runs the dual-direction
 * pass over its own rows, so the per-column state has to agree across ranks.
 * kb2_col_heads: the first min(count, cap) entries of every column buffer (after
 *   kb2_col_compact: the best cap so far), EMPTY-padded -> out_entries [ny][cap] packed
 *   (key bits << 32 | row + row_offset) and / or out_keys [ny][cap] fp32 (+inf padded); either
 *   may be NULL.  The keys are all-gathered between row segments, the entries exchanged with an
 *   all-to-all after the last one (each rank finishes a shard of the columns).
 * kb2_kth_key: tau[col] = min(tau[col], kth smallest of keys[p * part_stride + col * L + j],
 *   p < nparts, j < L): the threshold every rank continues with (nparts * L <= 1024). */
int kb2_col_heads(const uint64_t *col_buf, const uint32_t *col_cnt, int64_t ny, int col_cap,
                  int cap, int64_t row_offset, uint64_t *out_entries, float *out_keys,
                  void *stream);
int kb2_kth_key(const float *keys, int nparts, int64_t part_stride, int64_t ny, int L, int kth,
                float *tau, void *stream);

/*
 * Screening candidate search (knn_screen.cu): the same stage of the reference as
 * kb2_knn_candidates / kb2_knn_fused (the ArgKmin contraction behind
 * sklearn_nearest_neighbors.py:96-101, both directions of hubness_reduction/base.py:37-42,92-94)
 * and the same contract, but the keys come from ONE TF32 product (the hi halves only) -- 3x fewer
 * MMAs, and the 128 x dpad query tile stays resident in shared memory.  Its candidate lists
 * are proposals: kb2_refine_topk_checked proves per row that the exact top k lies inside the
 * list and flags the rows where the proof fails, which the caller searches again with
 * kb2_knn_candidates (3xTF32).  Results are therefore the same as with the 3xTF32 search.
 *   steps/chained from kb2_screen_plan: the index is cut into `steps` contiguous ranges;
 *     chained = 1: ranges are sized for L2, searched in order per query tile, the lists are
 *       carried from range to range -> cand_idx/cand_key [nq][cap]; needs
 *       chain_flag [ceil(nq/128)*4] int32 scratch (zeroed by the call)
 *     chained = 0: independent ranges -> cand_idx/cand_key [nq][steps*cap]
 *   cand_key is required (fp32 screen keys, ascending per list, +inf padded): the last key of
 *     a list is the tau of the proof.
 *   dual-direction form when tau_col != NULL: additionally appends to col_buf/col_cnt like
 *     kb2_knn_fused (q_key = the row-side selection terms, row_id_base = first row of the
 *     segment).
 * kb2_screen_stages: ring slots (16 KB each) the kernel gets for (dpad, cap); 0 = shape not
 *   supported (dpad > 1024, cap > 128).  kb2_screen_config additionally reports the append-buffer
 *   slots per list and how many of the dpad/32 K chunks of the query tile stay resident in
 *   shared memory (the rest is streamed through the ring with the index tiles); max_smem <= 0 =
 *   the current device's opt-in limit (pass 232448 to evaluate the B200 plan without a device);
 *   ny = index rows one list sweeps (0 = long index): short indexes with long lists get longer
 *   append buffers (fewer list merges) at the price of resident query chunks.
 */
int kb2_screen_stages(int dpad, int cap, int dual);
int kb2_screen_config(int dpad, int cap, int dual, int max_smem, int64_t ny, int *slots,
                      int *resident);
int kb2_screen_plan(int64_t nq, int64_t ny, int dpad, int cap, int sm_count, int *steps,
                    int *chained);
int kb2_knn_screen(const float *q_hi, const float *q_key, int64_t nq, const float *y_hi,
                   const float *y_key, int64_t ny, int dpad, int cap, int steps, int chained,
                   int32_t *cand_idx, float *cand_key, int32_t *chain_flag, const float *tau_col,
                   uint32_t *col_cnt, uint64_t *col_buf, int col_cap, int64_t row_id_base,
                   void *stream);
/* max(0, max_i x[i]) -> *out (device): the largest selection term ||y - center||^2 of an
 * index, input of the completeness proof (part of the index build that replaces
 * sklearn_nearest_neighbors.py:83-94). */
int kb2_max_f32(const float *x, int64_t n, float *out, void *stream);
/* Rounding-error terms of the TF32 split, the other input of the completeness proof (same stage:
 * the index build that replaces sklearn_nearest_neighbors.py:83-94).  With w the centred row,
 * hi = rn_tf32(w) and lo = rn_tf32(w - hi) as written by kb2_prepare_rows:
 *   err_term [n] fp32 out: an upper bound of ||w - hi||^2 (= ||lo||^2 (1 + 2^-9), rounded up)
 *   err_max  device scalar out: its maximum over the rows. */
int kb2_split_error_terms(const float *lo, int64_t n, int dpad, float *err_term, float *err_max,
                          void *stream);

/*
 * Exact finish -- recomputes the distance of every candidate in float64 from
 * the raw fp32 rows (the oracle upcasts fp32 to fp64 too,
 * sklearn _middle_term_computer.pyx.tp:309-318), sorts (distance, id)
 * ascending and writes the best k: the (dist, ind) pair NNAlgorithm._kneighbors
 * must return (neighbor_algorithm_base.py:112-136).
 *   q, y     raw rows, elem_size 4 (fp32) or 8 (fp64: float64 callers such as the
 *            README example keep their exact values for the finish)
 *   cand_idx [nq][ncand] local ids (-1 skipped); index_base is added on output
 *   q_sqnorm,y_sqnorm: fp64 ||.||^2 of the raw rows, cosine only (else NULL)
 *   exclude_self: drop candidate j where j + self_offset == query row -- sklearn's
 *            kneighbors(X=None) semantics (neighbors/_base.py:937-958).  The search keeps
 *            the query's own row like any other candidate (it costs one of the list's
 *            margin slots); it is removed here, by id.
 */
int kb2_refine_topk(const void *q, int64_t nq, int64_t ldq, const void *y, int64_t ny,
                    int64_t ldy, int d, int elem_size, const double *q_sqnorm,
                    const double *y_sqnorm, const int32_t *cand_idx, int ncand, int metric,
                    int64_t index_base, int exclude_self, int64_t self_offset, int k,
                    double *out_dist, int64_t *out_ind, void *stream);

/*
 * Exact finish + completeness proof for lists proposed by kb2_knn_screen: the (dist, ind)
 * NNAlgorithm._kneighbors must return (neighbor_algorithm_base.py:112-136).  Same outputs as
 * kb2_refine_topk; additionally unverified[row] = 0 when the float64 k-th best distance
 * proves that no index row outside the list can be among the k nearest, else 1:
 *   every non-candidate has screen key >= tau_row = min_j tau[row*tau_row_stride + j*tau_step],
 *   j < tau_count (+inf = the list never filled: nothing was left out);
 *   |screen key - exact key| <= E = 2 (dq ym + qn (1 + 2^-11) dym + eps_acc qn ym) + 2^-21 (ym^2 + 2 qn ym)
 *   with qn = ||q-c||, ym = max ||y-c||, dq = ||q-c - hi(q-c)||, dym = max ||y-c - hi(y-c)||
 *   (the exact TF32 rounding errors of the operands, Cauchy-Schwarz per term);  verified iff
 *   exact key of the k-th best < tau_row - E.
 *   q_key [nq] / y_key_max (device scalar): selection terms of the queries / their maximum
 *   over the index (kb2_max_f32); ignored for cosine (unit rows).
 *   q_err [nq] / y_err_max (device scalar): kb2_split_error_terms of the queries / the index.
 *   eps_acc: bound of the fp32 accumulation error relative to qn ym (dpad 2^-22 + 2^-21).
 */
int kb2_refine_topk_checked(const void *q, int64_t nq, int64_t ldq, const void *y, int64_t ny,
                            int64_t ldy, int d, int elem_size, const double *q_sqnorm,
                            const double *y_sqnorm, const int32_t *cand_idx, int ncand,
                            int metric, int64_t index_base, int exclude_self,
                            int64_t self_offset, int k, double *out_dist, int64_t *out_ind,
                            const float *tau, int64_t tau_row_stride, int tau_step,
                            int tau_count, const float *q_key, const float *y_key_max,
                            const float *q_err, const float *y_err_max, double eps_acc,
                            int32_t *unverified, void *stream);

/*
 * Row-wise top-k of (dist, ind) pairs, ascending, ties by input position, NaN last.
 * Replaces HubnessReduction._sort (kiez/hubness_reduction/base.py:72-87) and is the
 * multi-GPU merge kernel: with nparts > 1 row r is the concatenation of
 * dist[p*part_stride + r*c .. +c) for p in [0, nparts).
 */
int kb2_topk_rows(const double *dist, const int64_t *ind, int64_t n, int c, int nparts,
                  int64_t part_stride, int k, double *out_dist, int64_t *out_ind,
                  void *stream);

/* Per-row statistics of the reverse neighbourhoods (CSLS/NICDM mean, LS c-th
 * neighbour, MutualProximity mu/sigma with ddof=0: csls.py:90,
 * local_scaling.py:136,143, mutual_proximity.py:101-103).  Any output may be NULL. */
int kb2_row_stats(const double *dist, int64_t n, int c, double *mean, double *sd,
                  double *last, void *stream);

/*
 * Rescale + final sort, one pass (transform of csls.py:85-96, local_scaling.py:129-151,
 * mutual_proximity.py:166-183 numpy branch, fused with HubnessReduction._sort base.py:72-87):
 * out = top-k over the c candidates of
 *   CSLS     2 d - mean_c(d_row) - stat_a[ind]
 *   LS       1 - exp(-d^2 / (d_row[c-1] * stat_a[ind]))
 *   NICDM    d / sqrt(mean_c(d_row) * stat_a[ind])
 *   MP_GAUSS 1 - sf(d; mu_row, sd_row) * sf(d; stat_a[ind], stat_b[ind])
 * k == 0: write the unsorted (n, c) transform (HubnessReduction.transform contract).
 */
int kb2_rescale_topk(int mode, const double *dist, const int64_t *ind, int64_t n, int c,
                     const double *stat_a, const double *stat_b, int64_t n_stats, int k,
                     double *out_dist, int64_t *out_ind, void *stream);

/* MutualProximity(method="empiric") transform, mutual_proximity.py:185-212,
 * including its source-id / target-id index-space mix, as c x c membership
 * tests instead of an O(max_ind) table per pair. */
int kb2_mp_empiric_topk(const double *dist, const int64_t *ind, int64_t n, int c,
                        const double *rev_dist, const int64_t *rev_ind, int64_t m,
                        int c_rev, int k, double *out_dist, int64_t *out_ind, void *stream);

/* DisSimLocal._fit (dis_sim.py:95-108): centroid of the c_rev reverse neighbours
 * (source rows) of every target row and ||target - centroid||^2, in fp64.
 *   centroids [m][d] fp64 out or NULL, dist_to_cent [m] fp64 out */
int kb2_dsl_fit(const void *source, int64_t n_source, int64_t lds, const void *target,
                int64_t m, int64_t ldt, int d, int elem_size, const int64_t *rev_ind, int c_rev,
                double *centroids, double *dist_to_cent, void *stream);

/* DisSimLocal.transform (dis_sim.py:139-181), stage 1: raw (n, c) values
 * ||q-t||^2 - ||q-c_q||^2 - dist_to_cent[ind] and their global minimum
 * (*global_min must be initialised to +inf by the caller; with row-sharded
 * multi-GPU runs all-reduce(min) it before stage 2). */
int kb2_dsl_transform(const void *query, int64_t n, int64_t ldq, const void *target,
                      int64_t m, int64_t ldt, int d, int elem_size, const int64_t *ind, int c,
                      const double *dist_to_cent, double *raw, double *global_min,
                      void *stream);
/* stage 2 (dis_sim.py:171-177 + HubnessReduction._sort base.py:72-87): shift by -min if
 * negative, sqrt unless squared, top-k (k==0: unsorted). */
int kb2_dsl_finish_topk(const double *raw, const int64_t *ind, int64_t n, int c,
                        const double *global_min, int squared, int k, double *out_dist,
                        int64_t *out_ind, void *stream);

/*
 * kiez.analysis.hubness_score (kiez/analysis/estimation.py:272-351) on device.
 * kb2_index_range: min and max id over the first k columns -> out[2] (device).
 * kb2_k_occurrence: bincount of the first k columns, negatives dropped (:286-295).
 * kb2_hub_moments: one pass over the histogram; out[0..9] (device, fp64) =
 *   sum, sum (x-mean)^2, sum (x-mean)^3, sum |x-mean|, sum sqrt(x), max,
 *   #zeros (antihubs), #(x >= hub_thresh) (hubs), sum x over hubs, sum x^2
 * kb2_compact_ids: ascending ids whose count is ==0 (mode 0) or >= thresh (mode 1);
 *   scratch must hold (nbins+1023)/1024 + 1 int64; *out_count receives the total.
 * kb2_gini_numerator: sum_ij |x_i - x_j| exactly (int64) via a value histogram;
 *   scratch must hold max_value+1 int64.
 */
int kb2_index_range(const int64_t *ind, int64_t n, int64_t ld, int k, int64_t *out,
                    void *stream);
int kb2_k_occurrence(const int64_t *ind, int64_t n, int64_t ld, int k, int64_t nbins,
                     int64_t *hist, void *stream);
int kb2_hub_moments(const int64_t *hist, int64_t nbins, double mean, double hub_thresh,
                    double *out, void *stream);
int kb2_compact_ids(const int64_t *hist, int64_t nbins, int mode, double thresh,
                    int64_t *scratch, int64_t *out_ids, int64_t *out_count, void *stream);
int kb2_gini_numerator(const int64_t *hist, int64_t nbins, int64_t max_value,
                       int64_t *scratch, int64_t *out, void *stream);

/* kiez.evaluate.hits (kiez/evaluate/eval_metrics.py:23-61): for each of nks
 * cut-offs ks[i], the number of rows r with gold[r] among ind[r][0..ks[i]).
 * gold[r] < 0 = row not evaluated.  counts [nks] int64 (device) is accumulated into. */
int kb2_hits(const int64_t *ind, int64_t n, int64_t ld, int k, const int64_t *gold,
             const int32_t *ks, int nks, int64_t *counts, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* KIEZ_B200_H */
