// C-ABI glue: error reporting, capability queries and the candidate-search dispatch.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace kb2 {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int launch_knn_simt(const float *, const float *, int64_t, const float *, const float *,
                    const float *, int64_t, int, int, int, int32_t *, float *, cudaStream_t);
int launch_knn_tc(const float *, const float *, int64_t, const float *, const float *, const float *,
                  int64_t, int, int, int, int32_t *, float *, cudaStream_t);
int launch_knn_tc2(const float *, const float *, int64_t, const float *, const float *, const float *,
                   int64_t, int, int, int, int32_t *, float *, cudaStream_t);

}  // namespace kb2

extern "C" int kb2_version(void) { return KB2_VERSION; }
extern "C" const char *kb2_last_error(void) { return kb2::g_error; }
extern "C" int kb2_max_candidates(void) { return 128; }

extern "C" int kb2_suggest_splits(int64_t nq, int64_t ny, int cap, int sm_count) {
    // Work units are (128-row query tile, index split).  Split the index only when the
    // query tiles alone cannot fill ~2 waves of SMs; keep >= 4 index tiles of 256 per split.
    if (sm_count <= 0) sm_count = 148;
    const int64_t q_tiles = (nq + 127) / 128;
    if (q_tiles >= 2 * (int64_t)sm_count) return 1;
    int64_t s = (2 * (int64_t)sm_count + q_tiles - 1) / q_tiles;
    const int64_t max_by_rows = ny / 1024 > 0 ? ny / 1024 : 1;
    if (s > max_by_rows) s = max_by_rows;
    const int64_t max_by_cand = 2048 / (cap > 0 ? cap : 1);
    if (s > max_by_cand) s = max_by_cand;
    return (int)(s < 1 ? 1 : s);
}

extern "C" int kb2_knn_candidates(int impl, const float *q_hi, const float *q_lo, int64_t nq,
                                  const float *y_hi, const float *y_lo, const float *y_key,
                                  int64_t ny, int dpad, int cap, int splits, int32_t *cand_idx,
                                  float *cand_key, void *stream) {
    KB2_CHECK(nq >= 0 && ny > 0, "knn_candidates: bad shape nq=%lld ny=%lld", (long long)nq,
              (long long)ny);
    KB2_CHECK(nq < (1LL << 31) - 256 && ny < (1LL << 31) - 256,
              "knn_candidates: more than 2^31 rows per call; shard the call");
    KB2_CHECK(dpad > 0 && dpad % 32 == 0, "knn_candidates: dpad=%d must be a multiple of 32", dpad);
    KB2_CHECK(cap > 0 && cap <= kb2_max_candidates(), "knn_candidates: cap=%d outside (0, %d]", cap,
              kb2_max_candidates());
    KB2_CHECK(splits >= 1 && (int64_t)splits * cap <= 2048,
              "knn_candidates: splits*cap=%lld exceeds 2048", (long long)splits * cap);
    KB2_CHECK(impl >= 0 && impl <= 3, "knn_candidates: unknown impl %d", impl);
    if (nq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (impl == KB2_KNN_SIMT)
        return kb2::launch_knn_simt(q_hi, q_lo, nq, y_hi, y_lo, y_key, ny, dpad, cap, splits,
                                    cand_idx, cand_key, st);
    // tensor-core search: CTA-pair kernel unless the single-CTA one is asked for
    // (impl KB2_KNN_TC1, or KB2_TC_MODE=1 in the environment for A/B runs)
    const char *mode = getenv("KB2_TC_MODE");
    const bool single = impl == KB2_KNN_TC1 || (mode && mode[0] == '1');
    if (!single) {
        const int rc = kb2::launch_knn_tc2(q_hi, q_lo, nq, y_hi, y_lo, y_key, ny, dpad, cap, splits,
                                           cand_idx, cand_key, st);
        if (rc >= 0) return rc;
    }
    return kb2::launch_knn_tc(q_hi, q_lo, nq, y_hi, y_lo, y_key, ny, dpad, cap, splits, cand_idx,
                              cand_key, st);
}
