"""Diagnostic: run both candidate-search kernels on one problem and compare their raw
candidate lists (keys + ids) with a float64 numpy computation of the same selection key."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from kiez_b200 import B200, _lib as lib


def run(nq, ny, d, cap, splits, impl, seed=0, exclude_self=False):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    y = rng.standard_normal((ny, d)).astype(np.float32)
    algo = B200(n_candidates=cap, impl=impl, center=False)
    qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
    cand = torch.full((nq, splits * cap), -7, dtype=torch.int32, device="cuda")
    key = torch.full((nq, splits * cap), float("nan"), dtype=torch.float32, device="cuda")
    code = {"tc": lib.KNN_TC, "tc1": lib.KNN_TC1, "simt": lib.KNN_SIMT}[impl]
    lib.call("kb2_knn_candidates", code, lib.ptr(qp.hi), lib.ptr(qp.lo), nq, lib.ptr(yp.hi),
             lib.ptr(yp.lo), lib.ptr(yp.key), ny, qp.dpad, cap, splits,
             lib.ptr(cand), lib.ptr(key), lib.stream_ptr())
    torch.cuda.synchronize()
    # float64 reference of the selection key
    K = (y.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * q.astype(np.float64) @ y.astype(np.float64).T
    return cand.cpu().numpy(), key.cpu().numpy(), K


def check(nq, ny, d, cap, splits, impl):
    cand, key, K = run(nq, ny, d, cap, splits, impl)
    bad_rows = 0
    max_err = 0.0
    per = None
    for r in range(nq):
        ids = cand[r]
        valid = ids >= 0
        want = np.sort(K[r])[: min(cap, ny)]
        got = np.sort(key[r][valid])[: min(cap, ny)]
        if splits == 1:
            if got.shape != want.shape or not np.allclose(got, want, rtol=1e-4, atol=1e-3):
                bad_rows += 1
                if bad_rows <= 3:
                    print(f"  row {r}: got {got[:6]} want {want[:6]} ids {ids[:6]}")
            else:
                max_err = max(max_err, float(np.abs(got - want).max()))
        # keys must equal K at the returned ids
        kk = K[r][ids[valid]]
        if not np.allclose(key[r][valid], kk, rtol=1e-4, atol=1e-3):
            bad_rows += 1
            if bad_rows <= 3:
                print(f"  row {r}: key/id mismatch key {key[r][valid][:6]} K[id] {kk[:6]} ids {ids[valid][:6]}")
    print(f"{impl} nq={nq} ny={ny} d={d} cap={cap} splits={splits}: bad_rows={bad_rows} max_err={max_err:.3g}")
    return bad_rows == 0


if __name__ == "__main__":
    impls = sys.argv[1:] or ["simt", "tc"]
    ok = True
    for impl in impls:
        for (nq, ny, d, cap, splits) in [(128, 256, 32, 16, 1), (128, 512, 64, 16, 1),
                                         (100, 300, 40, 16, 1), (300, 1000, 256, 32, 1),
                                         (256, 4096, 128, 64, 2), (130, 700, 96, 112, 1),
                                         (700, 3000, 256, 16, 1), (1100, 5000, 128, 56, 3)]:
            ok &= check(nq, ny, d, cap, splits, impl)
    print("DIAG", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
