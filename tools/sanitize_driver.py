"""Small-shape pass over EVERY kernel family, for `compute-sanitizer` (tools/sanitize.sh):
the 1xTF32 screen (one-direction, dual-direction, chained ranges), the 3xTF32 pair / single-CTA /
dual-direction kernels, the SIMT cross-check, exact finish + proof, column select / compact, every
rescaler, top-k / merge, analysis and hits.  Results are checked against the oracle so that a
sanitizer-clean run is also a correct one.  Shapes are tiny: the tools slow kernels 10-100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from kiez_b200 import B200, Kiez, hits, hubness_score
from oracle import kiez_oracle as O


def _np(x):
    return x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def main():
    families = sys.argv[1:] or ["screen", "tf32x3", "simt", "rescale", "analysis"]
    rng = np.random.default_rng(0)
    n, m, d, c, k = 700, 900, 96, 10, 5
    source = rng.standard_normal((n, d)).astype(np.float32)
    target = rng.standard_normal((m, d)).astype(np.float32)
    s64, t64 = source.astype(np.float64), target.astype(np.float64)
    want = {h: O.kiez_kneighbors(s64, t64, hubness=h, n_candidates=c, k=k)
            for h in ("csls", "ls", "nicdm", "mp_gaussian", "mp_empiric", "dsl", "no")}

    def run(label, hub_key, hub, kw, **algo_kw):
        algo = B200(n_candidates=c, **algo_kw)
        algo.FUSED_SEGMENT_MIN_ROWS = 256
        inst = Kiez(n_candidates=c, algorithm=algo, hubness=hub, hubness_kwargs=dict(kw))
        inst.fit(source, target)
        got_d, got_i = inst.kneighbors(k)
        O.assert_neighbors_match(_np(got_d), _np(got_i), *want[hub_key], rtol=1e-5, atol=5e-6, what=label,
                                 max_bad_rows=0.03 if hub_key == "mp_empiric" else 0.0)
        torch.cuda.synchronize()
        print("ok", label, flush=True)
        return got_i

    ind = None
    if "screen" in families:
        ind = run("screen dual", "csls", "CSLS", {}, impl="tc", precision="screen", fused=True)
        run("screen one-direction", "csls", "CSLS", {}, impl="tc", precision="screen", fused=False)
        os.environ["KB2_SCREEN_RANGE_MB"] = "0.25"
        q = rng.standard_normal((40000, 64)).astype(np.float32)
        y = rng.standard_normal((3000, 64)).astype(np.float32)
        algo = B200(n_candidates=10, precision="screen", fused=False)
        algo.fit(q, y)
        dist, idx = algo.kneighbors(k=10)
        rows = rng.choice(40000, 200, replace=False)
        wd, wi = O.knn_brute(q[rows].astype(np.float64), y.astype(np.float64), 10, "euclidean")
        O.assert_neighbors_match(_np(dist)[rows], _np(idx)[rows], wd, wi, 1e-5, 5e-6, what="chained")
        del os.environ["KB2_SCREEN_RANGE_MB"]
        print("ok screen chained ranges", flush=True)
        # long lists at d = 256: 5 of 8 K chunks of the query tile resident, the rest streamed
        q = rng.standard_normal((600, 256)).astype(np.float32)
        y = rng.standard_normal((3000, 256)).astype(np.float32)
        for fused in (False, True):
            algo = B200(n_candidates=50, precision="screen", fused=fused)
            qp, yp = algo._prepare(q, cache=False), algo._prepare(y, cache=False)
            if fused:
                (dist, idx), _rev = algo.search_both(qp, yp, 50, 50)
            else:
                dist, idx = algo.search(qp, yp, 50)
            wd, wi = O.knn_brute(q.astype(np.float64), y.astype(np.float64), 50, "euclidean")
            O.assert_neighbors_match(_np(dist), _np(idx), wd, wi, 1e-5, 5e-6, what="partial residency")
        print("ok screen partially resident query tile", flush=True)
    if "tf32x3" in families:
        ind = run("3xTF32 pair dual", "csls", "CSLS", {}, impl="tc", precision="tf32x3", fused=True)
        run("3xTF32 pair", "nicdm", "LocalScaling", {"method": "nicdm"}, impl="tc",
            precision="tf32x3", fused=False)
        run("3xTF32 single CTA", "no", None, {}, impl="tc1", precision="tf32x3", fused=False)
    if "simt" in families:
        run("simt", "csls", "CSLS", {}, impl="simt", fused=False)
    if "rescale" in families:
        for key, hub, kw in [("ls", "LocalScaling", {"method": "standard"}),
                             ("mp_gaussian", "MutualProximity", {"method": "normal"}),
                             ("mp_empiric", "MutualProximity", {"method": "empiric"}),
                             ("dsl", "DisSimLocal", {})]:
            ind = run(f"rescale {key}", key, hub, kw, impl="simt", fused=False)
    if "analysis" in families:
        if ind is None:
            ind = want["csls"][1]
        ind = _np(ind)
        got = hubness_score(ind, m, k=k, return_value="all")
        ref = O.hubness_score(ind, m, k=k, return_value="all")
        assert abs(got["robinhood"] - ref["robinhood"]) < 1e-9 and abs(got["gini"] - ref["gini"]) < 1e-9
        gold = rng.integers(0, m, n)
        assert hits(np.asarray(ind), gold, k=[1, 5]) == O.hits(np.asarray(ind), gold, k=[1, 5])
        torch.cuda.synchronize()
        print("ok analysis + hits", flush=True)
    print("sanitize driver: all families ok")


if __name__ == "__main__":
    main()
