"""The rescale stage against the reference's OWN GPU route.

With a tensor-returning backend the reference's hubness classes run as unfused torch-eager ops
on the device (`_use_torch`, kiez/hubness_reduction/base.py:43-44; exercised by the reference's
tests/neighbors/test_faiss.py:66-116).  That -- not a hand-written kernel -- is the existing GPU
implementation of stage (2) of the north star, and the thing kb2_row_stats / kb2_rescale_topk /
kb2_dsl_* must beat (it substitutes for the Faiss row f3: faiss has no
sm_100 build in this image).  For each method this script times, on the SAME CUDA tensors:

  reference : <Class>._fit(rev_dist, rev_ind, source, target) + .transform(fwd_dist, fwd_ind,
              source) + HubnessReduction._sort(.., k)   -- the unmodified reference from
              baseline/_ref under oracle/ref_shim.py, torch branch, float64 (what the B200
              backend hands it) and float32 (what Faiss would)
  kiez_b200 : the same three calls of this package's classes (fused CUDA kernels)

and reports milliseconds (CUDA events, L2 flushed between runs), device launches (torch.profiler
activity count), peak extra device memory, and for the fused kernels the algorithmic bytes and
GB/s against the measured HBM peak.  Results of both are compared where the reference's torch
branch computes the same function (CSLS, LS, NICDM, DSL; its MP-Gaussian branch uses ddof=1 and
fp32 1-cdf -- a different function).

    python tools/bench_rescale_vs_reference.py [--n 1000000 --m 1000000 --c 10 --k 10 --d 256]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch.profiler import ProfilerActivity, profile


class _StubAlgo:
    """What the hubness classes read from a backend (hubness_reduction/base.py:24, dis_sim.py:47-49)."""

    def __init__(self, c, source, target):
        self.n_candidates, self.metric, self.p = c, "euclidean", 2
        self.source_, self.target_ = source, target
        self.device = source.device


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--c", type=int, default=10)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    n, m, d, c, k = args.n, args.m, args.d, args.c, args.k
    dev = torch.device("cuda", 0)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]

    from oracle import ref_shim

    ref = None
    if ref_shim.reference_available():
        ref_shim.load_reference()
        import kiez.hubness_reduction as ref

    import kiez_b200.hubness_reduction as mine

    g = torch.Generator(device=dev)
    g.manual_seed(0)
    source = torch.randn((n, d), device=dev, generator=g)
    target = torch.randn((m, d), device=dev, generator=g)
    # synthetic but well-formed kNN results: sorted positive distances, valid ids
    fwd_d = (torch.rand((n, c), dtype=torch.float64, device=dev, generator=g) + 20.0).sort(dim=1).values
    fwd_i = torch.randint(0, m, (n, c), device=dev, generator=g)
    rev_d = (torch.rand((m, c), dtype=torch.float64, device=dev, generator=g) + 20.0).sort(dim=1).values
    rev_i = torch.randint(0, n, (m, c), device=dev, generator=g)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed(fn):
        """Median device-side milliseconds of fn() (CUDA events on the current stream, L2 flushed
        before every run), its device launches and its peak extra device memory.  The events
        bracket the host code of fn too, so a pipeline whose host side is slower than its
        kernels is reported with its host-bound time -- what a caller sees."""
        for _ in range(3):                            # warm-up (allocator, kernel load)
            out = fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            flush.fill_(1)                            # L2 flush (512 MB > 126 MB L2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        torch.cuda.reset_peak_memory_stats(dev)
        base = torch.cuda.memory_allocated(dev)
        out = fn()
        torch.cuda.synchronize()
        extra = torch.cuda.max_memory_allocated(dev) - base
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        kernel_ms = 1e-3 * sum(e.time_range.end - e.time_range.start for e in evs)
        return sorted(times)[len(times) // 2], len(evs), extra, out, kernel_ms

    def pipeline(cls, kwargs, fd, fi, rd, ri, src, tgt, use_torch):
        def run():
            hub = cls(nn_algo=_StubAlgo(c, src, tgt), **kwargs)
            hub._use_torch = use_torch
            hub._fit(rd, ri, src, tgt)
            if hasattr(hub, "_transform_topk"):       # kiez_b200: fused rescale + top-k
                return hub._transform_topk(fd, fi, src, k)
            out_d, out_i = hub.transform(fd, fi, src)
            return type(hub)._sort(out_d, out_i, k)
        return run

    # algorithmic bytes of the fused path per method
    stats_b = m * c * 8 + m * 8
    rescale_b = lambda gathers: n * c * 16 + n * c * 8 * gathers + n * k * 16
    dsl_b = (m * c * 8 + m * c * d * 4 + m * d * 4 + m * 8) + \
            (n * c * 8 + n * c * d * 4 + n * d * 4 + 2 * n * c * 8) + (n * c * 16 + n * k * 16)
    methods = [
        ("CSLS", "CSLS", {}, stats_b + rescale_b(1), True),
        ("LocalScaling[standard]", "LocalScaling", {"method": "standard"}, stats_b + rescale_b(1), True),
        ("LocalScaling[nicdm]", "LocalScaling", {"method": "nicdm"}, stats_b + rescale_b(1), True),
        ("MutualProximity[normal]", "MutualProximity", {"method": "normal"},
         stats_b + m * 8 + rescale_b(2), False),
        ("DisSimLocal", "DisSimLocal", {}, dsl_b, True),
    ]
    results = []
    for label, cls_name, kw, bytes_alg, same_function in methods:
        rec = {"method": label, "shape": {"n": n, "m": m, "c": c, "k": k, "d": d}}
        ms, launches, extra, out, k_ms = timed(pipeline(getattr(mine, cls_name), kw, fwd_d, fwd_i,
                                                        rev_d, rev_i, source, target, True))
        rec["kiez_b200"] = {"ms": ms, "kernel_ms": k_ms, "launches": launches,
                            "extra_device_bytes": int(extra), "algorithmic_bytes": int(bytes_alg),
                            "achieved_gbs": bytes_alg / (k_ms * 1e-3) / 1e9,
                            "frac_of_hbm_peak": bytes_alg / (k_ms * 1e-3) / 1e9 / peak,
                            "note": "ms = CUDA events around the three calls (host code included); "
                                    "kernel_ms = sum of the device activities; GB/s from kernel_ms"}
        if ref is not None:
            for tag, cast in (("f64", torch.float64), ("f32", torch.float32)):
                try:
                    r_ms, r_l, r_x, r_out, r_kms = timed(pipeline(
                        getattr(ref, cls_name), kw, fwd_d.to(cast), fwd_i, rev_d.to(cast), rev_i,
                        source.to(cast) if cls_name == "DisSimLocal" else source,
                        target.to(cast) if cls_name == "DisSimLocal" else target, True))
                    entry = {"ms": r_ms, "kernel_ms": r_kms, "launches": r_l,
                             "extra_device_bytes": int(r_x), "speedup_of_kiez_b200": r_ms / ms,
                             "kernel_speedup_of_kiez_b200": r_kms / k_ms}
                    if same_function and tag == "f64":
                        entry["max_abs_diff_vs_kiez_b200"] = float(
                            (r_out[0].double() - out[0]).abs().max())
                        entry["index_mismatch_rows"] = int((r_out[1] != out[1]).any(dim=1).sum())
                    rec[f"reference_torch_eager_{tag}"] = entry
                except Exception as exc:  # e.g. out of memory in the reference's temporaries
                    rec[f"reference_torch_eager_{tag}"] = {"failed": f"{type(exc).__name__}: {exc}"[:200]}
                torch.cuda.empty_cache()
        results.append(rec)
        print(json.dumps(rec), flush=True)
    out_path = args.out or os.path.join(ROOT, "gpurun_out", f"rescale_vs_reference_n{n}_c{c}.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump({"hbm_peak_gbs": peak, "results": results}, fh, indent=1)


if __name__ == "__main__":
    main()
