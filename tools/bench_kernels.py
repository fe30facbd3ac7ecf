"""HBM-roofline measurements of the memory-bound kernels (prepare, exact finish, rescale,
DisSimLocal, analysis): CUDA-event time per launch with an L2 flush between launches,
algorithmic bytes per launch, achieved GB/s and the fraction of the
measured HBM peak (MEASURED_PEAKS.json).  Writes one JSON object per kernel.

    python tools/bench_kernels.py [--n 1000000] [--m 1000000] [--d 256] [--c 10] [--k 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kiez_b200 import _lib as lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--c", type=int, default=10)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    n, m, d, c, k = args.n, args.m, args.d, args.c, args.k
    dev = torch.device("cuda", 0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peak = 6650.0
    src = "fallback"
    pk = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
        src = "MEASURED_PEAKS.json"
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    st = lib.stream_ptr()
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    x = torch.randn((n, d), device=dev, generator=g)
    y = torch.randn((m, d), device=dev, generator=g)
    dpad = lib.lib.kb2_padded_dim(d)
    hi = torch.empty((n, dpad), device=dev)
    lo = torch.empty((n, dpad), device=dev)
    key = torch.empty(n, device=dev)
    cap = 16 if c <= 10 else min(128, ((c + max(6, c // 8)) + 7) // 8 * 8)
    cand = torch.randint(0, m, (n, cap), device=dev, dtype=torch.int32, generator=g)
    fd = torch.empty((n, c), dtype=torch.float64, device=dev)
    fi = torch.empty((n, c), dtype=torch.int64, device=dev)
    rd = torch.rand((m, c), dtype=torch.float64, device=dev, generator=g).sort(dim=1).values
    ri = torch.randint(0, n, (m, c), device=dev, generator=g)
    mean = torch.empty(m, dtype=torch.float64, device=dev)
    sd = torch.empty(m, dtype=torch.float64, device=dev)
    od = torch.empty((n, k), dtype=torch.float64, device=dev)
    oi = torch.empty((n, k), dtype=torch.int64, device=dev)
    raw = torch.empty((n, c), dtype=torch.float64, device=dev)
    gmin = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
    d2c = torch.empty(m, dtype=torch.float64, device=dev)
    hist = torch.empty(max(n, m), dtype=torch.int64, device=dev)
    mom = torch.empty(10, dtype=torch.float64, device=dev)

    results = []

    def run(name, bytes_alg, fn, note=""):
        fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            flush.fill_(1)                           # L2 flush (512 MB > 126 MB L2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        gbs = bytes_alg / (ms * 1e-3) / 1e9
        rec = {"kernel": name, "ms": ms, "algorithmic_bytes": bytes_alg, "achieved_gbs": gbs,
               "peak_gbs": peak, "peak_source": src, "frac": gbs / peak, "note": note,
               "shape": {"n": n, "m": m, "d": d, "c": c, "k": k}}
        results.append(rec)
        print(json.dumps(rec), flush=True)

    run("prepare_rows", n * d * 4 + 2 * n * dpad * 4 + n * 4,
        lambda: lib.call("kb2_prepare_rows", lib.ptr(x), n, d, d, None, 0, lib.ptr(hi), lib.ptr(lo),
                         dpad, lib.ptr(key), None, st),
        "read n*d*4, write 2*n*dpad*4 + n*4")
    run("refine_topk", n * d * 4 + n * cap * (d * 4 + 4) + n * c * 16,
        lambda: lib.call("kb2_refine_topk", lib.ptr(x), n, d, lib.ptr(y), m, d, d, 4, None, None,
                         lib.ptr(cand), cap, 0, 0, 0, 0, c, lib.ptr(fd), lib.ptr(fi), st),
        "gather cap rows of d*4 B per query (random ids: worst case, no L2 reuse) + write n*c*16")
    run("row_stats(mean,sd)", m * c * 8 + m * 16,
        lambda: lib.call("kb2_row_stats", lib.ptr(rd), m, c, lib.ptr(mean), lib.ptr(sd), None, st))
    for mode, nm, g_ in ((0, "csls", 1), (1, "ls", 1), (2, "nicdm", 1), (3, "mp_gauss", 2)):
        run(f"rescale_topk[{nm}]", n * c * 16 + n * c * 8 * g_ + n * k * 16,
            lambda mode=mode: lib.call("kb2_rescale_topk", mode, lib.ptr(fd), lib.ptr(fi), n, c,
                                       lib.ptr(mean), lib.ptr(sd), m, k, lib.ptr(od), lib.ptr(oi), st),
            "read n*c*(8+8), gather n*c*8*g (32 B sectors actually move), write n*k*16")
    run("topk_rows", n * c * 16 + n * k * 16,
        lambda: lib.call("kb2_topk_rows", lib.ptr(fd), lib.ptr(fi), n, c, 1, 0, k, lib.ptr(od),
                         lib.ptr(oi), st))
    if c <= 64:
        run("mp_empiric_topk", n * c * 16 + n * c * c * 16 + n * k * 16,
            lambda: lib.call("kb2_mp_empiric_topk", lib.ptr(fd), lib.ptr(fi), n, c, lib.ptr(rd),
                             lib.ptr(ri), m, c, k, lib.ptr(od), lib.ptr(oi), st),
            "per candidate: its reverse row (c*16 B) is gathered")
    run("dsl_fit", m * c * 8 + m * c * d * 4 + m * d * 4 + m * 8,
        lambda: lib.call("kb2_dsl_fit", lib.ptr(x), n, d, lib.ptr(y), m, d, d, 4, lib.ptr(ri), c,
                         None, lib.ptr(d2c), st),
        "gather c source rows per target row")
    run("dsl_transform", n * c * 8 + n * c * d * 4 + n * d * 4 + n * c * 8 + n * c * 8,
        lambda: lib.call("kb2_dsl_transform", lib.ptr(x), n, d, lib.ptr(y), m, d, d, 4, lib.ptr(fi),
                         c, lib.ptr(d2c), lib.ptr(raw), lib.ptr(gmin), st),
        "gather c target rows per query")
    run("dsl_finish_topk", n * c * 16 + n * k * 16,
        lambda: lib.call("kb2_dsl_finish_topk", lib.ptr(raw), lib.ptr(fi), n, c, lib.ptr(gmin), 0, k,
                         lib.ptr(od), lib.ptr(oi), st))
    run("k_occurrence", n * k * 8 + m * 8 * 2,
        lambda: lib.call("kb2_k_occurrence", lib.ptr(oi.clamp_(0, m - 1)), n, k, k, m, lib.ptr(hist), st),
        "read n*k ids, memset + atomics on m bins")
    run("hub_moments", m * 8,
        lambda: lib.call("kb2_hub_moments", lib.ptr(hist), m, float(k), 2.0 * k, lib.ptr(mom), st))
    out = os.path.join(root, "gpurun_out", f"kernels_n{n}_c{c}_d{d}.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        json.dump(results, fh, indent=1)


if __name__ == "__main__":
    main()
