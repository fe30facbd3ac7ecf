#!/usr/bin/env python
"""Benchmark of the kiez hot path on B200: queries/s of exact kNN + hubness reduction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|c5|custom ...]
    python bench.py --impl reference ...      # the reference's CPU path, bounded sample

One *step* = one full ``Kiez(algorithm=B200, hubness=CSLS).fit(source, target)`` +
``kneighbors(k)`` (+ ``hubness_score``) over the whole synthetic workload, i.e. both kNN
passes, the rescale and the final sort for every query.  Default workload = BASELINE.json's
metric config (C4: 1M x 1M, d=256, CSLS, k=10).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs (--impl reference, and the
# parity oracle on rank 0) must be allowed all host threads, and every rank of the GPU arm its
# share of them (host-side copies of the upload), so undo that before numpy/sklearn/torch load.
if os.environ.get("OMP_NUM_THREADS") == "1":
    _cores = len(os.sched_getaffinity(0))
    _reference = "--impl" in sys.argv and "reference" in sys.argv
    _world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    if int(os.environ.get("RANK", "0")) == 0:
        os.environ["OMP_NUM_THREADS"] = str(_cores if _reference else max(1, _cores // 2))
    else:
        os.environ["OMP_NUM_THREADS"] = str(max(1, _cores // (2 * _world)))

import numpy as np

WORKLOADS = {
    # name: n (source rows), m (target rows), d, n_candidates, k, hubness
    "c1": dict(n=100, m=100, d=50, c=10, k=5, hubness="CSLS"),
    "c2": dict(n=15_000, m=15_000, d=256, c=50, k=10, hubness="CSLS"),
    "c3": dict(n=100_000, m=100_000, d=256, c=100, k=10, hubness="MutualProximity"),
    "c4": dict(n=1_000_000, m=1_000_000, d=256, c=10, k=10, hubness="CSLS"),
    "c5": dict(n=1_000_000, m=10_000_000, d=128, c=50, k=10, hubness="LocalScaling"),
    "c5dsl": dict(n=1_000_000, m=10_000_000, d=128, c=50, k=10, hubness="DisSimLocal"),
}
HUB_KWARGS = {"LocalScaling": {"method": "nicdm"}, "MutualProximity": {"method": "normal"}}
ORACLE_HUB = {"CSLS": "csls", "LocalScaling": "nicdm", "MutualProximity": "mp_gaussian",
              "DisSimLocal": "dsl", None: "no"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(WORKLOADS) + ["custom"])
    ap.add_argument("--n", type=int)
    ap.add_argument("--m", type=int)
    ap.add_argument("--d", type=int)
    ap.add_argument("--c", type=int)
    ap.add_argument("--k", type=int)
    ap.add_argument("--hubness", default=None)
    ap.add_argument("--search-impl", default="auto", choices=["auto", "tc", "tc1", "simt"])
    ap.add_argument("--fused", default="auto", choices=["auto", "on", "off"],
                    help="dual-direction pass (one contraction for reverse + forward kNN)")
    ap.add_argument("--precision", default="auto", choices=["auto", "tf32x3", "screen"],
                    help="candidate search: 3xTF32 keys, or 1xTF32 screen + float64 proof + 3xTF32 "
                         "re-search of unproven rows (same results)")
    ap.add_argument("--data", default="gaussian", choices=["gaussian", "hubby"],
                    help="synthetic distribution (see synth)")
    ap.add_argument("--shard-mode", default="rows", choices=["rows", "cols"],
                    help="N>1, dual-direction pass: shard the source rows (thresholds agreed across "
                         "ranks; default) or the target columns")
    ap.add_argument("--no-hub-scores", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows per direction for the CPU leg")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "port"],
                    help="CPU legs: the unmodified reference from baseline/_ref when present "
                         "(auto), or the oracle port")
    ap.add_argument("--parity-rows", type=int, default=1024,
                    help="rows per direction checked against the oracle outside the timed region")
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the second data distribution (data_variants)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    return ap.parse_args()


def workload(args):
    w = dict(WORKLOADS["c4" if args.workload == "custom" else args.workload])
    for key in ("n", "m", "d", "c", "k"):
        if getattr(args, key) is not None:
            w[key] = getattr(args, key)
    if args.hubness is not None:
        w["hubness"] = None if args.hubness.lower() in ("none", "no") else args.hubness
    w["name"] = (f"{args.workload}: {w['n']}x{w['m']} d={w['d']} fp32 {args.data}, exact kNN "
                 f"c={w['c']} + {w['hubness']} k={w['k']}")
    return w


def make_config(args, w, world):
    """The `config` object of the JSON line -- the same for both arms, so that the driver can
    pair the lines."""
    return {"workload": w["name"], "l2": "inputs (>=1 GB) exceed the 126 MB L2",
            "search_impl": args.search_impl, "fused": args.fused, "precision": args.precision,
            "hub_scores": not args.no_hub_scores,
            "parallelism": (f"{'source rows' if args.shard_mode == 'rows' else 'target columns'} "
                            f"sharded over {world} GPU(s), NCCL exchange + merge kernels")
            if world > 1 else "single GPU"}


def synth(n, d, seed, device, data="gaussian"):
    """Synthetic embeddings: i.i.d. standard normal (almost no near ties, mild hubness), or
    "hubby": a unit-normalised Gaussian mixture of 300 clusters shared by source and target
    (relative noise 0.25: neighbour gaps around the TF32 error bound, strong hubness)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if data == "gaussian":
        return torch.randn((n, d), generator=g, device=device, dtype=torch.float32)
    gc = torch.Generator(device=device)
    gc.manual_seed(12345)                                  # same centres for both sides
    centres = torch.randn((300, d), generator=gc, device=device, dtype=torch.float32)
    x = torch.randn((n, d), generator=g, device=device, dtype=torch.float32)
    assign = torch.randint(0, 300, (n,), generator=g, device=device)
    x.mul_(0.25).add_(centres[assign])
    return torch.nn.functional.normalize(x, dim=1)


# ---------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks line")
# ---------------------------------------------------------------------------
class ClockSampler:
    """ONE `nvidia-smi -lms 200` process for the whole timed region (the recipe's form): spawning
    nvidia-smi per sample initialises NVML every 200 ms and measurably perturbs the step."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self._proc = None

    def start(self):
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None

    def stop(self):
        out = ""
        if self._proc is not None:
            try:
                self._proc.terminate()
                out, _ = self._proc.communicate(timeout=10)
            except Exception:
                try:
                    self._proc.kill()
                except Exception:
                    pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            s = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for nm, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU leg: the reference's SklearnNN + numpy hubness path (oracle port), bounded sample
# ---------------------------------------------------------------------------
def cpu_reference_step(w, sample, source, target):
    """Forward: `sample` source rows vs the full target index; reverse: `sample` target rows vs
    the full source index; rescale + sort on the slice.  Brute-force cost per query is
    independent of the other queries, so queries/s extrapolates linearly."""
    from oracle import kiez_oracle as O

    c, k = w["c"], w["k"]
    t0 = time.perf_counter()
    fwd_d, fwd_i = O.knn_sklearn(source[:sample], target, min(c, target.shape[0]), n_jobs=-1)
    rev_d, _ = O.knn_sklearn(target[:sample], source, min(c, source.shape[0]), n_jobs=-1)
    # the rescale needs reverse statistics of the gathered targets: on the slice we take the
    # statistics of the sampled reverse rows (same arithmetic volume per query as the full run)
    stats_idx = fwd_i % rev_d.shape[0]
    hub = ORACLE_HUB[w["hubness"]]
    if hub == "csls":
        out = O.csls_transform(fwd_d, stats_idx, rev_d)
    elif hub == "nicdm":
        out = O.local_scaling_transform(fwd_d, stats_idx, rev_d, "nicdm")
    elif hub == "mp_gaussian":
        out = O.mp_gaussian_transform(fwd_d, stats_idx, rev_d)
    else:
        out = fwd_d
    O.sort_topk(out, fwd_i, k)
    return time.perf_counter() - t0


def cpu_sample_rows(w, requested):
    if requested:
        return min(requested, w["n"], w["m"])
    # ~10-30 s of CPU work: 4 n m d flop per full step at roughly 100-300 GFLOP/s sustained
    per_query = 4.0 * max(w["n"], w["m"]) * w["d"]
    rows = int(3.0e12 / per_query)
    return int(max(64, min(rows, w["n"], w["m"], 8192)))


def reference_kiez_step(w, sample, source, target):
    """One bounded step of the UNMODIFIED reference (baseline/_ref, loaded through the import
    shims of oracle/ref_shim.py): kiez.Kiez(algorithm=SklearnNN(brute, n_jobs=-1), hubness=...)
    .fit(source[:sample], target).kneighbors(k) -- kiez/kiez.py:160-223.  The reverse pass
    searches all m targets against the `sample` source rows and the forward pass the `sample`
    rows against all m targets: 4 m d flop per query, the per-query cost of the full problem."""
    import warnings

    from oracle import ref_shim

    kiez = ref_shim.load_reference()
    from kiez.neighbors import SklearnNN

    t0 = time.perf_counter()
    algo = SklearnNN(n_candidates=w["c"], metric="euclidean", algorithm="brute", n_jobs=-1)
    inst = kiez.Kiez(n_candidates=w["c"], algorithm=algo, hubness=w["hubness"],
                     hubness_kwargs=dict(HUB_KWARGS.get(w["hubness"], {})))
    inst.fit(source[:sample], target)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        inst.kneighbors(w["k"])
    return time.perf_counter() - t0


def run_reference(args, w):
    """--impl reference: the reference's own CPU implementation of the path on this box's host
    cores, all threads, on a bounded sample of the workload per step.  Runs the unmodified
    reference from baseline/_ref (tools/vendor_reference.sh; /root/reference does not exist on
    the GPU box) when it is there and its hubness class has a CPU path that finishes (the
    reference's DisSimLocal / MP-empiric are per-row Python loops), else the oracle port, which
    performs the same scikit-learn call (NearestNeighbors brute, n_jobs=-1) + numpy rescale."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    from oracle import ref_shim

    def synth_np(n, seed):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((n, w["d"]), dtype=np.float32)
        if args.data == "hubby":          # same distribution as synth() (other random streams)
            centres = np.random.default_rng(12345).standard_normal((300, w["d"]), dtype=np.float32)
            x = 0.25 * x + centres[rng.integers(0, 300, n)]
            x /= np.linalg.norm(x, axis=1, keepdims=True)
        return x

    source = synth_np(w["n"], 0)
    target = synth_np(w["m"], 1)
    sample = cpu_sample_rows(w, args.cpu_sample)
    real = ref_shim.reference_available() and args.cpu_kind != "port"
    step = reference_kiez_step if real else cpu_reference_step
    for _ in range(args.warmup):            # warm-up steps on a small sample (threads, caches)
        step(w, min(sample, 64), source, target)
    times = [step(w, sample, source, target) for _ in range(max(1, args.steps))]
    t = sum(times) / len(times)
    value = sample / t
    what = (f"unmodified kiez 0.5.0 from {ref_shim.REFERENCE_ROOT} (oracle/ref_shim.py): "
            f"Kiez(SklearnNN brute n_jobs=-1, {w['hubness']}).fit(source[:{sample}], target)"
            f".kneighbors({w['k']})") if real else \
           (f"oracle port: {sample} source rows vs all {w['m']} targets + {sample} target rows vs "
            f"all {w['n']} sources + rescale")
    line = {
        "impl": "reference", "metric": "queries_per_s", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, w, args.gpus),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores,
                         "kind": "reference" if real else "port",
                         "sample": what + ", extrapolated linearly in the number of queries"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_sustained": p.get("bf16_tflops_sustained"), "bf16_burst": p.get("bf16_tflops"),
                "hbm_gbs": p.get("hbm_gbs"), "source": "MEASURED_PEAKS.json"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def ncu_traffic(kind):
    """DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed
    `ncu --set full` summary under profiles/ (tools/ncu_summary.py), or None."""
    names = {"screen-dual": ("r02_ncu_knn_screen_dual.txt", "r01_ncu_knn_screen_dual_final.txt"),
             "screen": ("r02_ncu_knn_screen_c50.txt", "r01_ncu_knn_screen_rows.txt"),
             "tf32x3-dual": ("r01_ncu_knn_fused.txt",),
             "tf32x3": ("r01_ncu_knn_tc2_pair_bk32.txt",)}.get(kind, ())
    name = next((nm for nm in names if os.path.exists(os.path.join(ROOT, "profiles", nm))), None)
    path = os.path.join(ROOT, "profiles", name) if name else None
    if not path:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, ms = 0.0, None
    with open(path) as fh:
        for line in fh:
            parts = line.split()
            if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(parts[1]) * unit.get(parts[2], 1.0)
            if len(parts) >= 3 and parts[0] == "gpu__time_duration.sum" and parts[2] == "ms":
                ms = float(parts[1])
    if total <= 0:
        return None, None
    return total, f"profiles/{name}: dram read+write of the captured launch ({ms} ms under ncu)"


def tf32_cublas_tflops(device):
    """cuBLAS TF32 GEMM rate measured live (MEASURED_PEAKS.json has no TF32 entry)."""
    import torch

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((8192, 8192), device=device)
        b = torch.randn((8192, 8192), device=device)
        for _ in range(3):
            a @ b
        best = 0.0
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def oracle_parity(w, inst, final, source, target, rows_wanted, world, rank):
    """Outside the timed region: the kNN results of BOTH directions (and, for the gather-type
    rescalers, the final hubness-reduced neighbours) of a random row sample against the
    scikit-learn brute-force call the reference makes (oracle.knn_sklearn,
    sklearn_nearest_neighbors.py:96-101) with the tie tolerance of the parity tests.  Every rank
    takes part in the device side (the searches are collective at N > 1); rank 0 runs the CPU
    side.  Rows / columns that took the 3xTF32 re-search are force-included (up to 256 each)."""
    import torch

    algo = inst.algorithm
    n, m, c, k = w["n"], w["m"], w["c"], w["k"]
    fwd_d, fwd_i = algo.kneighbors(k=c)
    hub = inst.hubness
    if hasattr(hub, "r_dist_train_") and hasattr(hub, "r_ind_train_"):
        rev_d, rev_i = hub.r_dist_train_, hub.r_ind_train_
    else:
        rev_d, rev_i = algo.kneighbors(k=c, query=algo.target_, s_to_t=False)
    if rank != 0:
        return None
    from oracle import kiez_oracle as O

    t0 = time.perf_counter()
    rng = np.random.default_rng(2)
    rows = rng.choice(n, min(rows_wanted, n), replace=False)
    cols = rng.choice(m, min(rows_wanted, m), replace=False)
    forced = {"rows": 0, "cols": 0}
    researched = getattr(algo, "researched", None)
    if researched is not None:
        extra_r = researched["rows"].cpu().numpy()[:256]
        extra_c = researched["cols"].cpu().numpy()[:256]
        forced = {"rows": int(len(extra_r)), "cols": int(len(extra_c))}
        rows, cols = np.concatenate([rows, extra_r]), np.concatenate([cols, extra_c])
    rows = np.unique(rows)
    # float64 copies of the same fp32 values (what the tests feed the oracle); very large
    # workloads stay fp32 on the host -- scikit-learn upcasts its chunks to float64 itself
    big = (n + m) * w["d"] > 1.2e9
    s_h = source.cpu().numpy() if big else source.cpu().numpy().astype(np.float64)
    t_h = target.cpu().numpy() if big else target.cpu().numpy().astype(np.float64)
    want_fd, want_fi = O.knn_sklearn(s_h[rows], t_h, c, n_jobs=-1)
    r_t = torch.as_tensor(rows, device=fwd_d.device)
    bad_f, first_f = O.count_mismatched_rows(fwd_d[r_t].cpu().numpy(), fwd_i[r_t].cpu().numpy(),
                                             want_fd, want_fi, 1e-5, 5e-6)
    # final result on a few rows: needs the reverse statistics of every target they touch
    oracle_hub = ORACLE_HUB.get(w["hubness"])
    sel = np.sort(rng.choice(len(rows), min(64, len(rows)), replace=False)) \
        if oracle_hub in ("csls", "nicdm", "mp_gaussian") else np.zeros(0, np.int64)
    touched = np.unique(want_fi[sel]) if len(sel) else np.zeros(0, np.int64)
    cols = np.unique(np.concatenate([cols, touched]))
    want_rd, want_ri = O.knn_sklearn(t_h[cols], s_h, c, n_jobs=-1)
    c_t = torch.as_tensor(cols, device=rev_d.device)
    bad_r, first_r = O.count_mismatched_rows(rev_d[c_t].cpu().numpy(), rev_i[c_t].cpu().numpy(),
                                             want_rd, want_ri, 1e-5, 5e-6)
    bad_h, first_h = 0, None
    if len(sel):
        fd, fi = want_fd[sel], want_fi[sel]
        pos = np.searchsorted(cols, fi)                     # target id -> row of want_rd
        if oracle_hub == "csls":
            out = O.csls_transform(fd, pos, want_rd)
        elif oracle_hub == "nicdm":
            out = O.local_scaling_transform(fd, pos, want_rd, "nicdm")
        else:
            out = O.mp_gaussian_transform(fd, pos, want_rd)
        want_d, want_i = O.sort_topk(out, fi, k)
        got_d, got_i = (torch.as_tensor(x) for x in final)
        f_t = torch.as_tensor(rows[sel], device=got_d.device)
        bad_h, first_h = O.count_mismatched_rows(got_d[f_t].cpu().numpy(), got_i[f_t].cpu().numpy(),
                                                 want_d, want_i, 1e-5, 5e-6)
    final_rows = sel
    return {"rows": int(len(rows)), "columns": int(len(cols)), "final_rows": int(len(final_rows)),
            "forced_researched": forced,
            "mismatch": int(bad_f + bad_r + bad_h),
            "mismatch_forward": int(bad_f), "mismatch_reverse": int(bad_r),
            "mismatch_final": int(bad_h), "first": first_f or first_r or first_h,
            "oracle": "oracle.knn_sklearn (the reference's NearestNeighbors brute call) on "
                      + ("fp32 host copies" if big else "float64 copies") +
                      "; tolerance rtol 1e-5 / atol 5e-6, ids exact outside tied runs",
            "seconds": time.perf_counter() - t0}


def run_b200(args, w):
    import torch
    import torch.distributed as dist

    from kiez_b200 import B200, Kiez, _lib, hubness_score

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    source = synth(w["n"], w["d"], 0, device, args.data)
    target = synth(w["m"], w["d"], 1, device, args.data)
    hub_kwargs = HUB_KWARGS.get(w["hubness"], {})

    def make():
        algo = B200(n_candidates=w["c"], metric="euclidean", impl=args.search_impl,
                    distributed=world > 1, precision=args.precision,
                    shard_mode=args.shard_mode,
                    fused={"auto": "auto", "on": True, "off": False}[args.fused])
        algo.host_result = "rank0"          # numpy callers: only rank 0 downloads the result
        return Kiez(n_candidates=w["c"], algorithm=algo, hubness=w["hubness"],
                    hubness_kwargs=dict(hub_kwargs))

    last = {}

    def step(src, tgt, profile=None, collect_stats=False):
        inst = make()
        inst.algorithm._profile = profile
        inst.algorithm._collect_stats = collect_stats
        inst.fit(src, tgt)
        dist_, ind_ = inst.kneighbors(w["k"])
        if not args.no_hub_scores:
            scores = hubness_score(torch.as_tensor(ind_, device=device), w["m"], k=w["k"])
        else:
            scores = None
        last.update(inst=inst, out=(dist_, ind_), scores=scores)
        return dist_, ind_, scores

    def timed(src, tgt, warmup, steps, with_clocks):
        for _ in range(warmup):
            step(src, tgt)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and with_clocks:
            sampler.start()
        profile = []
        launches0 = _lib.launch_counter
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            step(src, tgt, profile)
        ev1.record()
        barrier()
        launches = _lib.launch_counter - launches0
        clocks = sampler.stop() if rank == 0 and with_clocks else None
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        return ms, profile, launches, clocks

    ms, profile, launches, clocks = timed(source, target, args.warmup, args.steps, True)
    ms_per_step = ms / args.steps
    value = w["n"] / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (candidate search) from CUDA events around its launches.
    # Launches are grouped by shape; the group with the most algorithmic flops is the dominant
    # kernel (the dual-direction pass, or the two one-direction passes), the others (threshold
    # sample, overflow re-search) are listed beside it.
    peaks = measured_peaks()
    tf32_peak = peaks["bf16_sustained"] / 2.0
    groups = {}
    for (a, b, nq, ny, d, kind) in profile:
        g = groups.setdefault((nq, ny, d, kind), {"ms": 0.0, "n": 0})
        g["ms"] += a.elapsed_time(b)
        g["n"] += 1
    detail = [{"nq": k[0], "ny": k[1], "d": k[2], "kind": k[3], "launches": v["n"],
               "avg_launch_ms": v["ms"] / v["n"],
               "algorithmic_tflops": 2.0 * k[0] * k[1] * k[2] / (v["ms"] / v["n"] * 1e-3) / 1e12}
              for k, v in groups.items()]
    detail.sort(key=lambda r: -r["nq"] * r["ny"])
    search_ms = sum(v["ms"] for v in groups.values())
    # the dominant kernel = the kind with the most algorithmic flops; its launches may differ in
    # shape (row segments of the dual-direction pass): achieved = sum(flop) / sum(time)
    by_kind = {}
    for r in detail:
        k = by_kind.setdefault(r["kind"], {"flop": 0.0, "ms": 0.0, "n": 0})
        k["flop"] += 2.0 * r["nq"] * r["ny"] * r["d"] * r["launches"]
        k["ms"] += r["avg_launch_ms"] * r["launches"]
        k["n"] += r["launches"]
    if by_kind:
        kind = max(by_kind, key=lambda kk: by_kind[kk]["flop"])
        top_ms, top_n = by_kind[kind]["ms"], by_kind[kind]["n"]
        top_flop = by_kind[kind]["flop"] / top_n
        achieved = by_kind[kind]["flop"] / (top_ms * 1e-3) / 1e12
    else:
        top_ms, top_n, top_flop, achieved, kind = 0.0, 0, 0.0, 0.0, "tf32x3"
    # emit / overflow statistics of the dual-direction pass need extra reductions and host
    # syncs: gathered in one more step OUTSIDE the timed region; the same step feeds the parity
    # check against the oracle
    step(source, target, collect_stats=True)
    barrier()
    algo = last["inst"].algorithm
    fused_stats = getattr(algo, "_fused_stats", None)
    search_stats = dict(getattr(algo, "search_stats", {}))
    parity = None
    if args.parity_rows > 0:
        parity = oracle_parity(w, last["inst"], last["out"], source, target, args.parity_rows,
                               world, rank)
        barrier()
    mmas = 1.0 if kind.startswith("screen") else 3.0      # MMAs issued per algorithmic MAC
    kernel_names = {
        "screen-dual": "knn_screen_kernel<dual> (1xTF32 tcgen05 cta_group::2, resident query tile, "
                       "dual-direction, fused top-c; float64 proof + 3xTF32 re-search of unproven rows)",
        "screen": "knn_screen_kernel (1xTF32 tcgen05 cta_group::2, resident query tile, fused top-c; "
                  "float64 proof + 3xTF32 re-search of unproven rows)",
        "tf32x3-dual": "knn_fused_kernel (3xTF32 tcgen05 cta_group::2, dual-direction, fused top-c)",
        "tf32x3": "knn_tc2_kernel (3xTF32 tcgen05 cta_group::2 + fused top-c)",
    }
    roofline = {
        "bound": "tensor",
        "kernel": kernel_names[kind],
        "achieved": achieved, "peak": tf32_peak / mmas, "unit": "TFLOP/s",
        "frac": achieved / (tf32_peak / mmas), "traffic": ncu_traffic(kind)[0],
        "traffic_unit": "bytes per captured launch (tensor-bound kernel: DRAM is ~1 % utilised)",
        "traffic_source": ncu_traffic(kind)[1],
        "issued_tf32_tflops": mmas * achieved, "tf32_peak": tf32_peak,
        "peak_source": f"{peaks['source']} bf16_tflops_sustained / 2 (TF32 rate)" +
                       (" / 3 (3xTF32 issues 3 MMAs per algorithmic MAC)" if mmas == 3.0 else ""),
        "launches": top_n, "avg_launch_ms": (top_ms / top_n) if top_n else None,
        "kernel_share_of_step": (top_ms / ms) if ms else None,
        "all_search_launches_share_of_step": (search_ms / ms) if ms else None,
        "algorithmic_flop_per_launch": top_flop,
        "algorithmic_flop_per_step": (top_flop * top_n / args.steps) if top_n else None,
        "search_launches": detail,
        "dual_direction": fused_stats, "screen": search_stats,
    }
    if rank == 0:
        try:
            roofline["tf32_cublas_tflops_live"] = tf32_cublas_tflops(device)
            # `frac` uses the driver's sustained bf16 figure / 2; under the power cap a TF32 GEMM
            # sustains a little more than half the bf16 rate, so the fraction of what cuBLAS
            # reaches in this very run is reported next to it
            roofline["frac_of_live_cublas"] = mmas * achieved / roofline["tf32_cublas_tflops_live"]
        except Exception as exc:  # pragma: no cover
            roofline["tf32_cublas_tflops_live"] = f"failed: {exc}"

    # end to end through the public API with HOST buffers, H2D + D2H inside the timing: first
    # with PAGEABLE numpy arrays (what a kiez user has: the headline `e2e.value`), then with
    # pinned ones.  One untimed step first (pinned staging buffers are allocated once per process).
    e2e = None
    if not args.no_e2e:
        def e2e_run(src_h, tgt_h, n_steps):
            step(src_h, tgt_h)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_steps):
                d_, i_, _s = step(src_h, tgt_h)       # numpy in -> numpy out (D2H inside)
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / n_steps), d_, i_

        n_e2e = max(1, args.e2e_steps)
        src_h, tgt_h = source.cpu(), target.cpu()
        t_page, d_, i_ = e2e_run(src_h.numpy(), tgt_h.numpy(), n_e2e)
        d2h = int(d_.nbytes + i_.nbytes) if isinstance(d_, np.ndarray) else 0
        # when the last chunk of each upload job of the last step was handed to the copy engine,
        # in ms after fit() began (jobs in order: row sample, target, source; N > 1: the slices)
        up_ms = None
        algo_e = last["inst"].algorithm
        if getattr(algo_e, "_uploader", None) is not None:
            up_ms = [round(1e3 * (j.t_enqueued[-1] - algo_e._t_fit_start), 1)
                     for j in algo_e._uploader.jobs]
        src_p, tgt_p = src_h.pin_memory(), tgt_h.pin_memory()
        del src_h, tgt_h
        t_pin, _d, _i = e2e_run(src_p.numpy(), tgt_p.numpy(), n_e2e)
        # whole-job bytes: with N ranks every rank uploads its 1/N slice of both matrices (the
        # rest arrives over NVLink) and rank 0 downloads the result
        e2e = {"value": w["n"] / t_page, "unit": "queries/s",
               "h2d_bytes_per_step": int(src_p.numel() * 4 + tgt_p.numel() * 4),
               "d2h_bytes_per_step": d2h, "steps": n_e2e, "host_buffers": "pageable numpy",
               "ms_per_step": 1e3 * t_page,
               "pinned": {"value": w["n"] / t_pin, "ms_per_step": 1e3 * t_pin},
               "fraction_of_device_value": (w["n"] / t_page) / value,
               "upload_jobs_enqueued_ms": up_ms,
               "timer": "host wall clock around fit+kneighbors(+hub scores) incl. copies, max over ranks"}
        del src_p, tgt_p

    # the same measurement on the other synthetic distribution (fewer steps)
    variants = None
    if not args.no_variants:
        other = "hubby" if args.data == "gaussian" else "gaussian"
        del source, target
        src2 = synth(w["n"], w["d"], 0, device, other)
        tgt2 = synth(w["m"], w["d"], 1, device, other)
        ms2, prof2, _l, _c = timed(src2, tgt2, 1, min(args.steps, 3), False)
        ms2 /= min(args.steps, 3)
        step(src2, tgt2, collect_stats=True)
        barrier()
        algo2 = last["inst"].algorithm
        par2 = oracle_parity(w, last["inst"], last["out"], src2, tgt2, min(args.parity_rows, 256),
                             world, rank) if args.parity_rows > 0 else None
        barrier()
        kinds = sorted({p[5] for p in prof2})
        variants = {other: {"value": w["n"] / (ms2 * 1e-3), "ms_per_step": ms2,
                            "steps": min(args.steps, 3), "search_kinds": kinds,
                            "screen": dict(algo2.search_stats), "parity_check": par2}}
        source, target = src2, tgt2

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # reported at N=1 only
        from oracle import ref_shim

        cores = len(os.sched_getaffinity(0))
        sample = cpu_sample_rows(w, args.cpu_sample)
        src_np, tgt_np = source.cpu().numpy(), target.cpu().numpy()   # same shapes either way
        real = ref_shim.reference_available() and args.cpu_kind != "port"
        fn = reference_kiez_step if real else cpu_reference_step
        fn(w, min(sample, 64), src_np, tgt_np)
        t_cpu = fn(w, sample, src_np, tgt_np)
        cpu = {"value": sample / t_cpu, "unit": "queries/s", "cores": cores,
               "kind": "reference" if real else "port",
               "sample": (f"unmodified kiez 0.5.0 (baseline/_ref): Kiez(SklearnNN brute, "
                          f"{w['hubness']}).fit(source[:{sample}], target).kneighbors({w['k']})"
                          if real else
                          f"oracle port: {sample} source rows vs all {w['m']} targets + {sample} "
                          f"target rows vs all {w['n']} sources + rescale")
                         + f" ({t_cpu:.1f} s), extrapolated linearly"}

    if rank == 0:
        line = {
            "metric": "queries_per_s", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "tf32" if kind.startswith("screen") else "tf32x3", "data": "synthetic",
            "config": make_config(args, w, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches, "parity_check": parity, "data_variants": variants,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    w = workload(args)
    # Exactly ONE line may reach stdout (the JSON): libraries such as NCCL print banners to
    # fd 1, so route fd 1 to stderr for the run and print the JSON line to the real stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)
    real_stdout.flush()


if __name__ == "__main__":
    main()
