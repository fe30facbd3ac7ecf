"""A/B timing of the candidate-search kernel of one library build (KB2_LIB) under sustained
load: ms per launch, SM clock and board power (the kernel is power-capped on B200)."""
import os
import statistics
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kiez_b200 import _lib as lib


def main():
    nq, ny, d, cap = 131072, 262144, 256, int(os.environ.get("AB_CAP", "16"))
    launches = int(os.environ.get("AB_LAUNCHES", "30"))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    q = torch.randn((nq, d), device=dev, generator=g)
    y = torch.randn((ny, d), device=dev, generator=g)
    dpad = lib.lib.kb2_padded_dim(d)
    st = lib.stream_ptr()

    def prep(x):
        n = x.shape[0]
        hi = torch.empty((n, dpad), device=dev)
        lo = torch.empty((n, dpad), device=dev)
        key = torch.empty(n, device=dev)
        lib.call("kb2_prepare_rows", lib.ptr(x), n, d, d, None, 0, lib.ptr(hi), lib.ptr(lo), dpad,
                 lib.ptr(key), None, st)
        return hi, lo, key

    qh, ql, _ = prep(q)
    yh, yl, yk = prep(y)
    cand = torch.empty((nq, cap), dtype=torch.int32, device=dev)

    def launch():
        lib.call("kb2_knn_candidates", lib.KNN_TC, lib.ptr(qh), lib.ptr(ql), nq, lib.ptr(yh),
                 lib.ptr(yl), lib.ptr(yk), ny, dpad, cap, 1, lib.ptr(cand), None, st)

    samples = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw",
                                  "--format=csv,noheader,nounits", "-i", "0"],
                                 capture_output=True, text=True).stdout.strip().split(",")
            try:
                samples.append((time.time(), float(out[0]), float(out[1])))
            except Exception:
                pass
            stop.wait(0.1)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    thr = threading.Thread(target=sampler, daemon=True)
    thr.start()
    evs = []
    for _ in range(launches):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    stop.set()
    thr.join()
    ms = [a.elapsed_time(b) for a, b in evs]
    tail = ms[len(ms) // 3:]
    half = samples[len(samples) // 3:]
    clk = statistics.median(s[1] for s in half) if half else -1
    pw = statistics.median(s[2] for s in half) if half else -1
    flop = 2.0 * nq * ny * d
    avg = sum(tail) / len(tail)
    print(f"{os.environ.get('AB_NAME', '?'):22s} cfg={os.environ.get('KB2_TC_CONFIG', 'auto'):7s} "
          f"ms/launch={avg:8.2f} (first {ms[0]:.1f}) issuedTF={3 * flop / avg / 1e9:7.1f} "
          f"clk={clk:6.0f} MHz power={pw:6.0f} W  cycles/launch={avg * clk * 1e3:.3e}", flush=True)


if __name__ == "__main__":
    main()
