"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches_c4.csv "<command>" > profiles/rNN_launches_x_summary.txt

Kernels of the synthetic-data generator and of the live cuBLAS TF32 peak probe are listed but
left out of the step total (they are not part of a step).
"""
import csv
import re
import sys
from collections import OrderedDict

NOT_STEP = ("distribution_elementwise", "cutlass", "gemm", "at::native", "at::<unnamed>")


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name[:72]


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r["Metric Unit"]
            ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            rows.append((r["Kernel Name"], ms))
    agg = OrderedDict()
    for name, ms in rows:
        a = agg.setdefault(short(name), [0, 0.0, name])
        a[0] += 1
        a[1] += ms
    step_total = sum(a[1] for k, a in agg.items() if not any(t in a[2] for t in NOT_STEP)
                     or "reduce_kernel" in a[2])
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}")
    print("# per-launch times are cold-cache and serialised: compare SHARES of the step, not absolutes")
    print(f"# {len(rows)} launches recorded; step kernels total {step_total:.1f} ms")
    print(f"{'kernel':<74s}{'n':>5s}{'total ms':>12s}{'share':>9s}")
    for k, (n, ms, full) in agg.items():
        in_step = not any(t in full for t in NOT_STEP) or "reduce_kernel" in full
        share = f"{100.0 * ms / step_total:7.3f}%" if in_step else "      -"
        print(f"{k:<74s}{n:>5d}{ms:>12.3f}  {share}")


if __name__ == "__main__":
    main()
