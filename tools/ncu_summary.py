"""Summarise an ncu report (--set full) of one kernel into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg",
    "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__mem_tensor_reads_op_ldt.sum.pct_of_peak_sustained_elapsed",
]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True,
                         text=True).stdout
    return out


def cuda_lines(path, tot_s, tot_i):
    """Samples / instructions aggregated per CUDA source line (needs -lineinfo + --import-source)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    fname, hdr, lines = "?", None, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                lines.append((int(d["# Samples"]), int(d["Instructions Executed"]), fname, r[0],
                              r[1].strip()))
            except (KeyError, ValueError):
                pass
    if not lines:
        return
    print("# hottest CUDA source lines (samples %, executed warp instructions %)")
    for smp, ins, f, ln, src in sorted(lines, reverse=True)[:30]:
        print(f"  {100.0 * smp / tot_s:5.2f}% {100.0 * ins / tot_i:5.2f}%  {f}:{ln:>4s}  {src[:90]}")


def main():
    path = sys.argv[1]
    raw = list(csv.reader(io.StringIO(ncu(path, "raw"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print(f"# ncu --set full summary of {path}")
    print(f"kernel: {d.get('Kernel Name', ('', '?'))[1]}")
    for k in KEYS:
        if k in d:
            print(f"{k:100s} {d[k][1]:>20s} {d[k][0]}")
    src = list(csv.reader(io.StringIO(ncu(path, "source"))))
    # find the header row of the source page
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[hi]
    rows = [dict(zip(h, r)) for r in src[hi + 1:] if len(r) == len(h)]
    tot_s = sum(int(r["# Samples"]) for r in rows) or 1
    tot_i = sum(int(r["Instructions Executed"]) for r in rows) or 1
    print(f"\n# warp-state samples: {tot_s}, warp instructions executed: {tot_i}")
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[c]) for r in rows) for c in stalls}
    print("# stall reasons (all warps, % of samples)")
    for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  {c:28s} {100.0 * v / tot_s:6.2f}%")
    print("# hottest SASS instructions by samples")
    for r in sorted(rows, key=lambda r: -int(r["# Samples"]))[:25]:
        top = max(stalls, key=lambda c: int(r[c]))
        print(f"  {100.0 * int(r['# Samples']) / tot_s:5.2f}% samples {100.0 * int(r['Instructions Executed']) / tot_i:5.2f}% inst  "
              f"{r['Source'][:64]:64s} {top}")
    cuda_lines(path, tot_s, tot_i)
    mem = [r for r in rows if "LDL" in r["Source"] or "STL" in r["Source"]]
    print(f"# local-memory instructions in SASS: {len(mem)} "
          f"(executed {sum(int(r['Instructions Executed']) for r in mem)})")


if __name__ == "__main__":
    main()
