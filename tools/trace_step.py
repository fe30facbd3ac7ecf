"""Where does a C4 step spend GPU time OUTSIDE kernels?  Runs warm-up steps, then one step under
torch.profiler (CUPTI), and prints the idle gaps between consecutive kernels on the device.

    python tools/trace_step.py [--n 1000000 --m 1000000 --d 256 --c 10 --k 10] > gpurun_out/trace.txt
    torchrun --nproc-per-node N tools/trace_step.py ...     # one report per rank on stdout
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.profiler import ProfilerActivity, profile

from kiez_b200 import B200, Kiez


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=256)
    ap.add_argument("--c", type=int, default=10)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--min-gap-us", type=float, default=200.0)
    ap.add_argument("--hubness", default="CSLS")
    ap.add_argument("--method", default=None, help="hubness method kwarg (nicdm, normal, ...)")
    ap.add_argument("--fused", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--shard-mode", default="rows", choices=["rows", "cols"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    src = torch.randn((args.n, args.d), generator=g, device=dev)
    g.manual_seed(1)
    tgt = torch.randn((args.m, args.d), generator=g, device=dev)

    def step():
        inst = Kiez(n_candidates=args.c,
                    algorithm=B200(n_candidates=args.c, distributed=world > 1,
                                   shard_mode=args.shard_mode,
                                   fused={"auto": "auto", "on": True, "off": False}[args.fused]),
                    hubness=None if args.hubness.lower() == "none" else args.hubness,
                    hubness_kwargs={"method": args.method} if args.method else {})
        inst.fit(src, tgt)
        return inst.kneighbors(args.k)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA events recorded")
        return
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    print(f"rank {os.environ.get('RANK', '0')}/{world}: step span {1e-3 * (t1 - t0):.1f} ms, kernels+memops busy {1e-3 * busy:.1f} ms, "
          f"idle {1e-3 * (t1 - t0 - busy):.1f} ms, {len(evs)} device activities")
    print(f"gaps >= {args.min_gap_us:.0f} us (after -> before):")
    end = evs[0].time_range.end
    prev = evs[0]
    for e in evs[1:]:
        gap = e.time_range.start - end
        if gap >= args.min_gap_us:
            print(f"  {1e-3 * gap:8.2f} ms   t={1e-3 * (end - t0):8.1f} ms   {prev.name[:60]:<60s} -> {e.name[:60]}")
        if e.time_range.end > end:
            end, prev = e.time_range.end, e
    print("largest device activities:")
    for e in sorted(evs, key=lambda e: e.time_range.start - e.time_range.end)[:12]:
        print(f"  {1e-3 * (e.time_range.end - e.time_range.start):8.2f} ms  {e.name[:90]}")
    print("device time by activity name:")
    by_name = {}
    for e in evs:
        t = by_name.setdefault(e.name[:70], [0.0, 0])
        t[0] += e.time_range.end - e.time_range.start
        t[1] += 1
    for name, (us, cnt) in sorted(by_name.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"  {1e-3 * us:9.2f} ms  x{cnt:<4d} {name}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
