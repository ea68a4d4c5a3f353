#!/usr/bin/env python
"""Steady-state DRAM traffic of the fused step.  Run UNDER ncu with the cache state left
alone between kernels, so that what one step leaves in the 126 MB L2 (its dirty observation lines, its inputs)
is what the next steps see, exactly as in bench.py's timed loop:

    ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,\
lts__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:mpe_kernel --csv --log-file gpurun_out/traffic_X.csv \
        python tools/traffic.py --scenario simple_spread --num-envs 65536 --passes 3
    python tools/ncu_traffic_summary.py gpurun_out/traffic_X.csv --scenario simple_spread --num-envs 65536

The script steps bench.py's own ring (same batches, same sizing on input bytes) `passes` times in order with
plain (un-graphed) launches; the summary drops the ring-construction launches and the first pass and averages
DRAM bytes per fused-step launch over the rest.  Dirty lines written by launch k are evicted -- and counted -- while
later launches run, so the per-launch average over whole passes is the steady-state traffic."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--num-agents", type=int, default=None)
    ap.add_argument("--passes", type=int, default=3)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build(quiet=True)
    import bench
    kw = {"num_agents": args.num_agents} if args.num_agents is not None else {}
    dev = torch.device("cuda", 0)
    ring = bench.Ring(args.scenario, kw, args.num_envs, dev, 0, 1)
    with torch.cuda.stream(ring.stream):
        for _ in range(args.passes):
            for i in range(ring.R):
                ring.step_slot(i)
        ring.stream.synchronize()
    print("traffic.py: ring=%d passes=%d fused launches=%d (+%d at ring construction)" % (ring.R, args.passes, args.passes * ring.R, ring.R))


if __name__ == "__main__":
    main()
