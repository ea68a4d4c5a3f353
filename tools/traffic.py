#!/usr/bin/env python
"""Steady-state DRAM traffic of the fused step.  Run UNDER ncu in RANGE replay with the cache
state left alone, so that the whole steady-state loop -- kernels overlapping through programmatic dependent launch,
dirty observation lines being written back while later steps run, inputs possibly still L2-resident from their last
visit -- is one measured range:

    ncu --replay-mode range --cache-control none --clock-control none \\
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \\
        --csv --log-file gpurun_out/traffic_X.csv python tools/traffic.py --scenario simple_spread --num-envs 65536
    python tools/ncu_traffic_summary.py gpurun_out/traffic_X.csv --scenario simple_spread --num-envs 65536 --launches <printed>

The script steps bench.py's own ring (same batches, same sizing) for `--warm` passes, then brackets `--passes` whole
passes with cudaProfilerStart/Stop.  DRAM bytes of the range / launches in the range = steady-state traffic per
launch.  (Per-KERNEL counters cannot see it: ncu serialises kernels and the write-back of a step's 18 MB of dirty lines
happens after the kernel that produced them has finished -- round 2's first attempt read 35 KB of writes per launch.)"""
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
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--ring", type=int, default=0)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build(quiet=True)
    import bench
    kw = {"num_agents": args.num_agents} if args.num_agents is not None else {}
    dev = torch.device("cuda", 0)
    ring = bench.Ring(args.scenario, kw, args.num_envs, dev, 0, 1, requested_ring=args.ring)
    with torch.cuda.stream(ring.stream):
        for _ in range(args.warm):
            for i in range(ring.R):
                ring.step_slot(i)
        ring.flush_l2()
        ring.stream.synchronize()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(args.passes):
            for i in range(ring.R):
                ring.step_slot(i)
        ring.stream.synchronize()
        torch.cuda.profiler.stop()
    print("traffic.py: ring=%d launches_in_range=%d" % (ring.R, args.passes * ring.R))


if __name__ == "__main__":
    main()
