#!/usr/bin/env python
"""Batch-size sweep of the fused step (BASELINE north star: simple_spread N=3 and simple_tag at 1k..1M worlds on
1 / 2 / 4 / 8 B200): device-resident, strictly serialized launches replayed from CUDA graphs over bench.py's ring
(inputs > 2x L2), env-steps/s and achieved fraction of the measured HBM roofline per point.

    python tools/sweep.py --out profiles/r2_sweep_n1.jsonl                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
        --master-port 29511 tools/sweep.py --out profiles/r2_sweep_n8.jsonl                    # --num-envs is per GPU

Under torchrun every rank sweeps its own shard (weak scaling); per point the ranks exchange one all-gather of
(env_steps, seconds) and rank 0 writes whole-job throughput with time = max over ranks."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SCENARIOS = {
    "simple_spread": dict(name="simple_spread", kw={}),
    "simple_tag": dict(name="simple_tag", kw={}),
    "simple_spread_n6": dict(name="simple_spread", kw={"num_agents": 6}),
    "simple_world_comm": dict(name="simple_world_comm", kw={}),
    "simple": dict(name="simple", kw={}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", default="simple_spread,simple_tag,simple_spread_n6,simple_world_comm")
    ap.add_argument("--sizes", default="1024,4096,16384,65536,262144,1048576")
    ap.add_argument("--seconds", type=float, default=0.2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    import bench
    bench.pin_to_gpu_numa(local_rank)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build(quiet=True)
    from multiagent_particle_envs_b200.sharding import aggregate_counters
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, src = bench.measured_peak()
    out = open(args.out, "w") if (args.out and rank == 0) else None
    spin = int(200e-6 * getattr(torch.cuda.get_device_properties(dev), "clock_rate", 1.9e6) * 1e3)
    for sc in args.scenarios.split(","):
        spec = SCENARIOS.get(sc, dict(name=sc, kw={}))
        for n in [int(x) for x in args.sizes.split(",")]:
            ring = bench.Ring(spec["name"], spec["kw"], n, dev, rank, world, max_ring=64)
            k = ring.R * max(1, 64 // ring.R)                 # whole ring passes, 64+ launches per graph
            plan = ring.plan(k)
            with torch.cuda.stream(ring.stream):
                for _ in range(3):
                    ring.run(plan)
                ring.stream.synchronize()
            est = ring.timed(plan, spin)
            reps = max(3, int(args.seconds / max(est, 1e-6)))
            if world > 1:
                t = torch.tensor([reps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                reps = int(t.item())
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ring.stream):
                torch.cuda._sleep(spin)
                e0.record(ring.stream)
                for _ in range(reps):
                    ring.run(plan)
                e1.record(ring.stream)
                ring.stream.synchronize()
            sec = e0.elapsed_time(e1) / 1e3
            steps = reps * k
            total, mx, _ = aggregate_counters(n * steps, sec)
            if rank == 0:
                per_gpu_gbs = ring.bytes_per_env * n / (mx / steps) / 1e9
                rec = {"scenario": sc, "n_gpus": world, "n_env_per_gpu": n, "ring": ring.R,
                       "ring_inputs_exceed_2xL2": ring.R * ring.input_bytes_per_env * n > 2 * bench.L2_BYTES, "steps": steps,
                       "us_per_step": 1e6 * mx / steps, "env_steps_per_sec": total / mx,
                       "bytes_per_env_step": ring.bytes_per_env, "achieved_gbs_per_gpu": per_gpu_gbs, "peak_gbs": peak,
                       "frac": per_gpu_gbs / peak, "peak_source": src}
                line = json.dumps(rec)
                print(line, flush=True)
                if out:
                    out.write(line + "\n")
                    out.flush()
            del ring, plan
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
