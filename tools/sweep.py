#!/usr/bin/env python
"""Batch-size sweep of the fused step kernel (device-resident, CUDA-graph replay over a ring of batches
whose working set exceeds 2x L2): env-steps/s and achieved fraction of the measured HBM roofline for the
BASELINE.json scenarios at 1k..1M worlds.  Writes JSON lines (one per point) to stdout / --out."""
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
    ap.add_argument("--seconds", type=float, default=0.25)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build(quiet=True)
    from multiagent_particle_envs_b200 import _lib, make_env
    from bench import measured_peak
    peak, src = measured_peak()
    dev = torch.device("cuda", 0)
    out = open(args.out, "w") if args.out else None
    for sc in args.scenarios.split(","):
        for n in [int(x) for x in args.sizes.split(",")]:
            spec = SCENARIOS.get(sc, dict(name=sc, kw={}))
            probe = make_env(spec["name"], num_envs=n, device=dev, **spec["kw"])
            bpe = probe.world.native_shapes().bytes_per_env_step
            ring_n = max(2, min(64, int(2.2 * 126 * 2**20 / (bpe * n)) + 1))
            ring = []
            for b in range(ring_n):
                env = make_env(spec["name"], num_envs=n, device=dev, seed=b, **spec["kw"])
                env.reset()
                nw = env.world.native
                gen = torch.Generator(device=dev).manual_seed(b)
                acts = []
                for d in nw.act_dims:
                    p = torch.softmax(torch.randn(n, 5, device=dev, generator=gen), 1)
                    if d > 5:
                        p = torch.cat([p, torch.rand(n, d - 5, device=dev, generator=gen)], 1)
                    acts.append(p.contiguous())
                ring.append((env, nw, acts, _lib.ptr_array([t.data_ptr() for t in acts]), env._flags()))
            stream = torch.cuda.Stream(dev)
            with torch.cuda.stream(stream):
                for env, nw, acts, ptrs, flags in ring:
                    nw.step(ptrs, nw.out, flags)
                stream.synchronize()
                graph = torch.cuda.CUDAGraph()
                reps_in_graph = max(1, 64 // ring_n)
                with torch.cuda.graph(graph, stream=stream):
                    for _ in range(reps_in_graph):
                        for env, nw, acts, ptrs, flags in ring:
                            nw.step(ptrs, nw.out, flags)
                steps_per_replay = reps_in_graph * ring_n
                for _ in range(3):
                    graph.replay()
                stream.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                graph.replay()
                e1.record(stream)
                stream.synchronize()
                est = e0.elapsed_time(e1) / 1e3
                replays = max(3, int(args.seconds / max(est, 1e-6)))
                e0.record(stream)
                for _ in range(replays):
                    graph.replay()
                e1.record(stream)
                stream.synchronize()
                sec = e0.elapsed_time(e1) / 1e3
            steps = replays * steps_per_replay
            us = 1e6 * sec / steps
            gbs = bpe * n / (sec / steps) / 1e9
            rec = {"scenario": sc, "n_env": n, "ring": ring_n, "us_per_step": us, "env_steps_per_sec": n * steps / sec,
                   "bytes_per_env_step": bpe, "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak, "peak_source": src}
            line = json.dumps(rec)
            print(line, flush=True)
            if out:
                out.write(line + "\n")
                out.flush()
            del ring, graph
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
