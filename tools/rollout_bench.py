#!/usr/bin/env python
"""Time the K-step open-loop rollout kernel (mpe_rollout) against K launches of the fused step on the same
pre-generated actions: device time per env-step and HBM bytes per env-step of each form."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build(quiet=True)
    from multiagent_particle_envs_b200 import _lib, make_env
    dev = torch.device("cuda", 0)
    n, T = args.num_envs, args.steps
    env = make_env(args.scenario, num_envs=n, device=dev)
    env.reuse_buffers = True
    env.reset()
    nw = env.world.native
    gen = torch.Generator(device=dev).manual_seed(0)
    ring = 8      # independent action sequences so that action reads come from HBM
    seqs = [[torch.softmax(torch.randn(T, n, d, device=dev, generator=gen), -1).contiguous() for d in nw.act_dims] for _ in range(ring)]
    ptrs = [_lib.ptr_array([t.data_ptr() for t in s]) for s in seqs]
    step_ptrs = [[_lib.ptr_array([t[k].data_ptr() for t in s]) for k in range(T)] for s in seqs]
    flags = env._flags()
    stream = torch.cuda.Stream(dev)
    res = {}
    with torch.cuda.stream(stream):
        for name in ("rollout", "steps"):
            def body(r):
                if name == "rollout":
                    nw.rollout(ptrs[r % ring], T, nw.out, flags)
                else:
                    for k in range(T):
                        nw.step(step_ptrs[r % ring][k], nw.out, flags)
            body(0)
            stream.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                for r in range(ring):
                    body(r)
            graph.replay()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                graph.replay()
            e1.record(stream)
            stream.synchronize()
            sec = e0.elapsed_time(e1) / 1e3 / (args.reps * ring * T)
            res[name] = {"us_per_env_batch_step": 1e6 * sec, "env_steps_per_sec": n / sec}
    act_b = 4 * sum(nw.act_dims)
    res["bytes_per_env_step"] = {"steps": nw.bytes_per_env_step, "rollout": act_b + (nw.bytes_per_env_step - act_b) / T}
    res["config"] = {"scenario": args.scenario, "n_env": n, "T": T}
    res["speedup"] = res["steps"]["us_per_env_batch_step"] / res["rollout"]["us_per_env_batch_step"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
