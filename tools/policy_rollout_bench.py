#!/usr/bin/env python
"""Closed-loop rollouts: T steps of obs -> per-agent actor (Linear-ReLU-Linear-softmax) -> env.step, as
  (a) ONE launch of mpe_rollout_policy (actors evaluated inside the kernel, state and observations in registers),
  (b) the same actors as torch modules + env.step, all captured in one CUDA graph (rollout.GraphedRollout).
Device time per 65 536-world step of each."""
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
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build(quiet=True)
    from multiagent_particle_envs_b200 import make_env
    from multiagent_particle_envs_b200.rollout import GraphedRollout
    dev = torch.device("cuda", 0)
    n, T, H = args.num_envs, args.steps, args.hidden
    env = make_env(args.scenario, num_envs=n, device=dev)
    env.reuse_buffers = True
    env.reset()
    nw = env.world.native
    torch.manual_seed(0)
    mods = [torch.nn.Sequential(torch.nn.Linear(od, H), torch.nn.ReLU(), torch.nn.Linear(H, 5)).to(dev) for od in nw.obs_dims]
    res = {"config": {"scenario": args.scenario, "n_env": n, "T": T, "hidden": H}}
    # (a) in-kernel actors
    for _ in range(2):
        env.rollout_policy(mods, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        env.rollout_policy(mods, T)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / (args.reps * T)
    res["in_kernel"] = {"us_per_step": 1e6 * sec, "env_steps_per_sec": n / sec}
    # (b) torch actors + env.step in one CUDA graph
    env2 = make_env(args.scenario, num_envs=n, device=dev)
    env2.reset()

    def policy(obs_n):
        return [torch.softmax(m(o), -1) for m, o in zip(mods, obs_n)]

    ro = GraphedRollout(env2, policy, steps=T)
    for _ in range(2):
        ro.run()
    torch.cuda.synchronize()
    e0.record(ro.stream)
    with torch.cuda.stream(ro.stream):
        e0.record(ro.stream)
        for _ in range(args.reps):
            ro.run()
        e1.record(ro.stream)
    torch.cuda.synchronize()
    sec2 = e0.elapsed_time(e1) / 1e3 / (args.reps * T)
    res["graphed_torch"] = {"us_per_step": 1e6 * sec2, "env_steps_per_sec": n / sec2}
    res["speedup"] = sec2 / sec
    print(json.dumps(res))


if __name__ == "__main__":
    main()
