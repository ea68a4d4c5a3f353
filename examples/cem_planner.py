#!/usr/bin/env python
"""Sampling-based planning with the K-step rollout kernel: cross-entropy-method search for an open-loop action sequence
that covers the landmarks of ONE simple_spread world.  Every candidate sequence is scored in its own copy of that world;
`env.rollout` advances all copies H steps in a single launch (state in registers, only actions read per step).

    python examples/cem_planner.py --candidates 65536 --horizon 25 --iters 6
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from make_env import make_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=65536)
    ap.add_argument("--horizon", type=int, default=25)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--elite", type=float, default=0.02)
    args = ap.parse_args()
    N, H = args.candidates, args.horizon
    env = make_env("simple_spread", num_envs=N, seed=3)
    env.reset()
    nw = env.world.native
    dev = nw.device
    # every candidate starts from the same world: broadcast world 0's initial state
    pv0 = nw.agent_pv[:, :1].clone()
    lm0 = nw.lm_p[:, :1].clone()
    mean = torch.zeros(env.n, H, 5, device=dev)            # logits of the movement distribution per agent and step
    std = torch.ones(env.n, H, 5, device=dev) * 2.0
    k = max(2, int(args.elite * N))
    for it in range(args.iters):
        logits = mean[:, :, None, :] + std[:, :, None, :] * torch.randn(env.n, H, N, 5, device=dev)
        seqs = [torch.softmax(logits[i], -1).contiguous() for i in range(env.n)]      # [H, N, 5] per agent
        nw.agent_pv.copy_(pv0.expand_as(nw.agent_pv))
        nw.lm_p.copy_(lm0.expand_as(nw.lm_p))
        obs_n, ret_n, done_n, _ = env.rollout(seqs)        # ONE launch = H steps of N worlds
        ret = ret_n[0]                                      # shared reward: identical for every agent
        top = torch.topk(ret, k).indices
        mean = logits[:, :, top, :].mean(2)
        std = logits[:, :, top, :].std(2) + 0.05
        print("iter %d: best return %.3f, elite mean %.3f, population mean %.3f"
              % (it, float(ret.max()), float(ret[top].mean()), float(ret.mean())))
    print("planned %d-step sequence for %d agents; final elite return %.3f" % (H, env.n, float(ret[top].mean())))


if __name__ == "__main__":
    main()
