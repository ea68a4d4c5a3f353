#!/usr/bin/env python
"""A device-resident rollout the way a GPU trainer would run it: per-agent MLP policies (obs -> 64 -> 64 -> 5 logits,
softmax) and `env.step`, 25 steps + an episode reset, captured ONCE in a CUDA graph (GraphedRollout) and replayed.

    python examples/graphed_mlp_rollout.py --scenario simple_spread --num-envs 65536 --episodes 40
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from make_env import make_env  # noqa: E402
from multiagent_particle_envs_b200.rollout import GraphedRollout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--episodes", type=int, default=40)
    ap.add_argument("--episode-length", type=int, default=25)
    ap.add_argument("--hidden", type=int, default=64)
    args = ap.parse_args()

    env = make_env(args.scenario, num_envs=args.num_envs)
    act_dims = env.world.native_shapes().act_dims
    obs_dims = [s.shape[0] for s in env.observation_space]
    torch.manual_seed(0)
    nets = [torch.nn.Sequential(torch.nn.Linear(o, args.hidden), torch.nn.ReLU(), torch.nn.Linear(args.hidden, args.hidden),
                                torch.nn.ReLU(), torch.nn.Linear(args.hidden, a)).cuda() for o, a in zip(obs_dims, act_dims)]

    def policy(obs_n):
        out = []
        for net, o, a in zip(nets, obs_n, act_dims):
            z = net(o)
            out.append(torch.cat([torch.softmax(z[:, :5], 1), torch.sigmoid(z[:, 5:])], 1) if a > 5 else
                       (torch.softmax(z, 1) if a == 5 else torch.sigmoid(z)))
        return out

    roll = GraphedRollout(env, policy, steps=args.episode_length, reset_every=args.episode_length)
    roll.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.episodes):
        obs, ret = roll.run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = args.episodes * args.episode_length
    print("%s, %d worlds: %d episodes x %d steps in %.3f s = %.3g env-steps/s including %d MLP policies "
          "(%.1f us per step); mean episode return of agent 0: %.3f"
          % (args.scenario, args.num_envs, args.episodes, args.episode_length, dt, steps * args.num_envs / dt, env.n,
             1e6 * dt / steps, float(ret[0].mean())))


if __name__ == "__main__":
    main()
