#!/usr/bin/env python
"""Drop-in usage example: the loop an existing MADDPG-style trainer runs, unchanged except for `num_envs`.

    python examples/random_rollout.py --scenario simple_tag --num-envs 65536 --steps 100

Uses the reference's import paths (`make_env`, `multiagent.*`); actions are softmax-random policies.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from make_env import make_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="simple_spread")
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--episode", type=int, default=25)
    args = ap.parse_args()

    env = make_env(args.scenario, num_envs=args.num_envs)
    print("scenario %s: n=%d agents, action spaces %s, observation shapes %s"
          % (args.scenario, env.n, env.action_space, [s.shape for s in env.observation_space]))
    act_dims = env.world.native_shapes().act_dims
    obs_n = env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ret = torch.zeros(env.n, args.num_envs, device=obs_n[0].device)
    for t in range(args.steps):
        action_n = []
        for d in act_dims:   # 5 movement probabilities (if the agent moves) followed by the utterance
            a = torch.rand(args.num_envs, d, device=obs_n[0].device)
            if d >= 5:
                a[:, :5] = torch.softmax(3 * a[:, :5], dim=1)
            action_n.append(a)
        obs_n, rew_n, done_n, info_n = env.step(action_n)
        ret += torch.stack(list(rew_n))
        if (t + 1) % args.episode == 0:
            obs_n = env.reset()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%d steps x %d worlds in %.3f s: %.3g env-steps/s (incl. the random policy); mean return per agent %s"
          % (args.steps, args.num_envs, dt, args.steps * args.num_envs / dt, [round(float(r.mean()), 3) for r in ret]))


if __name__ == "__main__":
    main()
