#!/usr/bin/env python
"""Host-side cost of one batched `env.step(cuda tensors)` call (the kernel itself takes ~6.5 us at 65536 worlds):
fresh output tensors per step (default, reference-like ownership) vs env.reuse_buffers = True."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from make_env import make_env  # noqa: E402

n = 65536
for name in ("simple_spread", "simple_world_comm"):
    for reuse in (False, True):
        env = make_env(name, num_envs=n)
        env.reuse_buffers = reuse
        env.reset()
        acts = [torch.rand(n, d, device="cuda") for d in env.world.native_shapes().act_dims]
        for _ in range(200):
            env.step(acts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        k = 5000
        for _ in range(k):
            obs_n, rew_n, done_n, info_n = env.step(acts)
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        print("%-18s reuse_buffers=%-5s  %.1f us per env.step call (host), %.1f us per step incl. GPU drain"
              % (name, reuse, 1e6 * t_issue / k, 1e6 * t_all / k))
