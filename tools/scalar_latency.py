#!/usr/bin/env python
"""Latency of the scalar (batch-1, NumPy) convention -- BASELINE config 1: `simple`, one world.
This mode exists for plumbing / bit-match, not for speed: every step is a pinned H2D copy, one kernel launch,
a D2H copy and a stream synchronisation for a single world."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from make_env import make_env  # noqa: E402

for name in ("simple", "simple_spread", "simple_tag", "simple_world_comm"):
    env = make_env(name)
    env.reset()
    dims = env.world.native_shapes().act_dims
    rng = np.random.RandomState(0)
    acts = [[rng.dirichlet(np.ones(d)) for d in dims] for _ in range(64)]
    for t in range(200):
        env.step(acts[t % 64])
    t0 = time.perf_counter()
    n = 3000
    for t in range(n):
        env.step(acts[t % 64])
    dt = time.perf_counter() - t0
    print("%-18s scalar mode: %.1f us per env.step (%.0f env-steps/s)" % (name, 1e6 * dt / n, n / dt))
