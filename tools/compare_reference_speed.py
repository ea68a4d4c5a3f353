#!/usr/bin/env python
"""Build-container only (needs /root/reference): times the UNMODIFIED Python reference next to oracle/np_port.py, the
per-world NumPy port that bench.py uses as the CPU arm on the GPU box, to show that the port is a faithful stand-in for
"the reference's own NumPy path" in speed as well as in results.  One world per process, all cores, softmax actions."""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(v, "1")


def ref_worker(args):
    name, seed, warmup, steps = args
    import numpy as np
    import refshim
    np.random.seed(seed)
    env = refshim.make_reference_env(name)
    env.reset()
    dims = [int(s.n) if hasattr(s, "n") else int(sum(s.high - s.low + 1)) for s in env.action_space]
    rng = np.random.RandomState(seed)

    def acts():
        out = []
        for d in dims:
            z = rng.randn(5)
            e = np.exp(z - z.max())
            out.append(np.concatenate([e / e.sum(), rng.uniform(0, 1, d - 5)]))
        return out

    for _ in range(warmup):
        env.step(acts())
    t0 = time.perf_counter()
    for t in range(steps):
        if t % 25 == 0:
            env.reset()
        env.step(acts())
    return steps / (time.perf_counter() - t0)


def main():
    import np_port
    from multiagent_particle_envs_b200 import make_env
    procs = len(os.sched_getaffinity(0))
    out = {"cores": procs, "scenarios": {}}
    for name in ("simple", "simple_spread", "simple_tag", "simple_world_comm"):
        with mp.get_context("fork").Pool(procs) as pool:
            ref = sum(pool.map(ref_worker, [(name, 100 + p, 100, 2000) for p in range(procs)]))
        port, _ = np_port.timed_throughput(make_env(name).world.descriptor(), procs, 100, 2000,
                                           shared_reward=(name == "simple_spread"))
        out["scenarios"][name] = {"reference_env_steps_per_s": ref, "np_port_env_steps_per_s": port, "port_over_reference": port / ref}
        print("%-18s reference %8.0f  np_port %8.0f env-steps/s on %d processes  (port / reference = %.2f)"
              % (name, ref, port, procs, port / ref))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r1_cpu_reference_vs_port.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
