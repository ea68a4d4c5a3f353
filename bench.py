#!/usr/bin/env python
"""bench.py -- env-steps/sec of the multiagent-particle-envs hot path on B200.

    python bench.py --gpus N --steps K --warmup W              # this repo (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference path's NumPy port

A "step" is one pass of the hot path (MultiAgentEnv.step: action decode -> World.step -> observation /
reward / done, environment.py:80-104) over ONE batch of `n_env` worlds = one fused-kernel launch.
Workload (BASELINE.json configs[1]): simple_spread N=3, 65 536 worlds per GPU.  Under torchrun every
rank steps its own 65 536-world shard (weak scaling, no data-path collective); the only exchange is one
all-gather of the (env_steps, seconds) counters.

Timing hygiene: the timed steps rotate over a ring of R independent batches whose combined working set
is > 2x the 126 MB L2 ("inputs larger than L2"); W >= 3 warm-up steps; CUDA events on the launching
stream with a barrier + synchronize on both sides; max over ranks; nvidia-smi clocks sampled during the
timed region.  Episodes are reset every 25 steps of each batch (MADDPG's episode length).

value      device-resident throughput: inputs already in HBM, K fused launches replayed from CUDA graphs
e2e        the same metric through the public API `env.step(host actions)`: pinned H2D of the actions,
           the fused step, D2H of observations / rewards / dones, every step (mpe_step_host)
roofline   achieved = algorithmic bytes per launch (411 B x n_env) / mean launch time
cpu_baseline  the reference path's per-world NumPy port (oracle/np_port.py), one world per process on all host
              cores; the C oracle (oracle/mpe_oracle.c) on the same cores is reported under cpu_baseline.c_oracle
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the CPU arms run one world per process: keep NumPy / torch thread pools from oversubscribing the host
for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENARIO = "simple_spread"
N_ENV = 65536
EPISODE = 25
L2_BYTES = 126 * 1024 * 1024
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


SCENARIO_KW = {}
BYTES_PER_ENV_STEP = 411
N_AGENTS = 3


def workload_config(n_gpus, ring):
    headline = (SCENARIO == "simple_spread" and N_ENV == 65536 and not SCENARIO_KW)
    return {"workload": ("simple_spread N=3 agents/landmarks, batch=65536 worlds per GPU (BASELINE configs[1])" if headline
                         else "%s %s, batch=%d worlds per GPU" % (SCENARIO, SCENARIO_KW or "", N_ENV)),
            "scenario": SCENARIO, "n_env_per_gpu": N_ENV, "global_n_env": N_ENV * n_gpus,
            "agents": N_AGENTS, "episode_length": EPISODE, "ring_batches": ring,
            "l2_policy": ("inputs larger than L2: steps rotate over %d independent batches (%.0f MB > 2x126 MB)"
                          % (ring, ring * N_ENV * BYTES_PER_ENV_STEP / 1e6)) if ring > 1 else "n/a (CPU arm)",
            "actions": "softmax of N(0,1) logits, pre-generated per batch, resident in HBM",
            "parallelism": "dp%d (independent shards, no data-path collective)" % n_gpus}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[1]) for r in rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(rows[0][2]) if rows and len(rows[0]) > 2 else None,
                "power_w_max": max([float(r[3]) for r in rows if len(r) > 3 and r[3].replace(".", "").isdigit()] or [0.0]),
                "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU arm.  Two ports of the reference path live under oracle/ (test + baseline infrastructure):
#   np_port.py     per-world NumPy float64, the reference's own granularity (one world per process) --
#                  the stand-in for "the reference's own NumPy path"; measured next to the real reference in the
#                  build container it is 1.1-1.5x FASTER than it (tools/compare_reference_speed.py,
#                  profiles/r1_cpu_reference_vs_port.json), i.e. a conservative baseline
#   mpe_oracle.c   the C checker; ~300x faster than the reference itself; reported alongside
# ------------------------------------------------------------------------------------------------
def _spread_desc():
    from multiagent_particle_envs_b200 import make_env
    return make_env(SCENARIO, **SCENARIO_KW).world.descriptor()


def _best_process_count(desc):
    """`os.sched_getaffinity` can exceed what the container may really use (CPU quota, SMT): probe a few process
    counts briefly and keep the one with the highest aggregate throughput -- "all the host threads it can use"."""
    import np_port
    cores = len(os.sched_getaffinity(0))
    cands = sorted({max(1, cores // 8), max(1, cores // 4), max(1, cores // 2), cores})
    best = (0.0, cores)
    for p in cands:
        rate, _ = np_port.timed_throughput(desc, p, 10, 150)
        if rate > best[0]:
            best = (rate, p)
    return best[1], best[0]


def cpu_numpy_port(budget_s, steps=None, warmup=100, max_seconds=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_port
    desc = _spread_desc()
    procs, probe = _best_process_count(desc)
    per_proc = probe / procs
    if steps is None:
        steps = int(max(200, min(20000, budget_s * per_proc)))
    if max_seconds is not None:
        steps = int(max(50, min(steps, max_seconds * per_proc)))
    t0 = time.perf_counter()
    total, rates = np_port.timed_throughput(desc, procs, warmup, steps)
    dt = time.perf_counter() - t0
    cpu_model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                cpu_model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"value": total, "unit": UNIT, "cores": procs, "kind": "port", "cpu_model": cpu_model,
            "sample": "%d processes (best of the probed counts; affinity reports %d CPUs) x %d env.step calls of one "
                      "%s world each (oracle/np_port.py: per-world NumPy float64 restatement at the reference's "
                      "granularity, softmax actions, reset every 25 steps); %.1f s wall"
                      % (procs, len(os.sched_getaffinity(0)), steps, SCENARIO, dt),
            "per_process": total / procs, "steps_per_process": steps, "seconds": dt}


def cpu_c_oracle(budget_s, n_sample=N_ENV):
    """env-steps/s of oracle/mpe_oracle.c (fp64) stepping n_sample worlds split over all host threads; the
    step loop runs inside C (ctypes releases the GIL)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import Oracle
    desc = _spread_desc()
    cores = len(os.sched_getaffinity(0))
    rng = np.random.RandomState(0)
    A, L = 3, 3
    per = (n_sample + cores - 1) // cores
    chunks = []
    for c in range(cores):
        m = min(per, n_sample - c * per)
        if m <= 0:
            break
        pv = np.zeros((m, A, 4))
        pv[:, :, :2] = rng.uniform(-1, 1, (m, A, 2))
        logits = rng.randn(m, A, 5)
        act = np.ascontiguousarray((np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)).reshape(m, 15))
        chunks.append((Oracle(desc, "f64"), pv, rng.uniform(-1, 1, (m, L, 2)), np.zeros((m, A, 2)), act))
    pool = ThreadPoolExecutor(len(chunks))

    def run(steps):
        t0 = time.perf_counter()
        list(pool.map(lambda ch: ch[0].rollout(ch[1], ch[2], ch[3], ch[4], steps, 1), chunks))
        return time.perf_counter() - t0

    run(1)
    one = run(2) / 2
    steps = int(max(2, min(5000, budget_s / max(one, 1e-6))))
    dt = run(steps)
    pool.shutdown()
    return {"value": n_sample * steps / dt, "unit": UNIT, "cores": len(chunks), "kind": "port",
            "sample": "%d steps x %d worlds through oracle/mpe_oracle.c (fp64 C restatement, step loop in C, one thread "
                      "per core); %.1f s" % (steps, n_sample, dt)}


def cpu_baseline_block(numpy_seconds, c_seconds):
    cb = cpu_numpy_port(numpy_seconds)
    out = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model")}
    out["per_process"] = cb["per_process"]
    if c_seconds > 0:
        c = cpu_c_oracle(c_seconds)
        out["c_oracle"] = {k: c[k] for k in ("value", "unit", "cores", "sample")}
    return out, cb


def run_reference_arm(args, rank, world):
    """The reference's CPU path on this box's host cores: the per-world NumPy port, one world per process
    on every core.  One bench "step" here = one env.step in each of the `cores` worlds (a bounded sample of
    the 65536-world batch); the timed run is capped at ~90 s."""
    if rank != 0:
        return
    t_all = time.perf_counter()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    res = cpu_numpy_port(0, steps=args.steps, warmup=min(max(args.warmup, 3), 300), max_seconds=75.0)
    steps_timed = res["steps_per_process"]
    cores = res["cores"]
    cb = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model")}
    cb["per_process"] = res["per_process"]
    if SCENARIO == "simple_spread" and not SCENARIO_KW:
        c = cpu_c_oracle(5.0)
        cb["c_oracle"] = {k: c[k] for k in ("value", "unit", "cores", "sample")}
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "steps_timed_per_process": steps_timed, "warmup": args.warmup,
            "ms_per_step": 1e3 * cores / res["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus, 1), "cpu_baseline": cb,
            "agent_steps_per_sec": 3 * res["value"],
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
            "note": "reference arm = oracle/np_port.py, the per-world NumPy port of the reference path (the Python reference "
                    "itself cannot travel to the GPU box); %d processes, one world each; the C oracle on the same cores is "
                    "reported under cpu_baseline.c_oracle" % cores}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from multiagent_particle_envs_b200 import _lib, make_env
    from multiagent_particle_envs_b200.sharding import aggregate_counters

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints a "NCCL version ..." banner on stdout when its communicator is created; rank 0 must print
        # exactly ONE line on stdout, so file descriptor 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    lib = _lib.load()

    # ---- ring of independent batches, each with resident actions and outputs -------------------
    bytes_per_step = None
    ring = []
    R = args.ring
    for b in range(R):
        env = make_env(SCENARIO, num_envs=N_ENV * world, device=dev, seed=1000 + b, rank=rank, world_size=world,
                       **SCENARIO_KW)
        env.reset()
        nw = env.world.native
        bytes_per_step = nw.bytes_per_env_step * N_ENV
        g = torch.Generator(device=dev).manual_seed(7 * b + rank)
        acts = []
        for d_act, ag in zip(nw.act_dims, env.agents):     # 5 movement probabilities, then the utterance
            parts = []
            if ag.movable:
                parts.append(torch.softmax(torch.randn(N_ENV, 5, device=dev, generator=g), 1))
            if d_act - (5 if ag.movable else 0) > 0:
                parts.append(torch.rand(N_ENV, d_act - (5 if ag.movable else 0), device=dev, generator=g))
            acts.append(torch.cat(parts, 1).contiguous())
        ring.append((env, nw, acts, _lib.ptr_array([t.data_ptr() for t in acts]), env._flags()))
    if R * bytes_per_step <= 2 * L2_BYTES:
        raise SystemExit("ring working set (%d x %.0f MB) must exceed 2x L2: raise --ring" % (R, bytes_per_step / 1e6))
    stream = torch.cuda.Stream(dev)
    launches = [0]

    def step_ring_once():
        for env, nw, acts, ptrs, flags in ring:
            nw.step(ptrs, nw.out, flags)
            launches[0] += 1

    def reset_ring():
        for env, nw, acts, ptrs, flags in ring:
            nw.reset()
            launches[0] += 1

    unit_steps = R * EPISODE
    with torch.cuda.stream(stream):
        step_ring_once()  # first launches outside capture (module load)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = launches[0]
        with torch.cuda.graph(graph, stream=stream):
            for _ in range(EPISODE):
                step_ring_once()
        launches[0] = l0  # capture is not execution

        def run_steps(k):
            """exactly k fused steps: whole episodes from the graph (+ the episode resets), then the rest"""
            units, rem = divmod(k, unit_steps)
            for _ in range(units):
                graph.replay()
                launches[0] += unit_steps
                reset_ring()
            i = 0
            while rem > 0:
                env, nw, acts, ptrs, flags = ring[i % R]
                nw.step(ptrs, nw.out, flags)
                launches[0] += 1
                i += 1
                rem -= 1

        run_steps(max(args.warmup, 3))
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank).start()
        time.sleep(0.12)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches[0] = 0
        t0 = time.time()
        e0.record(stream)
        run_steps(args.steps)
        e1.record(stream)
        stream.synchronize()
        t1 = time.time()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop(t0, t1)
        seconds = e0.elapsed_time(e1) / 1e3
        # extra (not the headline): the same K steps with TWO batches in flight -- even ring slots on one stream,
        # odd ones on another, captured as a fork/join graph -- so that a step's launch latency, ramp-up and
        # store drain overlap the neighbouring batch's step.  Each batch still advances strictly in order.
        side = torch.cuda.Stream(dev)
        graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph2, stream=stream):
            side.wait_stream(stream)
            for _ in range(EPISODE):
                for slot, (env_s, nw_s, acts_s, ptrs_s, flags_s) in enumerate(ring):
                    with torch.cuda.stream(side if slot % 2 else stream):
                        nw_s.step(ptrs_s, nw_s.out, flags_s)
            stream.wait_stream(side)
        units2 = max(1, args.steps // unit_steps)
        for _ in range(2):
            graph2.replay()
        stream.synchronize()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(units2):
            graph2.replay()
        e3.record(stream)
        stream.synchronize()
        seconds2 = e2.elapsed_time(e3) / 1e3
    gpu_launches = launches[0]
    total_steps, max_seconds, per_rank = aggregate_counters(N_ENV * args.steps, seconds)
    total2, max2, _ = aggregate_counters(N_ENV * units2 * unit_steps, seconds2)
    value = total_steps / max_seconds

    # ---- end to end through the public API with host buffers ------------------------------------
    env, nw = ring[0][0], ring[0][1]
    k_e2e = max(3, min(args.steps, args.e2e_steps))
    host_acts = [[a.cpu().pin_memory() for a in ring[b % R][2]] for b in range(4)]
    h2d = sum(a.numel() * 4 for a in host_acts[0])
    for b in range(3):
        obs_n, rew_n, done_n, _ = env.step(host_acts[b % 4])
    d2h = sum(o.numel() * 4 for o in obs_n) + sum(r.numel() * 4 for r in rew_n) + sum(d.numel() for d in done_n)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    checksum = 0.0
    w0 = time.perf_counter()
    for b in range(k_e2e):
        obs_n, rew_n, done_n, _ = env.step(host_acts[b % 4])   # H2D + fused step + D2H + sync inside
        checksum += float(rew_n[0][0])                          # the caller reads the result on the host
    torch.cuda.synchronize()
    e2e_seconds = time.perf_counter() - w0
    e2e_total, e2e_max, _ = aggregate_counters(N_ENV * k_e2e, e2e_seconds)
    e2e_value = e2e_total / e2e_max

    # extra (not the headline): two env batches in flight through step_async / step_wait, so that the upload +
    # step of one overlaps the download of the other -- what a double-buffered host trainer would see
    env_b = ring[1][0]
    lanes = {id(env): torch.cuda.Stream(dev), id(env_b): torch.cuda.Stream(dev)}   # one stream per env batch

    def launch(e, acts):
        with torch.cuda.stream(lanes[id(e)]):
            e.step_async(acts)

    for b in range(2):
        launch(env, host_acts[b % 4]); launch(env_b, host_acts[(b + 1) % 4]); env.step_wait(); env_b.step_wait()
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    launch(env, host_acts[0])
    for b in range(k_e2e):
        cur, nxt = (env, env_b) if b % 2 == 0 else (env_b, env)
        if b + 1 < k_e2e:
            launch(nxt, host_acts[(b + 1) % 4])
        obs_n, rew_n, done_n, _ = cur.step_wait()
        checksum += float(rew_n[0][0])
    torch.cuda.synchronize()
    pipe_seconds = time.perf_counter() - w0
    pipe_total, pipe_max, _ = aggregate_counters(N_ENV * k_e2e, pipe_seconds)

    if rank == 0:
        peak, peak_src = measured_peak()
        launch_s = max_seconds / args.steps
        achieved = bytes_per_step / launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and SCENARIO == "simple_spread" and N_ENV == 65536 and not SCENARIO_KW:
            try:
                traffic = json.load(open(tp)).get("simple_spread_65536_dram_bytes_per_launch")
            except Exception:  # noqa: BLE001
                traffic = None
        cpu = None
        if world == 1 and args.cpu_seconds > 0:
            cpu = cpu_baseline_block(args.cpu_seconds, 5.0 if SCENARIO == "simple_spread" and not SCENARIO_KW else 0.0)[0]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * launch_s, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world, R),
            "agent_steps_per_sec": N_AGENTS * value,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": k_e2e, "ms_per_step": 1e3 * e2e_max / k_e2e, "api": "MultiAgentEnv.step(pinned host tensors)"},
            "value_two_batches_in_flight": {"value": total2 / max2, "unit": UNIT, "ms_per_step": 1e3 * max2 / (units2 * unit_steps),
                                            "frac": bytes_per_step / (max2 / (units2 * unit_steps)) / 1e9 / peak,
                                            "note": "extra: alternate ring slots on two streams (fork/join CUDA graph)"},
            "e2e_two_batches_in_flight": {"value": pipe_total / pipe_max, "unit": UNIT, "ms_per_step": 1e3 * pipe_max / k_e2e,
                                          "api": "step_async / step_wait alternating over two env batches"},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_step,
                         "kernel": "mpe_kernel<%s program, kFusedStep>" % SCENARIO},
            "per_rank": per_rank,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60000)
    ap.add_argument("--warmup", type=int, default=3000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ring", type=int, default=12)
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--scenario", default="simple_spread", help="other BASELINE configs: simple_tag, simple_world_comm, ...")
    ap.add_argument("--num-envs", type=int, default=65536, help="worlds per GPU")
    ap.add_argument("--num-agents", type=int, default=None, help="simple_spread only (N agents = N landmarks)")
    args = ap.parse_args()
    global SCENARIO, N_ENV, SCENARIO_KW, BYTES_PER_ENV_STEP, N_AGENTS
    SCENARIO, N_ENV = args.scenario, args.num_envs
    if args.num_agents is not None:
        SCENARIO_KW = {"num_agents": args.num_agents}
    from multiagent_particle_envs_b200 import make_env as _mk
    _probe = _mk(SCENARIO, **SCENARIO_KW)
    BYTES_PER_ENV_STEP, N_AGENTS = _probe.world.native_shapes().bytes_per_env_step, _probe.n
    if args.ring * N_ENV * BYTES_PER_ENV_STEP <= 2 * L2_BYTES:
        args.ring = int(2 * L2_BYTES / (N_ENV * BYTES_PER_ENV_STEP)) + 2
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29400 + os.getpid() % 500), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    import __graft_entry__ as g
    g.build(quiet=True)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
