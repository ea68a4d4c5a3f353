"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.

A NumPy float64, ONE-WORLD-AT-A-TIME restatement of the reference hot path, kept deliberately at the
reference's own granularity (tiny ndarrays per entity, Python loops over entities and pairs): it is the
stand-in for "the reference's own NumPy path" that can travel to the GPU box, where /root/reference
does not exist.  bench.py times it (one world per process, all host cores) for `cpu_baseline` and for
`--impl reference`; tests/test_oracle_golden.py pins it against the reference's golden fixtures.
The C oracle (mpe_oracle.c) is the fast checker; this file is the honest speed baseline.

Each function cites the reference file:line it follows.  State of one world:
    pos [E,2] (agents then landmarks), vel [A,2], comm [A,C]
"""
import numpy as np


class WorldSpec(object):
    """plain-Python copy of an mpe_desc (include/mpe_b200.h)"""

    def __init__(self, d):
        self.scenario = int(d.scenario)
        self.A, self.L, self.C = int(d.n_agents), int(d.n_landmarks), int(d.dim_c)
        self.dt, self.damping = float(d.dt), float(d.damping)
        self.contact_force, self.contact_margin = float(d.contact_force), float(d.contact_margin)
        A, L = self.A, self.L
        self.size = [float(d.agent_size[i]) for i in range(A)] + [float(d.landmark_size[l]) for l in range(L)]
        self.mass = [float(d.agent_mass[i]) for i in range(A)]
        self.sens = [float(d.agent_sens[i]) for i in range(A)]
        self.max_speed = [None if d.agent_max_speed[i] < 0 else float(d.agent_max_speed[i]) for i in range(A)]
        self.movable = [bool(d.agent_movable[i]) for i in range(A)] + [False] * L
        self.collide = [bool(d.agent_collide[i]) for i in range(A)] + [bool(d.landmark_collide[l]) for l in range(L)]
        self.silent = [bool(d.agent_silent[i]) for i in range(A)]
        self.adversary = [bool(d.agent_adversary[i]) for i in range(A)]
        self.leader = [bool(d.agent_leader[i]) for i in range(A)]
        self.n_obstacles, self.n_food = int(d.n_obstacles), int(d.n_food)
        # landmark colours the observation functions embed (simple_push.py:34-37, simple_speaker_listener.py:44-46,
        # simple_reference.py:37-39, simple_crypto.py:58-62)
        self.push_lm_color = [np.array([0.1 + (0.8 if c == l + 1 else 0.0) for c in range(3)]) for l in range(L)]
        self.sl_lm_color = [np.array([0.65 if c == l else 0.15 for c in range(3)]) for l in range(L)]
        self.ref_lm_color = [np.array([0.75 if c == l else 0.25 for c in range(3)]) for l in range(L)]
        self.crypto_color = [np.eye(max(self.C, 1))[l] for l in range(L)] if self.C >= L else []


def decode_actions(spec, action_n, force_discrete=False):
    """MultiAgentEnv._set_action (environment.py:144-192) -> (u [A,2], c [A,C])"""
    u = np.zeros((spec.A, 2))
    c = np.zeros((spec.A, spec.C))
    for i in range(spec.A):
        a = np.asarray(action_n[i], dtype=np.float64)
        k = 0
        if spec.movable[i]:
            p = a[0:5].copy()
            if force_discrete:                                  # :169-172
                d = np.argmax(p)
                p[:] = 0.0
                p[d] = 1.0
            u[i, 0] += p[1] - p[2]                              # :174
            u[i, 1] += p[3] - p[4]                              # :175
            u[i] *= spec.sens[i]                                # :178-181
            k = 5
        if not spec.silent[i]:
            c[i] = a[k:k + spec.C]                              # :183-190
    return u, c


def world_step(spec, pos, vel, comm, u, c):
    """World.step (core.py:117-131), in place"""
    E, A = spec.A + spec.L, spec.A
    force = [None] * E
    for i in range(A):                                          # apply_action_force :134-140
        if spec.movable[i]:
            force[i] = u[i] + 0.0
    for a in range(E):                                          # apply_environment_force :143-155
        for b in range(a + 1, E):
            if not (spec.collide[a] and spec.collide[b]):       # get_collision_force :181-182
                continue
            delta = pos[a] - pos[b]                             # :186
            dist = np.sqrt(np.sum(np.square(delta)))            # :187
            dist_min = spec.size[a] + spec.size[b]              # :189
            k = spec.contact_margin
            pen = np.logaddexp(0, -(dist - dist_min) / k) * k   # :192
            f = spec.contact_force * delta / dist * pen         # :193
            if spec.movable[a]:
                force[a] = f + (force[a] if force[a] is not None else 0.0)
            if spec.movable[b]:
                force[b] = -f + (force[b] if force[b] is not None else 0.0)
    for i in range(A):                                          # integrate_state :158-169
        if not spec.movable[i]:
            continue
        vel[i] = vel[i] * (1 - spec.damping)
        if force[i] is not None:
            vel[i] += (force[i] / spec.mass[i]) * spec.dt
        if spec.max_speed[i] is not None:
            speed = np.sqrt(np.square(vel[i][0]) + np.square(vel[i][1]))
            if speed > spec.max_speed[i]:
                vel[i] = vel[i] / speed * spec.max_speed[i]
        pos[i] += vel[i] * spec.dt
    for i in range(A):                                          # update_agent_state :171-177
        comm[i] = 0.0 if spec.silent[i] else c[i]


def _dist(p, q):
    return np.sqrt(np.sum(np.square(p - q)))


def _hit(spec, pos, a, b):
    return _dist(pos[a], pos[b]) < spec.size[a] + spec.size[b]   # is_collision


def _bound(x):                                                   # simple_tag.py:103-108
    if x < 0.9:
        return 0
    if x < 1.0:
        return (x - 0.9) * 10
    return min(np.exp(2 * x - 2), 10)


def observe(spec, pos, vel, comm, shared_reward=False, goal=None):
    """scenario.observation / reward for every agent + the step glue (environment.py:92-102); `goal` holds the
    per-world goal indices reset_world draws with np.random.choice (adversary, push, speaker_listener: [g];
    reference: [goal_b of agent 0, goal_b of agent 1]; crypto: [goal, key])"""
    A, L = spec.A, spec.L
    lm = pos[A:]
    obs, rew = [], []
    if spec.scenario == 0:                                       # simple.py:41-50
        for i in range(A):
            obs.append(np.concatenate([vel[i]] + [lm[l] - pos[i] for l in range(L)]))
            rew.append(-np.sum(np.square(pos[i] - lm[0])))
    elif spec.scenario == 1:                                     # simple_spread.py:72-100
        for i in range(A):
            r = 0
            for l in range(L):
                r -= min(_dist(pos[a], lm[l]) for a in range(A))
            if spec.collide[i]:
                for a in range(A):
                    if _hit(spec, pos, a, i):
                        r -= 1
            rew.append(r)
            others = [j for j in range(A) if j != i]
            obs.append(np.concatenate([vel[i], pos[i]] + [lm[l] - pos[i] for l in range(L)] +
                                      [pos[j] - pos[i] for j in others] + [comm[j] for j in others]))
    elif spec.scenario == 2:                                     # simple_tag.py:84-147
        adv = [i for i in range(A) if spec.adversary[i]]
        good = [i for i in range(A) if not spec.adversary[i]]
        for i in range(A):
            r = 0
            if spec.adversary[i]:
                if spec.collide[i]:
                    for g in good:
                        for a in adv:
                            if _hit(spec, pos, g, a):
                                r += 10
            else:
                if spec.collide[i]:
                    for a in adv:
                        if _hit(spec, pos, a, i):
                            r -= 10
                for p in range(2):
                    r -= _bound(abs(pos[i][p]))
            rew.append(r)
            others = [j for j in range(A) if j != i]
            obs.append(np.concatenate([vel[i], pos[i]] + [lm[l] - pos[i] for l in range(L)] +
                                      [pos[j] - pos[i] for j in others] +
                                      [vel[j] for j in others if not spec.adversary[j]]))
    elif spec.scenario == 3:                                     # simple_world_comm.py:142-287
        adv = [i for i in range(A) if spec.adversary[i]]
        good = [i for i in range(A) if not spec.adversary[i]]
        food = [A + spec.n_obstacles + f for f in range(spec.n_food)]
        forest = [A + spec.n_obstacles + spec.n_food + f for f in range(2)]
        leader = [i for i in range(A) if spec.leader[i]][0]
        for i in range(A):
            r = 0
            if spec.adversary[i]:
                r -= 0.1 * min(_dist(pos[g], pos[i]) for g in good)
                if spec.collide[i]:
                    for g in good:
                        for a in adv:
                            if _hit(spec, pos, g, a):
                                r += 5
            else:
                if spec.collide[i]:
                    for a in adv:
                        if _hit(spec, pos, a, i):
                            r -= 5
                for p in range(2):
                    r -= 2 * _bound(abs(pos[i][p]))
                for f in food:
                    if _hit(spec, pos, i, f):
                        r += 2
                r += 0.05 * min(_dist(pos[f], pos[i]) for f in food)
            rew.append(r)
            inf = [_hit(spec, pos, i, forest[0]), _hit(spec, pos, i, forest[1])]
            other_pos, other_vel = [], []
            for j in range(A):
                if j == i:
                    continue
                of = [_hit(spec, pos, j, forest[0]), _hit(spec, pos, j, forest[1])]
                vis = (inf[0] and of[0]) or (inf[1] and of[1]) or (not inf[0] and not of[0] and not inf[1] and not of[1]) \
                    or spec.leader[i]
                other_pos.append(pos[j] - pos[i] if vis else np.zeros(2))
                if not spec.adversary[j]:
                    other_vel.append(vel[j] if vis else np.zeros(2))
            in_forest = [np.array([1.0 if inf[0] else -1.0]), np.array([1.0 if inf[1] else -1.0])]
            ent = [lm[l] - pos[i] for l in range(L)]
            if spec.adversary[i]:
                obs.append(np.concatenate([vel[i], pos[i]] + ent + other_pos + other_vel + in_forest + [comm[leader]]))
            else:
                obs.append(np.concatenate([vel[i], pos[i]] + ent + other_pos + in_forest + other_vel))
    elif spec.scenario == 4:                                     # simple_adversary.py:76-139
        g = lm[goal[0]]
        adv = [i for i in range(A) if spec.adversary[i]]
        good = [i for i in range(A) if not spec.adversary[i]]
        for i in range(A):
            if spec.adversary[i]:
                rew.append(-np.sum(np.square(pos[i] - g)))                                   # :107-118
            else:
                adv_rew = sum(_dist(pos[a], g) for a in adv)                                 # :83
                pos_rew = -min(_dist(pos[a], g) for a in good)                               # :93-94
                rew.append(pos_rew + adv_rew)
            ent = [lm[l] - pos[i] for l in range(L)]
            other = [pos[j] - pos[i] for j in range(A) if j != i]
            obs.append(np.concatenate(ent + other) if spec.adversary[i] else np.concatenate([g - pos[i]] + ent + other))
    elif spec.scenario == 5:                                     # simple_push.py:58-96
        g = lm[goal[0]]
        good = [i for i in range(A) if not spec.adversary[i]]
        for i in range(A):
            if spec.adversary[i]:
                rew.append(min(_dist(pos[a], g) for a in good) - _dist(g, pos[i]))           # :66-74
            else:
                rew.append(-_dist(pos[i], g))                                                # :62-64
            ent = [lm[l] - pos[i] for l in range(L)]
            other = [pos[j] - pos[i] for j in range(A) if j != i]
            if spec.adversary[i]:
                obs.append(np.concatenate([vel[i]] + ent + other))
            else:
                color = np.array([0.25, 0.25, 0.25])
                color[goal[0] + 1] += 0.5                                                    # :47-53
                obs.append(np.concatenate([vel[i], g - pos[i], color] + ent + spec.push_lm_color + other))
    elif spec.scenario == 6:                                     # simple_speaker_listener.py:63-92
        r = -np.sum(np.square(pos[1] - lm[goal[0]]))
        rew = [r, r]
        obs.append(np.concatenate([spec.sl_lm_color[goal[0]]]))                              # speaker :87-88
        obs.append(np.concatenate([vel[1]] + [lm[l] - pos[1] for l in range(L)] + [comm[0]]))   # listener :90-92
    elif spec.scenario == 7:                                     # simple_reference.py:55-80
        for i in range(A):
            o = 1 - i
            rew.append(-np.sum(np.square(pos[o] - lm[goal[i]])))
            obs.append(np.concatenate([vel[i]] + [lm[l] - pos[i] for l in range(L)] + [spec.ref_lm_color[goal[i]], comm[o]]))
    elif spec.scenario == 8:                                     # simple_crypto.py:94-174
        gcol, key = spec.crypto_color[goal[0]], spec.crypto_color[goal[1]]
        zero = np.zeros(spec.C)
        for i in range(A):
            if spec.adversary[i]:                                                            # :115-121
                r = 0
                if not (comm[i] == zero).all():
                    r -= np.sum(np.square(comm[i] - gcol))
            else:                                                                            # :94-113
                good_rew, adv_rew = 0, 0
                for a in range(A):
                    if not spec.adversary[a] and a != 2 and not (comm[a] == zero).all():
                        good_rew -= np.sum(np.square(comm[a] - gcol))
                    if spec.adversary[a] and not (comm[a] == zero).all():
                        adv_rew += np.sum(np.square(comm[a] - gcol))
                r = adv_rew + good_rew
            rew.append(r)
            if i == 2:
                obs.append(np.concatenate([gcol, key]))
            elif not spec.adversary[i]:
                obs.append(np.concatenate([key, comm[2]]))
            else:
                obs.append(np.concatenate([comm[2]]))
    else:
        raise NotImplementedError("scenario %d" % spec.scenario)
    done = [False] * A                                           # environment.py:132-135
    if shared_reward:                                            # environment.py:100-102
        rew = [np.sum(rew)] * A
    return obs, rew, done


def env_step(spec, pos, vel, comm, action_n, shared_reward=False, force_discrete=False, goal=None):
    """MultiAgentEnv.step (environment.py:80-104) for one world"""
    u, c = decode_actions(spec, action_n, force_discrete)
    world_step(spec, pos, vel, comm, u, c)
    return observe(spec, pos, vel, comm, shared_reward, goal)


# ---- timing helper used by bench.py ---------------------------------------------------------------
def _worker(args):
    import time
    desc_bytes, seed, warmup, steps, shared = args
    import ctypes
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from multiagent_particle_envs_b200._lib import MpeDesc
    d = MpeDesc.from_buffer_copy(desc_bytes)
    spec = WorldSpec(d)
    rng = np.random.RandomState(seed)
    pos = np.concatenate([rng.uniform(-1, 1, (spec.A, 2)), rng.uniform(-1, 1, (spec.L, 2))])
    vel = np.zeros((spec.A, 2))
    comm = np.zeros((spec.A, spec.C))
    adims = [(5 if spec.movable[i] else 0) + (0 if spec.silent[i] else spec.C) for i in range(spec.A)]
    goal = [int(rng.randint(0, max(spec.L, 1))), int(rng.randint(0, max(spec.L, 1)))]

    def acts():
        out = []
        for i, dmn in enumerate(adims):
            if spec.movable[i]:
                z = rng.randn(5)
                e = np.exp(z - z.max())
                out.append(np.concatenate([e / e.sum(), rng.uniform(0, 1, dmn - 5)]))
            else:
                out.append(rng.uniform(0, 1, dmn))
        return out

    for t in range(warmup):
        env_step(spec, pos, vel, comm, acts(), shared, goal=goal)
    t0 = time.perf_counter()
    for t in range(steps):
        if t % 25 == 0:
            pos[:spec.A] = rng.uniform(-1, 1, (spec.A, 2))
            vel[:] = 0
        env_step(spec, pos, vel, comm, acts(), shared, goal=goal)
    return steps / (time.perf_counter() - t0)


def timed_throughput(desc, procs, warmup, steps, shared_reward=True):
    """env-steps/s of `procs` processes each stepping one world `steps` times (aggregate, per-process list)"""
    import ctypes
    import multiprocessing as mp
    raw = bytes(ctypes.string_at(ctypes.addressof(desc), ctypes.sizeof(desc)))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        rates = pool.map(_worker, [(raw, 1000 + p, warmup, steps, shared_reward) for p in range(procs)])
    return float(sum(rates)), [float(r) for r in rates]
