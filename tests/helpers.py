"""shared test helpers (the oracle is imported HERE, in tests/, only as the checker)"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# golden tag -> (scenario name, scenario kwargs)
CONFIGS = {
    "simple": ("simple", {}),
    "simple_spread_n3": ("simple_spread", {}),
    "simple_spread_n6": ("simple_spread", {"num_agents": 6}),
    "simple_tag": ("simple_tag", {}),
    "simple_world_comm": ("simple_world_comm", {}),
    "simple_adversary": ("simple_adversary", {}),
    "simple_push": ("simple_push", {}),
    "simple_speaker_listener": ("simple_speaker_listener", {}),
    "simple_reference": ("simple_reference", {}),
    "simple_crypto": ("simple_crypto", {}),
}
# entity-count variants the reference hard-codes away (no goldens; checked against the generic oracle)
VARIANTS = {
    "simple_tag_1v1": ("simple_tag", {"num_adversaries": 1, "num_good_agents": 1, "num_landmarks": 2}),
    "simple_tag_2v1": ("simple_tag", {"num_adversaries": 2, "num_good_agents": 1, "num_landmarks": 2}),
    "simple_tag_4v2": ("simple_tag", {"num_adversaries": 4, "num_good_agents": 2, "num_landmarks": 2}),
    "simple_tag_6v2": ("simple_tag", {"num_adversaries": 6, "num_good_agents": 2, "num_landmarks": 3}),
    "simple_adversary_n4": ("simple_adversary", {"num_agents": 4}),
}
NO_BENCHMARK = ("simple", "simple_push", "simple_speaker_listener", "simple_reference")


def load_golden(tag):
    return dict(np.load(os.path.join(GOLDEN, tag + ".npz")))


def make_product_env(tag, **kw):
    from multiagent_particle_envs_b200 import make_env
    name, skw = CONFIGS[tag] if tag in CONFIGS else VARIANTS[tag]
    kw.update(skw)
    return make_env(name, benchmark=(name not in NO_BENCHMARK), **kw)


def descriptor(tag):
    from multiagent_particle_envs_b200 import scenarios
    name, kw = CONFIGS[tag] if tag in CONFIGS else VARIANTS[tag]
    return scenarios.load(name).Scenario(**kw).make_world().descriptor()


def step_flags(tag_or_golden):
    from multiagent_particle_envs_b200 import _lib
    g = load_golden(tag_or_golden) if isinstance(tag_or_golden, str) else tag_or_golden
    f = 0
    if int(g["prop_shared_reward"]):
        f |= _lib.FLAG_SHARED_REWARD
    if int(g["force_discrete"]):
        f |= _lib.FLAG_FORCE_DISCRETE_ACTION
    return f


def random_states(desc, n, rng, mode="mixed"):
    """seeded synthetic worlds in the oracle layout: reset-like, squeezed (contacts), fast/outside"""
    A, L, C = desc.n_agents, desc.n_landmarks, desc.dim_c
    pv = np.zeros((n, A, 4))
    pv[:, :, 0:2] = rng.uniform(-1, 1, (n, A, 2))
    lm = rng.uniform(-0.9, 0.9, (n, L, 2))
    kind = rng.randint(0, 4, n) if mode == "mixed" else np.zeros(n, int)
    sq = kind == 1
    pv[sq, :, 0:2] *= 0.3
    lm[sq] *= 0.3
    fast = kind == 2
    pv[fast, :, 0:2] *= 1.25
    pv[fast, :, 2:4] = rng.uniform(-1.5, 1.5, (int(fast.sum()), A, 2))
    tight = kind == 3
    pv[tight, :, 0:2] = rng.uniform(-0.12, 0.12, (int(tight.sum()), A, 2))
    comm = np.zeros((n, A, C))
    for i in range(A):
        if not desc.agent_silent[i]:
            comm[:, i, :] = rng.uniform(0, 1, (n, C)) * (rng.uniform(0, 1, (n, 1)) > 0.1)
        if not desc.agent_movable[i]:
            pv[:, i, 2:4] = 0.0
    return pv, lm, comm


def random_goals(n_goals, n_landmarks, n, rng):
    return rng.randint(0, max(n_landmarks, 1), (n, n_goals)).astype(np.int32)


def random_actions(act_dims, n, rng, temperature=2.0, movable=None):
    """probability vectors as MADDPG emits (softmax of logits) + uniform comm"""
    parts = []
    for i, d in enumerate(act_dims):
        mov = True if movable is None else bool(movable[i])
        if mov:
            logits = temperature * rng.randn(n, 5)
            p = np.exp(logits - logits.max(axis=1, keepdims=True))
            p /= p.sum(axis=1, keepdims=True)
            parts.append(p)
            d -= 5
        if d > 0:
            parts.append(rng.uniform(0, 1, (n, d)) * (rng.uniform(0, 1, (n, 1)) > 0.1))
    return np.concatenate(parts, axis=1)


def split_cols(a, dims):
    out, c = [], 0
    for d in dims:
        out.append(a[..., c:c + d])
        c += d
    return out
