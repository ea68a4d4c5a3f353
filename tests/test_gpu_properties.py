"""Size-independent properties checked at BASELINE.json's full batch sizes (where running the CPU
oracle on everything would be slow): determinism, shard equivalence (the multi-GPU contract,
SURVEY.md 8(e)), world-permutation equivariance, collaborative-reward structure, reset
distributions and masked reset."""
import pytest

from helpers import make_product_env

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

FULL = [("simple_spread_n3", 65536), ("simple_tag", 262144), ("simple_spread_n6", 131072),
        ("simple_world_comm", 32768)]


def cuda_actions(nw, n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    acts = []
    for d in nw.act_dims:
        p = torch.softmax(2.0 * torch.randn(n, 5, device="cuda", generator=g), dim=1)
        if d > 5:
            p = torch.cat([p, torch.rand(n, d - 5, device="cuda", generator=g)], 1)
        acts.append(p.contiguous())
    return acts


def rollout(env, acts_per_step):
    outs = []
    for acts in acts_per_step:
        obs_n, rew_n, done_n, _ = env.step(acts)
        outs.append(([o.clone() for o in obs_n], [r.clone() for r in rew_n]))
    return outs


@pytest.mark.parametrize("tag,n", FULL)
def test_determinism_and_shard_equivalence(tag, n):
    """stepping the full batch == stepping two half batches with world offsets, bit for bit;
    and repeating the same rollout reproduces it exactly"""
    steps = 5
    full = make_product_env(tag, num_envs=n, seed=11)
    full.reset()
    acts = [cuda_actions(full.world.native, n, 100 + t) for t in range(steps)]
    ref = rollout(full, acts)
    again = make_product_env(tag, num_envs=n, seed=11)
    again.reset()
    rep = rollout(again, acts)
    for (o1, r1), (o2, r2) in zip(ref, rep):
        assert all(torch.equal(a, b) for a, b in zip(o1 + r1, o2 + r2))
    half = n // 2
    for rank in range(2):
        sh = make_product_env(tag, num_envs=n, seed=11, rank=rank, world_size=2)
        assert sh.world.batch_size == half and sh.world.world_offset == rank * half
        sh.reset()
        lo, hi = rank * half, (rank + 1) * half
        out = rollout(sh, [[a[lo:hi].contiguous() for a in step_acts] for step_acts in acts])
        for (o1, r1), (o2, r2) in zip(ref, out):
            assert all(torch.equal(a[lo:hi], b) for a, b in zip(o1 + r1, o2 + r2))
        assert torch.equal(full.world.native.agent_pv[:, lo:hi], sh.world.native.agent_pv)


@pytest.mark.parametrize("tag,n", FULL[:2])
def test_world_permutation_equivariance(tag, n):
    """worlds are independent: permuting the batch permutes the outputs, bit for bit"""
    a = make_product_env(tag, num_envs=n, seed=5)
    b = make_product_env(tag, num_envs=n, seed=5)
    a.reset()
    b.reset()
    na, nb = a.world.native, b.world.native
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    nb.agent_pv.copy_(na.agent_pv[:, perm])
    nb.lm_p.copy_(na.lm_p[:, perm])
    acts = cuda_actions(na, n, 9)
    oa, ra, _, _ = a.step(acts)
    ob, rb, _, _ = b.step([x[perm].contiguous() for x in acts])
    for x, y in zip(oa + ra, ob + rb):
        assert torch.equal(x[perm], y)


def test_spread_reward_structure_full_size():
    """collaborative world: every agent receives the same sum (environment.py:100-102); the sum is
    <= -A*... and contains the reference's self-collision constant (-1 per agent)"""
    n = 65536
    env = make_product_env("simple_spread_n3", num_envs=n, seed=2)
    env.reset()
    obs_n, rew_n, done_n, info_n = env.step(cuda_actions(env.world.native, n, 3))
    assert torch.equal(rew_n[0], rew_n[1]) and torch.equal(rew_n[1], rew_n[2])
    rew_i, coll, min_d, occ = info_n["n"][0]
    assert float(coll.min()) >= 1.0                      # a == i is counted (simple_spread.py:79-81)
    assert bool((rew_n[0] <= -3.0 + 1e-6).all())
    # shared reward == sum of the per-agent benchmark rewards
    total = sum(info_n["n"][i][0] for i in range(3))
    assert torch.allclose(total, rew_n[0], rtol=1e-6, atol=1e-5)
    assert not any(bool(d.any()) for d in done_n)
    # observation layout: own vel, own pos, then relative positions; trailing comm block is zero
    o = obs_n[0]
    nw = env.world.native
    assert torch.equal(o[:, 0:2], nw.agent_pv[0, :, 2:4]) and torch.equal(o[:, 2:4], nw.agent_pv[0, :, 0:2])
    assert torch.equal(o[:, 4:6], nw.lm_p[0] - nw.agent_pv[0, :, 0:2])
    assert float(o[:, 14:18].abs().max()) == 0.0


def test_tag_speed_limit_and_bounds_full_size():
    n = 262144
    env = make_product_env("simple_tag", num_envs=n, seed=4)
    env.reset()
    nw = env.world.native
    for t in range(30):
        env.reuse_buffers = True
        env.step(cuda_actions(nw, n, 50 + t // 10))      # persistent directions -> speed builds up
    speed = nw.agent_pv[:, :, 2:4].norm(dim=2)
    assert float(speed[:3].max()) <= 1.0 * (1 + 1e-6)    # max_speed of adversaries (simple_tag.py:25)
    assert float(speed[3].max()) <= 1.3 * (1 + 1e-6)
    assert float(speed[3].max()) > 1.25                  # the clamp is actually reached


def test_reset_distribution_mask_and_sharding():
    n = 200000
    env = make_product_env("simple_tag", num_envs=n, seed=123)
    obs_n = env.reset()
    nw = env.world.native
    pos = nw.agent_pv[:, :, 0:2]
    assert float(pos.min()) >= -1.0 and float(pos.max()) < 1.0
    assert abs(float(pos.mean())) < 5e-3 and abs(float(pos.var()) - 1.0 / 3.0) < 5e-3
    assert float(nw.agent_pv[:, :, 2:4].abs().max()) == 0.0
    lm = nw.lm_p
    assert float(lm.min()) >= -0.9 and float(lm.max()) < 0.9          # simple_tag.py:53
    assert abs(float(lm.var()) - 0.27) < 5e-3
    # no two entities share a draw
    flat = torch.cat([pos.reshape(-1, n, 2), lm], 0)[:, :1000].reshape(-1)
    assert flat.unique().numel() > 0.995 * flat.numel()      # 24-bit uniforms: a few birthday collisions
    # masked reset touches only the masked worlds, and successive resets differ
    before = nw.agent_pv.clone()
    mask = torch.zeros(n, dtype=torch.bool, device="cuda")
    mask[::3] = True
    env.step([torch.full((n, 5), 0.2, device="cuda") for _ in range(4)])
    moved = nw.agent_pv.clone()
    env.reset(mask=mask)
    assert torch.equal(nw.agent_pv[:, ~mask], moved[:, ~mask])
    assert not torch.equal(nw.agent_pv[:, mask], before[:, mask])
    assert float(nw.agent_pv[:, mask][:, :, 2:4].abs().max()) == 0.0
    # sharding: rank r of 4 draws exactly the global stream's slice
    g = make_product_env("simple_tag", num_envs=4096, seed=77)
    g.reset()
    for r in range(4):
        s = make_product_env("simple_tag", num_envs=4096, seed=77, rank=r, world_size=4)
        s.reset()
        assert torch.equal(s.world.native.agent_pv, g.world.native.agent_pv[:, r * 1024:(r + 1) * 1024])
        assert torch.equal(s.world.native.lm_p, g.world.native.lm_p[:, r * 1024:(r + 1) * 1024])
