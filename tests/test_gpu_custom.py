"""User-defined scenarios (TorchScenario): `_set_action` + `World.step` on the generic native program, the
scenario's observation / reward written with torch ops over the batched state.

(1) simple_spread re-expressed as a user scenario must reproduce the compiled Spread<3> program: state bit for
    bit (same primitives, same pair order), observations exactly, rewards to rounding;
(2) an entity table no built-in scenario has (immovable speaker, speed limits on some agents, colliding and
    non-colliding agents and landmarks, unequal masses) against the CPU oracle's generic _set_action / World.step.
"""
import numpy as np
import pytest

from helpers import make_product_env, random_actions, split_cols

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _spread_as_user_scenario():
    from multiagent_particle_envs_b200 import Agent, Landmark, TorchScenario, World

    class Scenario(TorchScenario):
        def make_world(self, num_envs=None, device=None):
            world = World()
            world.dim_c = 2
            world.collaborative = True
            world.agents = [Agent() for _ in range(3)]
            for i, agent in enumerate(world.agents):
                agent.name, agent.collide, agent.silent, agent.size = 'agent %d' % i, True, True, 0.15
            world.landmarks = [Landmark() for _ in range(3)]
            for i, lm in enumerate(world.landmarks):
                lm.name, lm.collide, lm.movable = 'landmark %d' % i, False, False
            return self._finish_world(world, num_envs, device)

        def reward(self, agent, world):                      # simple_spread.py:72-82 in torch, vectorised over worlds
            rew = 0
            for l in world.landmarks:
                dists = torch.stack([(a.state.p_pos - l.state.p_pos).square().sum(1).sqrt() for a in world.agents])
                rew = rew - dists.min(0).values
            for a in world.agents:
                d = (a.state.p_pos - agent.state.p_pos).square().sum(1).sqrt()
                rew = rew - (d < a.size + agent.size).float()
            return rew

        def observation(self, agent, world):                 # simple_spread.py:84-100
            ent = [l.state.p_pos - agent.state.p_pos for l in world.landmarks]
            others = [o for o in world.agents if o is not agent]
            return torch.cat([agent.state.p_vel, agent.state.p_pos] + ent +
                             [o.state.p_pos - agent.state.p_pos for o in others] + [o.state.c for o in others], dim=1)

    return Scenario()


def test_user_scenario_reproduces_compiled_spread():
    from multiagent_particle_envs_b200 import MultiAgentEnv
    n = 6000
    sc = _spread_as_user_scenario()
    world = sc.make_world(num_envs=n)
    env = MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation)
    assert env.n == 3 and [s.shape for s in env.observation_space] == [(18,)] * 3 and env.shared_reward
    ref = make_product_env("simple_spread_n3", num_envs=n)
    obs0 = env.reset()
    ref.reset()
    # squeeze half of the worlds so that contacts are frequent, then share the state
    world.native.agent_pv[:, ::2, 0:2] *= 0.3
    ref.world.native.agent_pv.copy_(world.native.agent_pv)
    ref.world.native.lm_p.copy_(world.native.lm_p)
    p = world.native.agent_pv[:, :, 0:2]
    touching = ((p[0] - p[1]).norm(dim=1) < 0.3) | ((p[0] - p[2]).norm(dim=1) < 0.3) | ((p[1] - p[2]).norm(dim=1) < 0.3)
    assert float(touching.float().mean()) > 0.2                 # the comparison really exercises contact forces
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(5):
        acts = [torch.softmax(3 * torch.randn(n, 5, device="cuda", generator=g), 1) for _ in range(3)]
        o1, r1, d1, i1 = env.step(acts)
        o2, r2, d2, i2 = ref.step(acts)
        assert torch.equal(world.native.agent_pv, ref.world.native.agent_pv)          # physics: bit-identical
        for a, b in zip(o1, o2):
            assert a.shape == b.shape and torch.equal(a, b)
        for a, b in zip(r1, r2):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
        assert not any(bool(x.any()) for x in d1) and i1 == {'n': [{}, {}, {}]}
    assert isinstance(obs0, list) and obs0[0].shape == (n, 18)


def test_generic_program_against_oracle_on_an_unusual_entity_table():
    from multiagent_particle_envs_b200 import Agent, Landmark, MultiAgentEnv, MultiDiscrete, TorchScenario, World
    from oracle import Oracle

    class Scenario(TorchScenario):
        def make_world(self, num_envs=None, device=None):
            world = World()
            world.dim_c = 3
            world.damping, world.dt = 0.2, 0.08
            world.agents = [Agent() for _ in range(5)]
            for i, ag in enumerate(world.agents):
                ag.name = 'agent %d' % i
                ag.size = 0.05 + 0.03 * i
                ag.collide = i != 3
                ag.silent = i not in (0, 2)
                ag.initial_mass = 1.0 + 0.5 * i
                ag.accel = None if i % 2 else 2.0 + i
                ag.max_speed = 0.7 if i in (1, 4) else None
            world.agents[0].movable = False                    # an immovable speaker
            world.landmarks = [Landmark() for _ in range(3)]
            for l, lm in enumerate(world.landmarks):
                lm.name, lm.size, lm.collide, lm.movable = 'landmark %d' % l, 0.1 + 0.1 * l, l != 1, False
            return self._finish_world(world, num_envs, device)

        def observation(self, agent, world):
            return torch.cat([agent.state.p_pos, agent.state.p_vel, world.agents[2].state.c], dim=1)

        def reward(self, agent, world):
            return -agent.state.p_pos.square().sum(1)

    n = 5000
    sc = Scenario()
    world = sc.make_world(num_envs=n)
    env = MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation)
    assert [s.shape for s in env.observation_space] == [(7,)] * 5
    assert env.action_space[0].n == 3 and isinstance(env.action_space[2], MultiDiscrete) and env.action_space[1].n == 5
    env.reset()
    nw, desc = world.native, world.descriptor()
    nw.agent_pv[:, :, 0:2] *= 0.35                               # crowd them: many contacts
    nw.agent_pv[1:, :, 2:4] = torch.empty(4, n, 2, device="cuda").uniform_(-1.5, 1.5)
    orc = Oracle(desc, "f64")
    assert orc.act_dims == [3, 5, 8, 5, 5]
    rng = np.random.RandomState(1)
    movable = [bool(desc.agent_movable[i]) for i in range(5)]
    for t in range(4):
        pv0 = nw.agent_pv.permute(1, 0, 2).cpu().numpy()
        lm0 = nw.lm_p.permute(1, 0, 2).cpu().numpy()
        act = random_actions(orc.act_dims, n, rng, movable=movable).astype(np.float32)
        obs_n, rew_n, done_n, _ = env.step([torch.as_tensor(np.ascontiguousarray(a), device="cuda")
                                            for a in split_cols(act, orc.act_dims)])
        u, c = orc.set_action(act)
        rpv, rcomm = orc.world_step(pv0, lm0, np.zeros((n, 5, 3)), u, c)
        pv = nw.agent_pv.permute(1, 0, 2).cpu().numpy()
        np.testing.assert_allclose(pv, rpv, rtol=1e-5, atol=1e-6)
        assert np.array_equal(pv[:, 0], pv0[:, 0])                               # the immovable agent stayed put
        np.testing.assert_allclose(world.agents[2].state.c.cpu().numpy(), rcomm[:, 2], rtol=1e-7)
        np.testing.assert_allclose(world.agents[0].state.c.cpu().numpy(), rcomm[:, 0], rtol=1e-7)
        assert float(world.agents[1].state.c.abs().max()) == 0.0                  # silent agents stay silent
        speed = np.linalg.norm(pv[:, [1, 4], 2:4], axis=2)
        assert 0.699 < speed.max() <= 0.7 * (1 + 1e-6)                             # the limit is reached, never exceeded
        assert torch.equal(obs_n[3][:, 4:7], world.agents[2].state.c) and rew_n[0].shape == (n,)


def test_user_scenarios_need_batched_mode_and_keep_failing_loudly_otherwise():
    from multiagent_particle_envs_b200 import MultiAgentEnv
    sc = _spread_as_user_scenario()
    world = sc.make_world()                                       # scalar mode
    with pytest.raises(NotImplementedError):
        MultiAgentEnv(world, sc.reset_world, sc.reward, sc.observation)
