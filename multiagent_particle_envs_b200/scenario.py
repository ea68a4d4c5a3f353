"""Scenario plugin base (reference: multiagent/scenario.py:4-10).

A scenario builds the World (`make_world`) and draws initial conditions (`reset_world`); by
convention it also provides `reward(agent, world)`, `observation(agent, world)` and optionally
`benchmark_data(agent, world)` (README.md:37-44), which MultiAgentEnv binds as callbacks
(make_env.py:41-43).

In this package the built-in scenarios' callbacks are *views onto the native program*: they
return the calling agent's slice of what the sm_100a observe kernel computed for the whole batch.
`native_program` names the compiled program (multiagent_particle_envs_b200/csrc/mpe_scenarios.cuh).
"""


class BaseScenario(object):
    native_program = None

    def make_world(self):
        raise NotImplementedError()

    def reset_world(self, world):
        raise NotImplementedError()


class _BatchedScenario(BaseScenario):
    def _finish_world(self, world, num_envs=None, device=None):
        world.native_program = self.native_program
        world.scenario = self
        if num_envs is not None:
            world.num_envs = num_envs
        if device is not None:
            world.device = device
        return world

    def reset_world(self, world, mask=None, seed=None):
        """i.i.d. uniform positions, zero velocity and comm (e.g. simple_spread.py:31-45) for the
        worlds selected by `mask` (all when None), drawn on the device from a Philox stream."""
        world.reset_states(mask=mask, seed=seed)


class TorchScenario(_BatchedScenario):
    """User-defined scenario.  `make_world()` fills the entity table exactly as in the reference (any mix of
    movable / colliding / silent agents and landmarks, up to 8 + 8, end it with `return self._finish_world(world,
    num_envs, device)`); `_set_action` and `World.step` then run on the generic native program, and the scenario's
    own `observation(agent, world)` / `reward(agent, world)` (optionally `benchmark_data`, and a `done` callback) are
    written with torch operations over the batched state: in a batched world `agent.state.p_pos`, `p_vel`, `c` and
    `landmark.state.p_pos` are CUDA tensors of shape [num_envs, 2] (or [num_envs, dim_c]); return [num_envs, obs_dim]
    and [num_envs].  Nothing runs on the CPU: the callbacks are vectorised over worlds on the GPU.  Batched mode
    only (`make_env(..., num_envs=N)` / `make_world(num_envs=N)`)."""
    native_program = "custom"


class NativeScenario(_BatchedScenario):
    """Shared implementation of the callback surface for scenarios that have a native program."""

    def observation(self, agent, world):
        return world.native_observation(agent)

    def reward(self, agent, world):
        return world.native_reward(agent)

    def benchmark_data(self, agent, world):
        return world.native_benchmark_data(agent)
