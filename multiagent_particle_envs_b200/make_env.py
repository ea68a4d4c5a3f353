"""Environment factory (reference: make_env.py:15-44).

    env = make_env('simple_spread')                      # scalar, reference-compatible
    env = make_env('simple_spread', num_envs=65536)      # one batch of worlds on the current GPU

Loads the scenario module, builds the World and wires the scenario's reset / reward / observation
(/ benchmark_data when benchmark=True) callbacks into MultiAgentEnv exactly as the reference does
(make_env.py:36-43).  Extra keyword arguments are batch extensions with reference-compatible
defaults: num_envs (None = scalar API), device, seed, and (rank, world_size) to take this
process's contiguous shard of a global batch of num_envs worlds.
"""


def make_env(scenario_name, benchmark=False, num_envs=None, device=None, seed=0, rank=0, world_size=1,
             **scenario_kwargs):
    from .environment import MultiAgentEnv
    from . import scenarios
    from .sharding import shard_range

    scenario = scenarios.load(scenario_name + ".py").Scenario(**scenario_kwargs)
    offset = 0
    if num_envs is not None and world_size > 1:
        offset, stop = shard_range(num_envs, rank, world_size)
        num_envs = stop - offset
    world = scenario.make_world(num_envs=num_envs, device=device)
    world.seed = seed
    world.world_offset = offset
    if benchmark:
        env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation,
                            scenario.benchmark_data)
    else:
        env = MultiAgentEnv(world, scenario.reset_world, scenario.reward, scenario.observation)
    return env
