"""multiagent_particle_envs_b200 -- B200-native batched multi-agent particle environments.

A from-scratch implementation of the hot path of openai/multiagent-particle-envs
(World.step physics + per-agent observation/reward) as hand-written sm_100a CUDA kernels behind
the reference's own Python API (make_env / MultiAgentEnv / World / Scenario).  See DESIGN.md.

Importing the package does not prompt (the reference blocks on input(), multiagent/__init__.py:31)
and does not need gym.  There is no CPU fallback: stepping a world requires the CUDA extension
(multiagent_particle_envs_b200/csrc/libmpe_b200.so) and a B200.
"""
from .core import World, Agent, Landmark, Entity, EntityState, AgentState, Action  # noqa: F401
from .environment import MultiAgentEnv  # noqa: F401
from .multi_discrete import MultiDiscrete  # noqa: F401
from .scenario import BaseScenario, NativeScenario, TorchScenario  # noqa: F401
from .make_env import make_env  # noqa: F401

__version__ = "0.1.0"
