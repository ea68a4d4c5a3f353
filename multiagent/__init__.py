"""Drop-in alias: `multiagent` -> multiagent_particle_envs_b200.

Existing MADDPG-style code does `from multiagent.environment import MultiAgentEnv`,
`import multiagent.scenarios as scenarios`, `from multiagent.core import World, Agent, Landmark`
and `from multiagent.scenario import BaseScenario`; these imports keep working unchanged and
resolve to the B200-native implementation.  No gym registration side effects, no input() prompt
(the reference's multiagent/__init__.py:9-32 does both)."""
import sys

import multiagent_particle_envs_b200 as _impl
from multiagent_particle_envs_b200 import core, environment, multi_discrete, scenario, scenarios  # noqa: F401

for _name in ("core", "environment", "multi_discrete", "scenario", "scenarios"):
    sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
