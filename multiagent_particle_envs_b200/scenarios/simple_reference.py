"""simple_reference: two agents, each knows the landmark the OTHER one must reach and says so over a
10-dim channel (reference: multiagent/scenarios/simple_reference.py).  Both agents move and speak, so
their action space is MultiDiscrete [5, 10] (flat 15-vector: 5 physical then 10 comm).

reward of agent i (:55-59): -|other.pos - landmark[goal_b_i]|^2; the world is collaborative, so the
env returns the sum.  Observation (:61-80): [vel, landmarks - pos, colour of goal_b_i, other's
utterance] (21 floats).  Per-world goals: `world.native.goal[i]` = agents[i].goal_b.
Native program: Reference in csrc/mpe_scenarios.cuh."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_reference"

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 10
        world.collaborative = True
        world.agents = [Agent() for _ in range(2)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.color = np.array([0.25, 0.25, 0.25])
        world.landmarks = [Landmark() for _ in range(3)]
        colors = ([0.75, 0.25, 0.25], [0.25, 0.75, 0.25], [0.25, 0.25, 0.75])
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.color = np.array(colors[i])
        return self._finish_world(world, num_envs, device)
