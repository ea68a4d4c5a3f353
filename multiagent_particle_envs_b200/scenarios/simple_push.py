"""simple_push: keep-away.  1 adversary + 1 good agent (they collide), 2 landmarks one of which is the
per-world goal (reference: multiagent/scenarios/simple_push.py).

good reward (:62-64): -|agent - goal|; adversary reward (:66-74): min_good |good - goal| - |adv - goal|.
Observation (:76-96): good = [vel, goal - pos, own colour, landmarks - pos, landmark colours, other - pos]
(19 floats; the agent's colour encodes the goal index, :47-53), adversary = [vel, landmarks - pos,
other - pos] (8 floats).  Native program: Push<1,1,2> in csrc/mpe_scenarios.cuh."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_push"

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 2
        num_agents, num_adversaries, num_landmarks = 2, 1, 2
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.adversary = i < num_adversaries
            agent.color = np.array([0.75, 0.25, 0.25]) if agent.adversary else np.array([0.25, 0.25, 0.25])
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.color = np.array([0.1, 0.1, 0.1])
            landmark.color[i + 1] += 0.8
            landmark.index = i
        return self._finish_world(world, num_envs, device)
