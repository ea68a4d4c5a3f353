"""simple: one agent, one landmark, nothing collides (reference: multiagent/scenarios/simple.py).
reward = -|p_agent - p_landmark|^2 (:41-43); observation = [vel, landmark - pos] (:45-50).
Native program: Simple<1,1> in csrc/mpe_scenarios.cuh."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple"

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.agents = [Agent()]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.silent = True
            agent.color = np.array([0.25, 0.25, 0.25])
        world.landmarks = [Landmark()]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.color = np.array([0.75, 0.25, 0.25])
        return self._finish_world(world, num_envs, device)
