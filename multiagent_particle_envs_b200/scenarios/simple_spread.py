"""simple_spread: cooperative navigation, N agents cover N landmarks
(reference: multiagent/scenarios/simple_spread.py; N = 3 there, :11-12).

Each agent's reward is -sum_l min_a |p_a - lm_l| minus 1 per agent it overlaps -- the reference's
loop includes the agent itself, so the constant -1 is reproduced (:72-82) -- and, the world being
`collaborative`, MultiAgentEnv hands every agent the sum (environment.py:100-102).
Observation: [vel, pos, landmarks - pos, others - pos, others' comm (zeros)] (:84-100).
Native program: Spread<N> in csrc/mpe_scenarios.cuh (N = 2..6 compiled)."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_spread"

    def __init__(self, num_agents=3):
        self.num_agents = num_agents

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 2
        world.collaborative = True
        world.agents = [Agent() for _ in range(self.num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.size = 0.15
            agent.color = np.array([0.35, 0.35, 0.85])
        world.landmarks = [Landmark() for _ in range(self.num_agents)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.color = np.array([0.25, 0.25, 0.25])
        return self._finish_world(world, num_envs, device)
