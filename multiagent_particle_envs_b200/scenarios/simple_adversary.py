"""simple_adversary: physical deception.  1 adversary + 2 good agents, 2 landmarks one of which is the
(per-world) goal; nothing collides (reference: multiagent/scenarios/simple_adversary.py).

good agents' reward (:76-105): -min_good |good - goal| + sum_adv |adv - goal|; adversary's (:107-118):
-|adv - goal|^2.  Observation (:121-139): good = [goal - pos, landmarks - pos, others - pos] (10 floats),
adversary = [landmarks - pos, others - pos] (8 floats; it does not see which landmark is the goal).
The goal index is drawn per world at reset (np.random.choice(world.landmarks), :44) and lives in
`world.native.goal[0]`.  Native program: Adversary<1,NGOOD,NGOOD> in csrc/mpe_scenarios.cuh, compiled for NGOOD = 2
(the reference) and 3 via Scenario(num_agents=4)."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_adversary"

    def __init__(self, num_agents=3):
        self.num_agents = num_agents

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 2
        num_agents, num_adversaries = self.num_agents, 1
        world.num_agents = num_agents
        world.agents = [Agent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.silent = True
            agent.adversary = i < num_adversaries
            agent.size = 0.15
            agent.color = np.array([0.85, 0.35, 0.35]) if agent.adversary else np.array([0.35, 0.35, 0.85])
        world.landmarks = [Landmark() for _ in range(num_agents - 1)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.08
            landmark.color = np.array([0.15, 0.15, 0.15])
        return self._finish_world(world, num_envs, device)

    def good_agents(self, world):
        return [agent for agent in world.agents if not agent.adversary]

    def adversaries(self, world):
        return [agent for agent in world.agents if agent.adversary]
