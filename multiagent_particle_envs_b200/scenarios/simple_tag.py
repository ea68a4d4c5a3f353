"""simple_tag: predator-prey.  3 slower adversaries chase 1 faster prey around 2 colliding
obstacles (reference: multiagent/scenarios/simple_tag.py).

prey reward: -10 per adversary in contact, minus bound(|x|) + bound(|y|) for leaving the arena
(:89-113); every adversary: +10 per (prey, adversary) pair in contact (:115-129).
Observation: [vel, pos, obstacles - pos, others - pos, prey velocities] (:131-147).
Native program: Tag<NADV,NGOOD,L> in csrc/mpe_scenarios.cuh; compiled for (3,1,2) -- the reference's counts --
and (1,1,2), (2,1,2), (4,2,2), (6,2,3) via Scenario(num_adversaries=, num_good_agents=, num_landmarks=)."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_tag"

    def __init__(self, num_adversaries=3, num_good_agents=1, num_landmarks=2):
        self.counts = (num_good_agents, num_adversaries, num_landmarks)

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 2
        num_good_agents, num_adversaries, num_landmarks = self.counts
        world.agents = [Agent() for _ in range(num_adversaries + num_good_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.silent = True
            agent.adversary = i < num_adversaries
            agent.size = 0.075 if agent.adversary else 0.05
            agent.accel = 3.0 if agent.adversary else 4.0
            agent.max_speed = 1.0 if agent.adversary else 1.3
            agent.color = np.array([0.85, 0.35, 0.35]) if agent.adversary else np.array([0.35, 0.85, 0.35])
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = True
            landmark.movable = False
            landmark.size = 0.2
            landmark.boundary = False
            landmark.color = np.array([0.25, 0.25, 0.25])
        return self._finish_world(world, num_envs, device)

    def good_agents(self, world):
        return [agent for agent in world.agents if not agent.adversary]

    def adversaries(self, world):
        return [agent for agent in world.agents if agent.adversary]
