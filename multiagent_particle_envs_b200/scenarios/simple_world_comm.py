"""simple_world_comm: simple_tag plus food, forests that hide agents, and a leader adversary
that broadcasts a 4-dim message (reference: multiagent/scenarios/simple_world_comm.py).

4 adversaries (agent 0 = leader, the only non-silent agent, MultiDiscrete action [5, 4]),
2 prey, landmarks = 1 obstacle ++ 2 food ++ 2 forests (:7-57).
prey reward (:155-183): -5 per adversary in contact, -2*bound per axis, +2 per food touched,
+0.05 * distance to the nearest food; adversary reward (:185-198): -0.1 * distance from *this*
adversary to the nearest prey, +5 per (prey, adversary) pair in contact.
Observation (:224-287): others are zeroed unless visible (same forest, or neither in a forest, or
the observer is the leader); adversaries get 34 floats (incl. the leader's message), prey 28.
Native program: WorldComm<4,2,1,2> in csrc/mpe_scenarios.cuh."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_world_comm"

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 4
        num_good_agents, num_adversaries = 2, 4
        num_landmarks, num_food, num_forests = 1, 2, 2
        world.agents = [Agent() for _ in range(num_adversaries + num_good_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = True
            agent.leader = (i == 0)
            agent.silent = (i > 0)
            agent.adversary = i < num_adversaries
            agent.size = 0.075 if agent.adversary else 0.045
            agent.accel = 3.0 if agent.adversary else 4.0
            agent.max_speed = 1.0 if agent.adversary else 1.3
            agent.color = np.array([0.95, 0.45, 0.45]) if agent.adversary else np.array([0.45, 0.95, 0.45])
            if agent.leader:
                agent.color = agent.color - np.array([0.3, 0.3, 0.3])

        def group(n, prefix, size, collide, color):
            out = [Landmark() for _ in range(n)]
            for i, lm in enumerate(out):
                lm.name = '%s %d' % (prefix, i)
                lm.collide = collide
                lm.movable = False
                lm.size = size
                lm.boundary = False
                lm.color = np.array(color)
            return out

        obstacles = group(num_landmarks, 'landmark', 0.2, True, [0.25, 0.25, 0.25])
        world.food = group(num_food, 'food', 0.03, False, [0.15, 0.15, 0.65])
        world.forests = group(num_forests, 'forest', 0.3, False, [0.6, 0.9, 0.6])
        world.landmarks = obstacles + world.food + world.forests
        return self._finish_world(world, num_envs, device)

    def good_agents(self, world):
        return [agent for agent in world.agents if not agent.adversary]

    def adversaries(self, world):
        return [agent for agent in world.agents if agent.adversary]
