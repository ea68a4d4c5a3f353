"""simple_crypto: covert communication.  Nobody moves: Alice (agent 2, the speaker) knows the goal
"colour" and a private key, Bob (agent 1) knows the key, Eve (agent 0, adversary) only hears Alice; all
three emit a 4-dim utterance (reference: multiagent/scenarios/simple_crypto.py).

good agents' reward (:94-113): -|bob.c - goal|^2 + |eve.c - goal|^2 (terms skipped while the utterance
is all zeros); Eve's (:115-121): -|eve.c - goal|^2.  Observation (:124-174): Alice [goal, key] (8),
Bob [key, alice.c] (8), Eve [alice.c] (4).  Per-world indices: `world.native.goal[0]` = goal landmark,
`goal[1]` = key landmark; colours are one-hot in dim_c (:58-62).
Native program: Crypto in csrc/mpe_scenarios.cuh."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class CryptoAgent(Agent):
    def __init__(self):
        super(CryptoAgent, self).__init__()
        self.key = None


class Scenario(NativeScenario):
    native_program = "simple_crypto"

    def make_world(self, num_envs=None, device=None):
        world = World()
        num_agents, num_adversaries, num_landmarks = 3, 1, 2
        world.dim_c = 4
        world.agents = [CryptoAgent() for _ in range(num_agents)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.adversary = i < num_adversaries
            agent.speaker = (i == 2)
            agent.movable = False
            agent.color = np.array([0.75, 0.25, 0.25]) if agent.adversary else np.array([0.25, 0.25, 0.25])
        world.landmarks = [Landmark() for _ in range(num_landmarks)]
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.color = np.eye(world.dim_c)[i]
        return self._finish_world(world, num_envs, device)

    def good_listeners(self, world):
        return [agent for agent in world.agents if not agent.adversary and not agent.speaker]

    def good_agents(self, world):
        return [agent for agent in world.agents if not agent.adversary]

    def adversaries(self, world):
        return [agent for agent in world.agents if agent.adversary]
