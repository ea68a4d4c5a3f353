"""simple_speaker_listener: an immovable speaker sees the goal landmark's colour and talks (3-dim
channel); a silent listener hears it and must move there (reference:
multiagent/scenarios/simple_speaker_listener.py).

reward (:63-67), same for both and summed by the collaborative env: -|listener.pos - goal|^2.
Observation (:69-92): speaker = [goal colour] (3 floats), listener = [vel, landmarks - pos, speaker's
utterance] (11 floats).  Action spaces: speaker Discrete(3), listener Discrete(5).
Native program: SpeakerListener in csrc/mpe_scenarios.cuh."""
import numpy as np

from ..core import World, Agent, Landmark
from ..scenario import NativeScenario


class Scenario(NativeScenario):
    native_program = "simple_speaker_listener"

    def make_world(self, num_envs=None, device=None):
        world = World()
        world.dim_c = 3
        world.collaborative = True
        world.agents = [Agent() for _ in range(2)]
        for i, agent in enumerate(world.agents):
            agent.name = 'agent %d' % i
            agent.collide = False
            agent.size = 0.075
            agent.color = np.array([0.25, 0.25, 0.25])
        world.agents[0].movable = False     # speaker
        world.agents[1].silent = True       # listener
        world.landmarks = [Landmark() for _ in range(3)]
        colors = ([0.65, 0.15, 0.15], [0.15, 0.65, 0.15], [0.15, 0.15, 0.65])
        for i, landmark in enumerate(world.landmarks):
            landmark.name = 'landmark %d' % i
            landmark.collide = False
            landmark.movable = False
            landmark.size = 0.04
            landmark.color = np.array(colors[i])
        return self._finish_world(world, num_envs, device)
