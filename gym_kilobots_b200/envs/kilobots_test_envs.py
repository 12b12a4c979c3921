"""Example scenes of gym_kilobots/envs/kilobots_test_envs.py, built with the same constructor calls."""
import numpy as np

from ..lib.body import CForm, CornerQuad, LForm, TForm, Triangle
from ..lib.kilobot import PhototaxisKilobot
from ..lib.light import CircularGradientLight
from .kilobots_env import KilobotsEnv


class QuadPushingEnv(KilobotsEnv):
    """Abstract shell in the reference (no _configure_environment, kilobots_test_envs.py:11-20)."""
    world_size = world_width, world_height = 1., .5


class QuadAssemblyKilobotsEnv(KilobotsEnv):
    def __init__(self, seed=None):
        self._rng = np.random.default_rng(seed)
        super().__init__()

    def _uniform(self, loc, scale):
        # scipy.stats.uniform(loc, scale).rvs() of the reference (:26-28)
        return np.asarray(loc) + self._rng.random(2) * np.asarray(scale)

    def _configure_environment(self):
        swarm_spawn_location = self._uniform((-.95, -.7), (.9, 1.4))
        obj_spawn_location = self._uniform((.05, -.7), (.9, .65))
        self._objects = [
            CornerQuad(world=self.world, width=.15, height=.15, position=(.45, .605)),
            CornerQuad(world=self.world, width=.15, height=.15, position=(.605, .605), orientation=-np.pi / 2),
            CornerQuad(world=self.world, width=.15, height=.15, position=(.605, .45), orientation=-np.pi),
            CornerQuad(world=self.world, width=.15, height=.15, position=obj_spawn_location, orientation=-np.pi / 2)
        ]
        self._light = CircularGradientLight(position=swarm_spawn_location.copy())
        offsets = [(.0, .0), (.03, .0), (.0, .03), (-.03, .0), (.0, -.03)] * 3
        self._kilobots = [PhototaxisKilobot(self.world, position=swarm_spawn_location + off, light=self._light)
                          for off in offsets]

    def has_finished(self, state, action):
        return False

    def get_reward(self, state, action, new_state):
        return 1.

    def get_info(self, state, action):
        return None


class TriangleTestEnv(KilobotsEnv):
    def _configure_environment(self):
        self._objects = [Triangle(world=self.world, width=.15, height=.15, position=(.0, .0)),
                         LForm(world=self.world, width=.15, height=.15, position=(.0, .3)),
                         TForm(world=self.world, width=.15, height=.15, position=(.0, -.3)),
                         CForm(world=self.world, width=.15, height=.15, position=(.3, .0))]

    def has_finished(self, state, action):
        return False

    def get_reward(self, state, action, new_state):
        return 1.

    def get_info(self, state, action):
        return None
