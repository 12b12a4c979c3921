"""YamlKilobotsEnv -- YAML-configured scenes with the reference's tags, shape names, spawn
distributions and Box spaces (gym_kilobots/envs/yaml_kilobots_env.py:16-369)."""
from random import shuffle

import numpy as np
import yaml

from .. import lib as _lib
from ..lib import CForm, Circle, CircularGradientLight, CompositeLight, CornerQuad, GradientLight, LForm, Quad, \
    TForm, Triangle
from ..lib.light import MomentumLight, SinglePositionLight
from ..spaces import Box
from .kilobots_env import KilobotsEnv, UnknownLightTypeException, UnknownObjectException


class _Conf(yaml.YAMLObject):
    def __eq__(self, other):
        for k in self.__dict__:
            if k not in other.__dict__:
                return False
            if not self.__getattribute__(k) == other.__getattribute__(k):
                return False
        return True

    __hash__ = None


class EnvConfiguration(_Conf):
    yaml_tag = '!EvalEnv'

    class ObjectConfiguration(_Conf):
        yaml_tag = '!ObjectConf'

        def __init__(self, idx, color, shape, width, height, init, symmetry=None):
            self.idx = idx
            self.shape = shape
            self.width = width
            self.height = height
            self.init = init
            self.color = color
            self.symmetry = symmetry

        @property
        def object_type(self):
            _type = self.shape
            if _type in ['corner_quad', 'corner-quad', 'quad']:
                _type = 'square'
            return _type

    class LightConfiguration(_Conf):
        yaml_tag = '!LightConf'

        def __init__(self, obj_type=None, init=None, radius=None, components=None, type=None):
            self.type = obj_type if obj_type is not None else type
            self.init = init
            self.radius = radius
            if components is not None:
                self.components = [c if isinstance(c, EnvConfiguration.LightConfiguration)
                                   else EnvConfiguration.LightConfiguration(**c) for c in components]

    class KilobotsConfiguration(_Conf):
        yaml_tag = '!KilobotsConf'

        def __init__(self, num, mean, std, type='SimplePhototaxisKilobot'):
            self.num = num
            self.mean = mean
            self.std = std
            self.type = type

    def __init__(self, width, height, resolution, objects, light, kilobots):
        self.width = width
        self.height = height
        self.resolution = resolution
        self.objects = [o if isinstance(o, self.ObjectConfiguration) else self.ObjectConfiguration(**o) for o in objects]
        self.light = light if isinstance(light, self.LightConfiguration) else self.LightConfiguration(**light)
        self.kilobots = kilobots if isinstance(kilobots, self.KilobotsConfiguration) \
            else self.KilobotsConfiguration(**kilobots)


class YamlKilobotsEnv(KilobotsEnv):
    def __new__(cls, *, configuration, **kwargs):
        cls.world_width = configuration.width
        cls.world_height = configuration.height
        cls.world_size = cls.world_width, cls.world_height
        cls.screen_width = int(configuration.resolution * configuration.width)
        cls.screen_height = int(configuration.resolution * configuration.height)
        cls.screen_size = cls.screen_width, cls.screen_width
        return super(YamlKilobotsEnv, cls).__new__(cls, **kwargs)

    def __eq__(self, other):
        return self.conf == other.conf

    __hash__ = None

    def __init__(self, *, configuration, **kwargs):
        self.conf = configuration
        self._progress_factor = 1.
        self._iteration_counter = 0
        super().__init__(**kwargs)

    @property
    def progress_factor(self):
        return self._progress_factor

    @progress_factor.setter
    def progress_factor(self, pf):
        assert .0 <= pf <= 1., 'progress_factor must be a value in the range [.0, 1.]'
        self._progress_factor = pf

    @property
    def iteration_counter(self):
        return self._iteration_counter

    @iteration_counter.setter
    def iteration_counter(self, ic):
        assert isinstance(ic, int) and 0 <= ic, 'iteration_counter must be a positive integer'
        self._iteration_counter = ic

    def inc_iteration_counter(self):
        self._iteration_counter += 1

    def _configure_environment(self):
        self._init_objects()
        self._init_light()
        self._init_kilobots(getattr(self.conf.kilobots, 'type', 'SimplePhototaxisKilobot'))  # honours `type` (D10)

    # ------------------------------------------------------------------------------ spaces
    @property
    def state_space(self):
        lo, hi = self.kilobots_state_space.low, self.kilobots_state_space.high
        if self.light_state_space:
            lo = np.concatenate((lo, self.light_state_space.low))
            hi = np.concatenate((hi, self.light_state_space.high))
        if self.object_state_space:
            lo = np.concatenate((lo, self.object_state_space.low))
            hi = np.concatenate((hi, self.object_state_space.high))
        return Box(low=lo, high=hi, dtype=np.float32)

    @property
    def observation_space(self):
        lo, hi = self.kilobots_state_space.low, self.kilobots_state_space.high
        if self.light_observation_space:
            lo = np.concatenate((lo, self.light_observation_space.low))
            hi = np.concatenate((hi, self.light_observation_space.high))
        if self.object_observation_space:
            lo = np.concatenate((lo, self.object_observation_space.low))
            hi = np.concatenate((hi, self.object_observation_space.high))
        return Box(low=lo, high=hi, dtype=np.float32)

    @property
    def object_state_space(self):
        lo = np.array([self.world_x_range[0], self.world_y_range[0], -np.inf] * len(self._objects))
        hi = np.array([self.world_x_range[1], self.world_y_range[1], np.inf] * len(self._objects))
        return Box(low=lo, high=hi, dtype=np.float64)

    @property
    def object_observation_space(self):
        lo = np.array([self.world_x_range[0], self.world_y_range[0], -1., -1.] * len(self._objects))
        hi = np.array([self.world_x_range[1], self.world_y_range[1], 1., 1.] * len(self._objects))
        return Box(low=lo, high=hi, dtype=np.float64)

    @property
    def action_space(self):
        if self._light:
            return self._light.action_space
        return None

    @property
    def light_state_space(self):
        if self._light:
            return self._light.observation_space
        return None

    @property
    def light_observation_space(self):
        if self._light and self._observe_light:
            return self._light.observation_space
        return None

    @property
    def kilobots_state_space(self):
        lo = np.array([self.world_x_range[0], self.world_y_range[0]] * len(self._kilobots))
        hi = np.array([self.world_x_range[1], self.world_y_range[1]] * len(self._kilobots))
        return Box(low=lo, high=hi, dtype=np.float64)

    @property
    def kilobots_observation_space(self):
        return self.kilobots_state_space

    # ----------------------------------------------------------------------------- objects
    def _init_objects(self):
        for o in self.conf.objects:
            self._init_object(o.shape, o.width, o.height, o.init, o.color)

    def _get_random_object_init(self):
        init_position = np.random.rand(2) * np.asarray(self.world_size) + self.world_bounds[0]
        init_position *= 0.7
        init_orientation = np.random.rand() * 2 * np.pi - np.pi
        return np.r_[init_position, init_orientation]

    def _init_object(self, object_shape, object_width, object_height, object_init, object_color=None):
        if isinstance(object_init, str) and object_init == 'random':
            object_init = self._get_random_object_init()
        kw = dict(position=object_init[:2], orientation=object_init[2], world=self.world)
        if object_shape in ['square', 'quad', 'rect']:
            obj = Quad(width=object_width, height=object_height, **kw)
        elif object_shape in ['corner_quad', 'corner-quad']:
            obj = CornerQuad(width=object_width, height=object_height, **kw)
        elif object_shape == 'triangle':
            obj = Triangle(width=object_width, height=object_height, **kw)
        elif object_shape == 'circle':
            obj = Circle(radius=object_width, **kw)   # sic: radius = width (yaml_kilobots_env.py:228-230)
        elif object_shape == 'l_shape':
            obj = LForm(width=object_width, height=object_height, **kw)
        elif object_shape == 't_shape':
            obj = TForm(width=object_width, height=object_height, **kw)
        elif object_shape == 'c_shape':
            obj = CForm(width=object_width, height=object_height, **kw)
        else:
            raise UnknownObjectException('Shape of form {} not known.'.format(object_shape))
        if object_color:
            obj.color = object_color
        self._add_object(obj)

    # ------------------------------------------------------------------------------- light
    def _init_light(self):
        if not hasattr(self.conf, 'light') or self.conf.light is None:
            return
        self._light = self._init_light_from_config(self.conf.light)

    def _get_random_light_init(self, at_object=False):
        if at_object:
            which_object = self._objects[np.random.choice(len(self._objects), 1)[0]]
            init_position = which_object.get_position()
            radius = 1.2 * max(which_object.width, which_object.height) / 2
            angle = np.random.rand() * 2 * np.pi - np.pi
            init_position += (np.cos(angle) * radius, np.sin(angle) * radius)
        else:
            init_position = np.random.rand(2) * np.asarray(self.world_size) + self.world_bounds[0]
        return init_position

    def _init_light_from_config(self, light_config):
        light = None
        if light_config.type in ['circular', 'momentum']:
            light_bounds = np.array(self.world_bounds) * 1.1
            if isinstance(light_config.init, str) and light_config.init == 'random':
                init_position = self._get_random_light_init()
            elif isinstance(light_config.init, str) and light_config.init == 'object':
                init_position = self._get_random_light_init(at_object=True)
            else:
                init_position = np.asarray(light_config.init, dtype=np.float64)
            action_bounds = np.array([-1, -1]) * .01, np.array([1, 1]) * .01
            if light_config.type == 'circular':
                light = CircularGradientLight(position=init_position, radius=light_config.radius,
                                              bounds=light_bounds, action_bounds=action_bounds)
            else:
                init_angle = np.random.rand() * 2 * np.pi - np.pi
                init_velocity = np.array([np.sin(init_angle), np.cos(init_angle)]) * .01
                light = MomentumLight(position=init_position, velocity=init_velocity, max_velocity=.01,
                                      radius=light_config.radius, bounds=light_bounds, action_bounds=action_bounds)
        elif light_config.type == 'linear':
            light = GradientLight(angle=light_config.init)
        elif light_config.type == 'composite':
            lights = []
            if isinstance(light_config.init, str) and light_config.init == 'random':
                shuffle(light_config.components)
            for _c in light_config.components:
                lights.append(self._init_light_from_config(_c))
            light = CompositeLight(lights)
        else:
            raise UnknownLightTypeException()
        return light

    # ---------------------------------------------------------------------------- kilobots
    def _init_kilobots(self, type='SimplePhototaxisKilobot'):
        num_kilobots = self.conf.kilobots.num
        spawn_mean = self.conf.kilobots.mean
        spawn_std = self.conf.kilobots.std
        if isinstance(spawn_mean, str) and spawn_mean == 'light':
            if isinstance(self._light, SinglePositionLight):
                spawn_mean = self._light.get_position()
            elif isinstance(self._light, CompositeLight):
                lights_positions = np.asarray([_l.get_position() for _l in self._light.lights])
                idx = np.random.choice(np.arange(len(lights_positions)), num_kilobots)
                spawn_mean = lights_positions[idx]
            else:
                spawn_mean = 'random'
        if isinstance(spawn_mean, str) and spawn_mean == 'random':
            spawn_mean = np.random.rand(2) * np.asarray(self.world_size) + self.world_bounds[0]
            spawn_mean *= 0.9
        kilobot_positions = np.random.normal(scale=spawn_std, size=(num_kilobots, 2))
        kilobot_positions += spawn_mean
        kb_class = getattr(_lib, type)
        for position in kilobot_positions:
            position = np.maximum(position, self.world_bounds[0] + 0.02)
            position = np.minimum(position, self.world_bounds[1] - 0.02)
            self._add_kilobot(kb_class(self.world, position=position, light=self._light))

    def get_reward(self, state, action, new_state):
        return .0
