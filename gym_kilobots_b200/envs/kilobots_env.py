"""KilobotsEnv -- the reference's abstract gym env (gym_kilobots/envs/kilobots_env.py:16-290) as an
E = 1 facade over the batched CUDA step.

Same public surface: reset()/step()/seed()/close(), kilobots / objects / num_kilobots / action_space,
get_state()/get_observation()/get_objects()/get_kilobots()/get_light(), and the subclass hooks
_configure_environment / get_reward / has_finished / get_info / _add_kilobot / _add_object.
`self.world` is a recorder (lib/world.py); `_configure_environment` builds the scene with the same
constructor calls as the reference and `reset()` uploads it.  The whole sub-step loop of
`step` (:168-190) is ONE kernel launch (kb_step_host); there is no CPU fallback.
"""
import abc

import numpy as np

from .. import _abi as abi
from .. import scene as S
from ..lib.body import Body, _world_scale  # noqa: F401
from ..lib.kilobot import Kilobot, SimpleAccelerationControlKilobot, SimpleVelocityControlKilobot
from ..lib.light import Light
from ..lib.world import World

try:
    import gym
    _Base = gym.Env
except Exception:  # gym is not installable in this image
    _Base = object


class UnknownObjectException(Exception):
    pass


class UnknownLightTypeException(Exception):
    pass


class KilobotsEnv(_Base):
    metadata = {'render.modes': ['human']}

    world_size = world_width, world_height = 2., 1.5
    screen_size = screen_width, screen_height = 1200, 900

    _observe_objects = False
    _observe_light = True

    __sim_steps_per_second = 10
    __sim_velocity_iterations = 10
    __sim_position_iterations = 10
    __steps_per_action = 10

    # Box2D-version switches; subclasses may override
    _wall_edges = 3
    _damping_mode = 0
    _enable_toi = True
    _device = 0

    def __new__(cls, **kwargs):
        cls.sim_steps_per_second = cls.__sim_steps_per_second
        cls.sim_step = 1. / cls.__sim_steps_per_second
        cls.world_x_range = -cls.world_width / 2, cls.world_width / 2
        cls.world_y_range = -cls.world_height / 2, cls.world_height / 2
        cls.world_bounds = (np.array([-cls.world_width / 2, -cls.world_height / 2]),
                            np.array([cls.world_width / 2, cls.world_height / 2]))
        return super(KilobotsEnv, cls).__new__(cls)

    def __init__(self, **kwargs):
        self.__sim_steps = 0
        self.__reset_counter = 0
        self.world = World()
        self.world._env = self
        self._real_time = False
        self._kilobots: [Kilobot] = []
        self._objects: [Body] = []
        self._light: Light = None
        self.__seed = 0
        self._screen = None
        self.render_mode = 'human'
        self.video_path = None
        self._batch = None
        self._batch_key = None
        self._configure_environment()
        # the reference drops the kilobots built by the constructor pass (kilobots_env.py:67-68)
        self.world._unregister(self._kilobots)
        self._kilobots = []

    # ------------------------------------------------------------------------- properties
    @property
    def _sim_steps(self):
        return self.__sim_steps

    @property
    def kilobots(self):
        return tuple(self._kilobots)

    @property
    def num_kilobots(self):
        return len(self._kilobots)

    @property
    def objects(self):
        return tuple(self._objects)

    @property
    def action_space(self):
        if self._light:
            return self._light.action_space

    @property
    def observation_space(self):
        return NotImplemented

    @property
    def state_space(self):
        return NotImplemented

    @property
    def _steps_per_action(self):
        return self.__steps_per_action

    def _add_kilobot(self, kilobot: Kilobot):
        self._kilobots.append(kilobot)

    def _add_object(self, body: Body):
        self._objects.append(body)

    @abc.abstractmethod
    def _configure_environment(self):
        raise NotImplementedError

    # ------------------------------------------------------------------------------ state
    def get_state(self):
        light = self._light.get_state() if self._light else np.zeros(0)   # reference crashes without a light (D7)
        return {'kilobots': np.array([k.get_state() for k in self._kilobots]),
                'objects': np.array([o.get_state() for o in self._objects]),
                'light': light}

    def get_observation(self):
        return self.get_state()

    @abc.abstractmethod
    def get_reward(self, state, action, new_state):
        raise NotImplementedError

    def has_finished(self, state, action):
        return False

    def get_info(self, state, action):
        return ""

    def destroy(self):
        self.world._unregister_all()
        del self._objects[:]
        del self._kilobots[:]
        self._light = None

    def close(self):
        self.destroy()
        if self._batch is not None:
            self._batch.close()
            self._batch = None

    def seed(self, seed=None):
        if seed is not None:
            self.__seed = seed
        return [self.__seed]

    # ------------------------------------------------------------------- scene -> device
    def _scene_spec(self):
        bodies = [o._spec() for o in self._objects] + [k._spec() for k in self._kilobots]
        lights = self._light._specs() if self._light else []
        return S.SceneSpec(bodies=bodies, num_objects=len(self._objects), lights=lights,
                           world_size=(self.world_width, self.world_height), wall_edges=self._wall_edges,
                           steps_per_action=self.__steps_per_action,
                           velocity_iterations=self.__sim_velocity_iterations,
                           position_iterations=self.__sim_position_iterations, dt=self.sim_step,
                           damping_mode=self._damping_mode, enable_toi=self._enable_toi)

    def _make_batch(self, spec):
        """Factory of the batched backend (E = 1).  Product path: the CUDA library, nothing else."""
        from .. import _native
        # max_contacts = -1: a slot for every proxy pair -- the reference's spawn may stack kilobots, and a single
        # env has no throughput reason to bound its contact list
        return _native.NativeBatch(spec, 1, max_contacts=-1, device=self._device)

    @staticmethod
    def _spec_key(spec):
        def fx(f):
            v = None if f.vertices is None else tuple(np.asarray(f.vertices, float).ravel())
            return (f.shape, f.density, f.friction, f.restitution, f.radius, f.hx, f.hy, v)
        return (tuple((b.kind, b.linear_damping, b.angular_damping, tuple(fx(f) for f in b.fixtures)) for b in spec.bodies),
                spec.num_objects,
                tuple((l.type, l.radius, tuple(np.asarray(l.bounds, float).ravel()),
                       tuple(np.asarray(l.action_bounds, float).ravel()), l.relative_actions, l.max_velocity)
                      for l in spec.lights),
                spec.world_size, spec.wall_edges, spec.damping_mode, spec.enable_toi)

    def _sync_mirror(self):
        b = self._batch
        self.world._mirror = b.bodies()[0]
        ctrl, light = b.controllers()
        for i, k in enumerate(self._kilobots):
            k._ctrl = ctrl[0, i].copy()
            if isinstance(k, SimpleVelocityControlKilobot):
                k._velocity = ctrl[0, i, :2].copy()
        if self._light:
            self._light._load_state(light[0])
        pairs, count = b.contacts()
        proxy_slot = {}
        p = self._scene_spec_cache.wall_edges
        for w in range(p):
            proxy_slot[w] = -1
        for slot, body in enumerate(self._scene_spec_cache.bodies):
            for _ in body.fixtures:
                proxy_slot[p] = slot
                p += 1
        self.world._contacts = (pairs[0], int(count[0]), proxy_slot)

    def _set_body_pose(self, slot, pose):
        b = self._batch
        raw = b.bodies()[0]
        poses = np.stack([raw[:, 8].astype(np.float64) / 25.0, raw[:, 9].astype(np.float64) / 25.0,
                          raw[:, 2].astype(np.float64)], axis=-1)
        poses[slot] = pose
        mask = np.zeros((1, len(poses)), np.uint8)
        mask[0, slot] = 1   # b2Body::SetTransform acts on this body only (lib/body.py:67-69)
        b.set_poses(poses[None], mask)
        self.world._mirror = b.bodies()[0]

    def _record_scene(self):
        """destroy + _configure_environment (kilobots_env.py:152-155), recorded instead of simulated:
        -> (SceneSpec, poses [B,3] m/rad, light state [L] or None, kilobot (v, omega) [N,2] or None)."""
        self.destroy()
        self._configure_environment()
        spec = self._scene_spec()
        ordered = list(self._objects) + list(self._kilobots)
        pose = np.stack([body._init_pose for body in ordered])
        light = self._light._state_vector() if self._light else None
        vel = None
        if any(isinstance(k, SimpleVelocityControlKilobot) for k in self._kilobots):
            vel = np.stack([np.asarray(getattr(k, '_velocity', np.zeros(2)), float) for k in self._kilobots])
        return spec, pose, light, vel

    def reset(self):
        self.__reset_counter += 1
        spec, pose, light, vel = self._record_scene()
        self.__sim_steps = 0
        key = self._spec_key(spec)
        if self._batch is None or key != self._batch_key:
            if self._batch is not None:
                self._batch.close()
            self._batch = self._make_batch(spec)
            self._batch_key = key
            self._out = None   # host output buffers are sized for the batch they were made for
        self._scene_spec_cache = spec
        ordered = list(self._objects) + list(self._kilobots)
        self.world._slot = {id(body): i for i, body in enumerate(ordered)}
        pose = pose[None]
        light = None if light is None else light[None]
        vel = None if vel is None else vel[None]
        # bodies at their poses, then one "step to resolve" (kilobots_env.py:157)
        self._batch.reset(pose, light, vel)
        self._check_status(self._batch.get_status())
        self._sync_mirror()
        return self.get_observation()

    def _check_status(self, status):
        """Capacity overflow / non-finite poses are errors in the drop-in path: the reference (Box2D) has no
        capacity to overflow, so a dropped pair would silently be different physics."""
        if np.any(np.asarray(status) != 0) and not getattr(self, "allow_status_flags", False):
            from .._native import KbStatusError, describe_status
            raise KbStatusError(describe_status(status))

    def _step_batch(self, action, mode):
        out = self._out = getattr(self, '_out', None) or {
            "kilobots": np.zeros((1, len(self._kilobots), 3), np.float32),
            "objects": np.zeros((1, len(self._objects), 3), np.float32),
            "light": np.zeros((1, self._batch.L), np.float64),
            "reward": np.zeros(1, np.float32), "done": np.zeros(1, np.uint8), "status": np.zeros(1, np.int32)}
        if (out["kilobots"].shape[1] != len(self._kilobots) or out["objects"].shape[1] != len(self._objects)
                or out["light"].shape[1] != self._batch.L):
            self._out = None
            return self._step_batch(action, mode)
        if hasattr(self._batch, "step_host"):
            self._batch.step_host(action, mode, out)
        else:
            out.update(self._batch.step(action, mode))
        self._check_status(out["status"])
        return out

    def step(self, action: np.ndarray):
        if self._batch is None:
            raise RuntimeError("KilobotsEnv.step called before reset()")
        state = self.get_state()
        if action is not None and self._light:
            act = np.asarray(action, dtype=np.float64).reshape(1, -1)
            self._step_batch(act, abi.KB_ACTION_LIGHT)
        else:
            self._step_batch(None, abi.KB_ACTION_NONE)
        self.__sim_steps += self.__steps_per_action
        self._sync_mirror()
        next_state = self.get_state()
        observation = self.get_observation()
        reward = self.get_reward(state, action, next_state)
        done = self.has_finished(next_state, action)
        info = self.get_info(next_state, action)
        return observation, reward, done, info

    def render(self, mode=None):
        """mode='rgb_array': the frame the reference draws through kb_rendering.KilobotsViewer (:221-275), rasterised
        on the device (kb_render) and returned as uint8 [screen_height, screen_width, 3].  There is no window:
        'human' mode (pygame display, real-time pacing, mp4 recording) is outside the accelerated path."""
        if mode is None:
            mode = self.render_mode
        if mode != 'rgb_array':
            raise NotImplementedError("only render(mode='rgb_array') is provided; the pygame viewer is out of scope")
        if self._batch is None:
            raise RuntimeError("KilobotsEnv.render called before reset()")
        img = self._batch.render((0,), self.screen_width, self.screen_height)
        return np.asarray(img.cpu() if hasattr(img, "cpu") else img)[0]

    def get_objects(self) -> [Body]:
        return self._objects

    def get_kilobots(self) -> [Kilobot]:
        return self._kilobots

    def get_light(self) -> Light:
        return self._light
