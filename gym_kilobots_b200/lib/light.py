"""Light models -- same classes and constructor signatures as gym_kilobots/lib/light.py.

On the hot path the light dynamics (`step`, lib/light.py:59-75, 122-127, 237-253, 300-316) and the
field evaluation (`value_and_gradients`, :176-189, :137-141) run inside the CUDA step kernel.  The
objects here carry the parameters into the scene template, hold the host mirror of the light state,
and keep a numpy `step` / `value_and_gradients` with the reference's semantics for callers that use
a light on its own.  Deviations from reference bugs are listed in DESIGN.md (D1-D3, D16).
"""
from typing import Callable, Iterable, Optional

import numpy as np

from .. import _abi as abi
from .. import scene as S
from ..spaces import Box


class Light(object):
    relative_actions = True
    interpolate_actions = True

    def __init__(self, **kwargs):
        self.observation_space = None
        self.action_space = None

    def step(self, action, time_step: float):
        raise NotImplementedError

    def get_value(self, position: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def get_gradient(self, position: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def value_and_gradients(self, position: np.ndarray):
        return self.get_value(position), self.get_gradient(position)

    def get_state(self):
        raise NotImplementedError

    # ---- scene template / device mirror -----------------------------------------------------
    def _specs(self):
        raise NotImplementedError

    def _state_vector(self):
        return np.asarray(self.get_state(), dtype=np.float64).ravel()

    def _load_state(self, vec):
        raise NotImplementedError


class SinglePositionLight(Light):
    _type = abi.KB_LIGHT_CIRCULAR

    def __init__(self, *, position: np.ndarray = None, bounds=None, action_bounds=None,
                 relative_actions: bool = True, **kwargs):
        super().__init__(**kwargs)
        self._position = np.array((.0, .0)) if position is None else np.asarray(position, dtype=np.float64)
        self._bounds = bounds
        if self._bounds is None:
            self._bounds = np.array([-np.inf, -np.inf]), np.array([np.inf, np.inf])
        self._relative_actions = relative_actions
        self._action_bounds = action_bounds
        if self._action_bounds is None:
            if self._relative_actions:
                self._action_bounds = np.array([-0.01, -0.01]), np.array([.01, .01])
            else:
                self._action_bounds = self._bounds
        self.action_space = Box(*self._action_bounds, dtype=np.float64)
        self.observation_space = Box(*self._bounds, dtype=np.float64)
        self._radius = np.inf

    def step(self, action: np.ndarray, time_step: float):
        if action is None:
            return
        action = np.asarray(action, dtype=np.float64).squeeze()
        action = np.minimum(np.maximum(action, self._action_bounds[0]), self._action_bounds[1])
        if self._relative_actions:
            self._position = self._position + action * time_step
        else:
            self._position = action
        self._position = np.minimum(np.maximum(self._position, self._bounds[0]), self._bounds[1])

    def get_value(self, position: np.ndarray):
        return -1 * np.linalg.norm(position - self._position, axis=1)

    def get_gradient(self, position: np.ndarray):
        return self.value_and_gradients(position)[1]

    def value_and_gradients(self, position: np.ndarray):
        gradients = -1 * (np.asarray(position, dtype=np.float64) - self._position)
        norms = np.linalg.norm(gradients, axis=1)
        safe = np.where(norms == 0.0, 1.0, norms)
        return -1 * norms, gradients / safe[:, None]   # row-wise normalisation (reference bug D2 fixed)

    def get_position(self):
        return self._position

    def get_state(self):
        return self._position

    def _specs(self):
        return [S.LightSpec(self._type, radius=float(self._radius), bounds=self._bounds,
                            action_bounds=self._action_bounds, relative_actions=self._relative_actions)]

    def _load_state(self, vec):
        self._position = np.array(vec[:2], dtype=np.float64)


class CircularGradientLight(SinglePositionLight):
    def __init__(self, radius=.2, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._radius = radius

    def get_value(self, position: np.ndarray):
        return self.value_and_gradients(position)[0]

    def value_and_gradients(self, position: np.ndarray):
        gradient = -1 * (np.asarray(position, dtype=np.float64) - self._position)
        norm_gradient = np.linalg.norm(gradient, axis=1)
        value = np.ones(gradient.shape[0])
        value -= norm_gradient / self._radius
        value = np.maximum(np.minimum(value, 1.), .0)
        value *= 255
        safe = np.where(norm_gradient == 0.0, 1.0, norm_gradient)   # reference: NaN at d == 0 (D3)
        gradient = gradient / safe[:, None]
        gradient[norm_gradient > self._radius] *= .0
        return value, gradient


class MomentumLight(CircularGradientLight):
    interpolate_actions = False
    _type = abi.KB_LIGHT_MOMENTUM

    def __init__(self, velocity=None, max_velocity=None, action_bounds=None, **kwargs):
        super().__init__(**kwargs)
        self._action_bounds = action_bounds
        if self._action_bounds is None:
            self._action_bounds = np.array([-.01, -.01]), np.array([.01, .01])
        self._velocity = np.array([.0, .0]) if velocity is None else np.asarray(velocity, dtype=np.float64)
        self.max_velocity = np.inf if max_velocity is None else max_velocity
        mv = self.max_velocity
        self._obs_bounds = np.r_[self._bounds[0], [-mv, -mv]], np.r_[self._bounds[1], [mv, mv]]
        self.action_space = Box(*self._action_bounds, dtype=np.float64)
        self.observation_space = Box(*self._obs_bounds, dtype=np.float64)

    def step(self, action: np.ndarray, time_step: float):
        if action is not None:
            action = np.asarray(action, dtype=np.float64).squeeze()
            action = np.minimum(np.maximum(action, self._action_bounds[0]), self._action_bounds[1])
            self._velocity = self._velocity + action * time_step
        n = np.linalg.norm(self._velocity)
        if self.max_velocity is not None and n > self.max_velocity:
            self._velocity = self._velocity * (self.max_velocity / n)
        self._position = self._position + self._velocity * time_step
        self._position = np.minimum(np.maximum(self._position, self._bounds[0]), self._bounds[1])

    def get_state(self):
        return np.r_[self._position, self._velocity]

    def _specs(self):
        return [S.LightSpec(self._type, radius=float(self._radius), bounds=self._bounds,
                            action_bounds=self._action_bounds, relative_actions=True,
                            max_velocity=float(self.max_velocity))]

    def _load_state(self, vec):
        self._position = np.array(vec[:2], dtype=np.float64)
        self._velocity = np.array(vec[2:4], dtype=np.float64)


class GradientLight(Light):
    """Linear gradient light.  The reference's get_value/get_gradient only run for N == 2 kilobots
    (lib/light.py:255-260); this implements the evident intent: value_i = pos_i . vec,
    gradient = vec for every kilobot."""
    relative_actions = False
    interpolate_actions = False

    def __init__(self, angle: float = .0):
        super().__init__()
        self._gradient_angle = np.array([angle], dtype=np.float64)
        self._gradient_vec = np.r_[np.cos(angle), np.sin(angle)]
        self._bounds = np.array([-np.pi]), np.array([np.pi])
        self._action_bounds = 2 * np.array([-np.pi]), 2 * np.array([np.pi])
        self.observation_space = Box(*self._bounds, dtype=np.float64)
        self.action_space = Box(*self._action_bounds, dtype=np.float64)

    def step(self, action, time_step):
        if action is None:
            return
        action = np.minimum(np.maximum(np.asarray(action, dtype=np.float64).reshape(1), self._action_bounds[0]),
                            self._action_bounds[1])
        self._gradient_angle = action
        if self._gradient_angle < self._bounds[0]:
            self._gradient_angle = self._gradient_angle + 2 * np.pi
        if self._gradient_angle > self._bounds[1]:
            self._gradient_angle = self._gradient_angle - 2 * np.pi
        self._gradient_vec = np.r_[np.cos(self._gradient_angle), np.sin(self._gradient_angle)]

    def get_value(self, position: np.ndarray):
        return np.asarray(position, dtype=np.float64) @ self._gradient_vec

    def get_gradient(self, position: np.ndarray):
        return np.broadcast_to(self._gradient_vec, np.asarray(position).shape).copy()

    def get_state(self):
        return self._gradient_angle

    def set_angle(self, angle):
        self._gradient_angle = np.array([angle], dtype=np.float64)
        self._gradient_vec = np.r_[np.cos(angle), np.sin(angle)]

    def _specs(self):
        return [S.LightSpec(abi.KB_LIGHT_LINEAR, bounds=(self._bounds[0], self._bounds[1]),
                            action_bounds=(self._action_bounds[0], self._action_bounds[1]), relative_actions=False)]

    def _load_state(self, vec):
        self.set_angle(float(vec[0]))


class CompositeLight(Light):
    def __init__(self, lights: Iterable[Light] = None, reducer: Callable[[np.ndarray, Optional[int]], float] = np.sum):
        super().__init__()
        self._lights = list(lights)
        self._reducer = reducer
        self.observation_space = Box(np.concatenate([l.observation_space.low for l in self._lights]),
                                     np.concatenate([l.observation_space.high for l in self._lights]), dtype=np.float64)
        self.action_space = Box(np.concatenate([l.action_space.low for l in self._lights]),
                                np.concatenate([l.action_space.high for l in self._lights]), dtype=np.float64)
        self._action_dims = [l.action_space.shape[0] for l in self._lights]

    @property
    def lights(self):
        return tuple(self._lights)

    def step(self, action, time_step):
        if action is not None:
            action = np.asarray(action, dtype=np.float64).squeeze()
            for l, ad in zip(self._lights, self._action_dims):
                l.step(action[:ad], time_step)
                action = action[ad:]

    def get_value(self, position: np.ndarray):
        return np.sum(np.array([l.get_value(position) for l in self._lights]), axis=0)

    def get_gradient(self, position: np.ndarray):
        return self.value_and_gradients(position)[1]

    def value_and_gradients(self, position: np.ndarray):
        values, grads = map(np.asarray, zip(*[l.value_and_gradients(position) for l in self._lights]))
        value = np.sum(values, axis=0)
        max_l = np.argmax(values, axis=0)
        return value, grads[max_l, range(np.asarray(position).shape[0])]

    def get_state(self):
        return np.concatenate([np.asarray(l.get_state(), dtype=np.float64).ravel() for l in self._lights])

    def _specs(self):
        out = []
        for l in self._lights:
            out.extend(l._specs())
        return out

    def _load_state(self, vec):
        off = 0
        for l in self._lights:
            n = len(l._state_vector())
            l._load_state(vec[off:off + n])
            off += n


class SmoothGridLight(Light):
    """All-NotImplementedError stub in the reference (lib/light.py:198-215); kept for name parity."""

    def __init__(self):
        super(SmoothGridLight, self).__init__()
