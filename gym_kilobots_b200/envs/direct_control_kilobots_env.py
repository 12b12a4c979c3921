"""DirectControlKilobotsEnv (gym_kilobots/envs/direct_control_kilobots_env.py:8-29): per-kilobot
actions routed to Kilobot.set_action, then the base step with action=None."""
import numpy as np

from .. import _abi as abi
from ..spaces import Box
from .kilobots_env import KilobotsEnv


class DirectControlKilobotsEnv(KilobotsEnv):
    def __init__(self, **kwargs):
        super(DirectControlKilobotsEnv, self).__init__(**kwargs)

    @property
    def action_space(self):
        as_low = np.array([kb.action_space.low for kb in self._kilobots])
        as_high = np.array([kb.action_space.high for kb in self._kilobots])
        return Box(as_low, as_high, dtype=np.float64)

    def step(self, actions: np.ndarray):
        if self._batch is None:
            raise RuntimeError("step called before reset()")
        state = self.get_state()
        if actions is not None:
            for kb, a in zip(self.kilobots, actions):
                kb.set_action(a)
            act = np.asarray(actions, dtype=np.float64).reshape(1, -1)
        else:
            for kb in self.kilobots:
                kb.set_action(None)
            act = None
        self._step_batch(act, abi.KB_ACTION_KILOBOTS)
        self._KilobotsEnv__sim_steps += self._steps_per_action
        self._sync_mirror()
        next_state = self.get_state()
        return (self.get_observation(), self.get_reward(state, None, next_state),
                self.has_finished(next_state, None), self.get_info(next_state, None))
