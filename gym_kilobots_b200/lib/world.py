"""World recorder -- stands in for the `b2World` the reference passes to every constructor
(gym_kilobots/envs/kilobots_env.py:45; lib/body.py:18,32).

Bodies register themselves here in construction order.  The owning env turns the registered bodies
into a `SceneSpec` + initial poses, pushes them to the GPU, and refreshes `_mirror` (raw float32
body state in Box2D units) after every reset/step so that the per-object getters of lib/body.py keep
working without a device round trip per call.
"""
import numpy as np


class World:
    def __init__(self):
        self._bodies = []      # registration order
        self._slot = {}        # id(body) -> body slot in the batch (objects first, then kilobots)
        self._mirror = None    # [B, 12] raw body state of env 0
        self._contacts = None  # (pairs [C,4], count, proxy->slot map)
        self._env = None

    # -- called by Body.__init__ ---------------------------------------------------------------
    def _register(self, body):
        self._bodies.append(body)
        self._mirror = None
        return len(self._bodies) - 1

    def _unregister_all(self):
        self._bodies = []
        self._slot = {}
        self._mirror = None
        self._contacts = None

    def _unregister(self, bodies):
        drop = {id(b) for b in bodies}
        self._bodies = [b for b in self._bodies if id(b) not in drop]
        self._mirror = None

    # -- used by the getters -------------------------------------------------------------------
    def _raw_pose(self, body):
        """(x, y, angle, sin, cos) in Box2D units as b2Body::GetPosition/GetAngle/m_xf.q would return."""
        if self._mirror is None or id(body) not in self._slot:
            p = body._init_pose
            x, y = np.float32(25.0 * p[0]), np.float32(25.0 * p[1])
            a = np.float32(p[2])
            return x, y, a, np.float32(np.sin(np.float64(a))), np.float32(np.cos(np.float64(a)))
        r = self._mirror[self._slot[id(body)]]
        return r[8], r[9], r[2], r[10], r[11]

    def _set_pose(self, body, pose):
        if self._env is not None and self._mirror is not None and id(body) in self._slot:
            self._env._set_body_pose(self._slot[id(body)], pose)
        else:
            body._init_pose = np.asarray(pose, dtype=np.float64).copy()

    def _touching(self, a, b):
        """Body.collides_with (lib/body.py:87-90): True if a touching contact joins the two bodies."""
        if self._contacts is None or id(a) not in self._slot or id(b) not in self._slot:
            return None
        pairs, count, proxy_slot = self._contacts
        sa, sb = self._slot[id(a)], self._slot[id(b)]
        for k in range(count):
            pa, pb, touching, _ = pairs[k]
            if touching and {proxy_slot[pa], proxy_slot[pb]} == {sa, sb}:
                return True
        return None
