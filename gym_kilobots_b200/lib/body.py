"""Pushable objects -- same constructors and accessors as gym_kilobots/lib/body.py.

The reference creates a Box2D body in `Body.__init__` (lib/body.py:32-38) and fixtures in the
subclasses (:136-142, :187-192, :245-251).  Here the constructors only RECORD a `BodySpec` with the
world recorder (`gym_kilobots_b200.lib.world.World`); the simulation state lives on the GPU and the
pose getters read the host mirror the owning env refreshes after every reset/step.
"""
import numpy as np

from .. import _abi as abi
from .. import scene as S

_world_scale = S.WORLD_SCALE  # lib/body.py:7


class Body:
    _density = S.OBJECT_DENSITY
    _friction = S.OBJECT_FRICTION
    _restitution = S.OBJECT_RESTITUTION
    _linear_damping = S.LINEAR_DAMPING
    _angular_damping = S.ANGULAR_DAMPING
    _kind = abi.KB_BODY_OBJECT

    def __init__(self, world, position=None, orientation=None):
        if self.__class__ == Body:
            raise NotImplementedError('Abstract class Body cannot be instantiated.')
        self._color = np.array((93, 133, 195))
        self._highlight_color = np.array((238, 80, 62))
        if position is None:
            position = [.0, .0]
        position = np.asarray(position, dtype=np.float64)
        if orientation is None:
            orientation = .0
        self._world = world
        self._init_pose = np.array([position[0], position[1], float(orientation)], dtype=np.float64)
        self._fixtures = []
        self._index = world._register(self)

    # ---- spec -------------------------------------------------------------------------------
    def _spec(self):
        return S.BodySpec(self._kind, list(self._fixtures), self._linear_damping, self._angular_damping)

    # ---- raw state (b2 units, float32) from the env's host mirror ------------------------------
    def _raw(self):
        return self._world._raw_pose(self)

    @property
    def width(self):
        raise NotImplementedError

    @property
    def height(self):
        raise NotImplementedError

    def get_position(self):
        x, y = self._raw()[:2]
        return np.asarray((x, y), dtype=np.float64) / _world_scale

    def set_position(self, position):
        pose = np.array(self.get_pose(), dtype=np.float64)
        pose[:2] = np.asarray(position, dtype=np.float64)
        self._world._set_pose(self, pose)

    def get_orientation(self):
        return float(self._raw()[2])

    def set_orientation(self, orientation):
        pose = np.array(self.get_pose(), dtype=np.float64)
        pose[2] = orientation
        self._world._set_pose(self, pose)

    def get_pose(self):
        position = self.get_position()
        return tuple((*position, self.get_orientation()))

    def set_pose(self, pose):
        self._world._set_pose(self, np.asarray(pose, dtype=np.float64))

    def get_state(self):
        return self.get_pose()

    def _rot(self):
        r = self._raw()   # (x, y, angle, sin, cos): b2Body::m_xf.q as synchronised on the device
        return np.float32(r[4]), np.float32(r[3])

    def get_world_point(self, point):
        p = (_world_scale * np.asarray(point, dtype=np.float64)).astype(np.float32)
        x, y = self._raw()[:2]
        c, s = self._rot()
        wx = (c * p[0] - s * p[1]) + np.float32(x)
        wy = (s * p[0] + c * p[1]) + np.float32(y)
        return np.asarray((wx, wy), dtype=np.float64) / _world_scale

    def get_local_point(self, point):
        p = (_world_scale * np.asarray(point, dtype=np.float64)).astype(np.float32)
        x, y = self._raw()[:2]
        c, s = self._rot()
        px, py = p[0] - np.float32(x), p[1] - np.float32(y)
        return np.asarray((c * px + s * py, -s * px + c * py), dtype=np.float64) / _world_scale

    def get_local_orientation(self, angle):
        return angle - self.get_orientation()

    def get_local_pose(self, pose):
        return tuple((*self.get_local_point(pose[:2]), self.get_local_orientation(pose[2])))

    def collides_with(self, other):
        return self._world._touching(self, other)

    @property
    def color(self):
        return self._color

    @color.setter
    def color(self, color):
        self._color = np.clip(np.asarray(color, dtype=np.int32), 0, 255)

    @property
    def highlight_color(self):
        return self._highlight_color

    @highlight_color.setter
    def highlight_color(self, color):
        self._highlight_color = np.clip(np.asarray(color, dtype=np.int32), 0, 255)


class Quad(Body):
    def __init__(self, width, height, **kwargs):
        super().__init__(**kwargs)
        self._width = width
        self._height = height
        self._fixtures = [S.box_fixture(width, height, self._density, self._friction, self._restitution)]

    @property
    def width(self):
        return self._width

    @property
    def height(self):
        return self._height

    @property
    def vertices(self):
        hx, hy = self._width / 2, self._height / 2
        local = [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
        return np.asarray([[self.get_world_point(v) for v in local]])

    def get_width(self):
        return self._width

    def get_height(self):
        return self._height


class CornerQuad(Quad):
    pass


class Circle(Body):
    def __init__(self, radius, **kwargs):
        super().__init__(**kwargs)
        self._radius = radius
        self._fixtures = [S.circle_fixture(radius, self._density, self._friction, self._restitution)]

    @property
    def width(self):
        return 2 * self._radius

    @property
    def height(self):
        return 2 * self._radius

    @property
    def vertices(self):
        return np.array([[self.get_position()]])

    def get_radius(self):
        return self._radius


class Polygon(Body):
    def __init__(self, width: float, height: float, **kwargs):
        super().__init__(**kwargs)
        self._width = width
        self._height = height
        self.__local_vertices = S.polygon_local_vertices(self._shape_vertices(), width, height)
        self.__local_vertices.setflags(write=False)
        self._fixtures = S.polygon_fixtures(self.__local_vertices, self._density, self._friction, self._restitution)

    @property
    def width(self):
        return self._width

    @property
    def height(self):
        return self._height

    @property
    def vertices(self):
        return np.array([[self.get_world_point(v) for v in vertices] for vertices in self.__local_vertices])

    @property
    def local_vertices(self):
        return self.__local_vertices

    @staticmethod
    def _shape_vertices() -> np.ndarray:
        raise NotImplementedError


class Triangle(Polygon):
    @staticmethod
    def _shape_vertices():
        return np.array(S.TRIANGLE_TEMPLATE)


class LForm(Polygon):
    @staticmethod
    def _shape_vertices():
        return np.array(S.LFORM_TEMPLATE)


class TForm(Polygon):
    @staticmethod
    def _shape_vertices():
        return np.array(S.TFORM_TEMPLATE)


class CForm(Polygon):
    @staticmethod
    def _shape_vertices():
        return np.array(S.CFORM_TEMPLATE)
