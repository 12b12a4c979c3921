#!/usr/bin/env python
"""Generate tests/golden/reference_python.npz by EXECUTING the reference's own Python code
(/root/reference/gym_kilobots/lib/{light,body,kilobot}.py).

pybox2d and gym are not installable here, so the two imports are satisfied by tiny stand-ins:
  * `gym.spaces.Box`  -> a plain container (the reference only stores it);
  * `Box2D`           -> a RECORDING fake: b2Vec2 with float32 x/y and float32 `/` `*` (as pybox2d's
                         SWIG operators), a world whose CreateDynamicBody returns a body that records
                         every CreatePolygonFixture / CreateCircleFixture call and every
                         linearVelocity / angularVelocity / linearDamping assignment, and implements
                         GetWorldVector / GetWorldPoint with float32 b2Mul arithmetic.
Everything numpy/float64 the reference computes around Box2D is therefore the reference's own
arithmetic; what is recorded is exactly what it would hand to Box2D.  Run from the repo root in the
build container (needs /root/reference); the .npz is committed.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_python.npz")


def install_fakes():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high, dtype=None, shape=None):
            self.low, self.high, self.dtype = np.asarray(low), np.asarray(high), dtype
            self.shape = self.low.shape

    class Env:
        pass

    spaces.Box = Box
    gym.spaces = spaces
    gym.Env = Env
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces

    b2 = types.ModuleType("Box2D")
    f32 = np.float32

    class b2Vec2:
        def __init__(self, *a):
            if len(a) == 1:
                a = tuple(a[0])
            self.x, self.y = f32(a[0]), f32(a[1])

        def __truediv__(self, s):
            return b2Vec2(self.x / f32(s), self.y / f32(s))

        def __mul__(self, s):
            return b2Vec2(self.x * f32(s), self.y * f32(s))

        __rmul__ = __mul__

        def __iter__(self):
            return iter((float(self.x), float(self.y)))

        def __len__(self):
            return 2

        def __getitem__(self, i):
            return float((self.x, self.y)[i])

    class b2PolygonShape:
        def __init__(self, vertices=None, box=None):
            self.vertices_in = vertices
            self.box = box

    class FakeBody:
        def __init__(self, position, angle, linearDamping, angularDamping):
            self.position_in = (f32(position.x), f32(position.y))
            self.angle = float(f32(angle))   # pybox2d returns a Python float holding the float32 value
            self.linearDamping = linearDamping
            self.angularDamping = angularDamping
            self.fixtures = []
            self.log = []
            a = np.float64(self.angle)
            self.s, self.c = f32(np.sin(a)), f32(np.cos(a))   # b2Rot::Set, correctly rounded

        @property
        def position(self):
            return b2Vec2(*self.position_in)

        def __setattr__(self, k, v):
            if k in ("linearVelocity", "angularVelocity") or (k == "linearDamping" and hasattr(self, "log")):
                val = (float(v.x), float(v.y)) if isinstance(v, b2Vec2) else float(np.float32(v))
                self.log.append((k, val))
            object.__setattr__(self, k, v)

        def CreatePolygonFixture(self, shape=None, box=None, **kw):
            if box is not None:
                self.fixtures.append(("box", (float(box.x), float(box.y)), kw))
            else:
                self.fixtures.append(("poly", [tuple(map(float, v)) for v in shape.vertices_in], kw))
            return self.fixtures[-1]

        def CreateCircleFixture(self, radius=None, **kw):
            self.fixtures.append(("circle", float(np.float32(radius)), kw))
            return self.fixtures[-1]

        def GetWorldVector(self, v):
            v = v if isinstance(v, b2Vec2) else b2Vec2(*v)
            return b2Vec2(self.c * v.x - self.s * v.y, self.s * v.x + self.c * v.y)

        def GetWorldPoint(self, v):
            v = v if isinstance(v, b2Vec2) else b2Vec2(*v)
            px, py = self.position_in
            return b2Vec2((self.c * v.x - self.s * v.y) + px, (self.s * v.x + self.c * v.y) + py)

    class b2World:
        def __init__(self, **kw):
            self.bodies = []

        def CreateDynamicBody(self, position=None, angle=0.0, linearDamping=0.0, angularDamping=0.0):
            b = FakeBody(position, angle, linearDamping, angularDamping)
            self.bodies.append(b)
            return b

        def DestroyBody(self, b):
            pass

    b2.b2Vec2, b2.b2PolygonShape, b2.b2World, b2.b2Body = b2Vec2, b2PolygonShape, b2World, FakeBody
    b2.b2ChainShape = object
    sys.modules["Box2D"] = b2
    return b2


def main():
    b2 = install_fakes()
    sys.path.insert(0, REF)
    # import the lib modules directly (the package __init__ of envs would need scipy/yaml, irrelevant here)
    from gym_kilobots.lib import body as rbody
    from gym_kilobots.lib import kilobot as rkb
    from gym_kilobots.lib import light as rlight

    out = {}
    rng = np.random.default_rng(20261018)

    # ---- lights ---------------------------------------------------------------------------
    pts = rng.uniform(-0.5, 0.5, size=(64, 2))
    L = rlight.CircularGradientLight(radius=.2, position=np.array([0.05, -0.1]))
    v, g = L.value_and_gradients(pts.copy())
    out["light_pts"], out["light_pos"], out["light_radius"] = pts, np.array([0.05, -0.1]), np.array(.2)
    out["light_value"], out["light_grad"] = v, g
    # SinglePositionLight.step: 3 env-steps x 10 sub-steps with clipping at bounds
    bounds = (np.array([-1.1, -0.825]), np.array([1.1, 0.825]))
    L2 = rlight.CircularGradientLight(radius=.2, position=np.array([1.095, -0.8]), bounds=bounds,
                                      action_bounds=(np.array([-.01, -.01]), np.array([.01, .01])))
    acts = np.array([[0.5, -0.004], [0.0031, -0.5], [-0.007, 0.002]])
    trace = []
    for a in acts:
        for _ in range(10):
            L2.step(a.copy(), 0.1)
            trace.append(np.array(L2.get_state()).copy())
    out["lstep_actions"], out["lstep_init"], out["lstep_trace"] = acts, np.array([1.095, -0.8]), np.array(trace)
    out["lstep_bounds"] = np.array(bounds)
    # MomentumLight.step
    L3 = rlight.MomentumLight(position=np.array([0.1, 0.2]), velocity=np.array([0.006, -0.008]), max_velocity=.01,
                              radius=.2, bounds=bounds, action_bounds=(np.array([-.01, -.01]), np.array([.01, .01])))
    trace = []
    macts = np.array([[0.01, 0.01], [-0.5, 0.003]])
    for a in macts:
        for _ in range(10):
            L3.step(a.copy(), 0.1)
            trace.append(np.array(L3.get_state()).copy())
    out["mstep_actions"], out["mstep_init"], out["mstep_trace"] = macts, np.array([0.1, 0.2, 0.006, -0.008]), np.array(trace)
    # CompositeLight of two circular lights
    CA = rlight.CircularGradientLight(radius=.2, position=np.array([-0.1, 0.0]))
    CB = rlight.CircularGradientLight(radius=.3, position=np.array([0.15, 0.05]))
    comp = rlight.CompositeLight([CA, CB])
    cpts = rng.uniform(-0.3, 0.3, size=(48, 2))
    cv, cg = comp.value_and_gradients(cpts.copy())
    out["comp_pts"], out["comp_value"], out["comp_grad"] = cpts, cv, cg

    # ---- body construction: what the reference hands to Box2D -------------------------------
    world = b2.b2World()
    shapes = {"Quad": lambda: rbody.Quad(width=.15, height=.15, world=world, position=(.3, -.2), orientation=.4),
              "Circle": lambda: rbody.Circle(radius=.075, world=world, position=(.1, .1)),
              "Triangle": lambda: rbody.Triangle(width=.15, height=.15, world=world),
              "LForm": lambda: rbody.LForm(width=.15, height=.15, world=world),
              "TForm": lambda: rbody.TForm(width=.15, height=.15, world=world),
              "CForm": lambda: rbody.CForm(width=.15, height=.15, world=world)}
    for name, mk in shapes.items():
        obj = mk()
        fx = obj._body.fixtures
        out["body_%s_pos" % name] = np.array(obj._body.position_in, dtype=np.float32)
        out["body_%s_damping" % name] = np.array([obj._body.linearDamping, obj._body.angularDamping])
        out["body_%s_material" % name] = np.array([fx[0][2]["density"], fx[0][2]["friction"], fx[0][2]["restitution"]])
        if fx[0][0] == "poly":
            out["body_%s_verts" % name] = np.array([f[1] for f in fx], dtype=np.float64)
        elif fx[0][0] == "box":
            out["body_%s_box" % name] = np.array(fx[0][1], dtype=np.float32)
        else:
            out["body_%s_radius" % name] = np.array(fx[0][1], dtype=np.float32)
        if hasattr(obj, "local_vertices"):
            out["body_%s_local" % name] = np.array(obj.local_vertices)

    # ---- kilobot controllers -------------------------------------------------------------
    def velocity_log(body):
        lin = [v for k, v in body.log if k == "linearVelocity"]
        ang = [v for k, v in body.log if k == "angularVelocity"]
        return lin, ang

    angles = rng.uniform(-3.0, 3.0, size=12).astype(np.float32)
    light_values = rng.uniform(0, 255, size=(12, 40))
    lin_all, ang_all = [], []
    for th, vals in zip(angles, light_values):
        kb = rkb.PhototaxisKilobot(world, position=(0.0, 0.0), orientation=float(th))
        kb._body.log.clear()
        for val in vals:
            kb.set_light_value_and_gradient(float(val), np.zeros(2))
            kb.step(0.1)
        lin, ang = velocity_log(kb._body)
        lin_all.append(lin)
        ang_all.append(ang)
    out["photo_angles"], out["photo_light"] = angles, light_values
    out["photo_lin"], out["photo_ang"] = np.array(lin_all, dtype=np.float32), np.array(ang_all, dtype=np.float32)
    kb = rkb.PhototaxisKilobot(world, position=(0.3, -0.2), orientation=0.7)
    out["sensor_pose"] = np.array([0.3, -0.2, 0.7])
    out["sensor_pos"] = np.array(kb.light_sensor_pos())
    out["kilobot_radius_b2"] = np.array(kb._body.fixtures[0][1], dtype=np.float32)
    out["kilobot_material"] = np.array([kb._body.fixtures[0][2][k] for k in ("density", "friction", "restitution")])

    grads = np.concatenate([rng.normal(size=(10, 2)) / 1.0, np.zeros((1, 2)), rng.normal(size=(5, 2)) * 1e-3])
    lin_all = []
    for gvec in grads:
        kb = rkb.SimplePhototaxisKilobot(world, position=(0.0, 0.0))
        kb._body.log.clear()
        kb.set_light_value_and_gradient(1.0, gvec.copy())
        kb.step(0.1)
        lin_all.append(velocity_log(kb._body)[0][0])
    out["simple_grads"], out["simple_lin"] = grads, np.array(lin_all, dtype=np.float32)

    vels = np.stack([rng.uniform(0, 0.01, 10), rng.uniform(-1.5, 1.5, 10)], axis=1)
    vangles = rng.uniform(-3, 3, 10).astype(np.float32)
    lin_all, ang_all = [], []
    for th, vv in zip(vangles, vels):
        kb = rkb.SimpleVelocityControlKilobot(world, velocity=[1e-9, 0.0], position=(0.0, 0.0), orientation=float(th))
        kb.set_action(vv.copy())
        kb._body.log.clear()
        kb.step(0.1)
        lin, ang = velocity_log(kb._body)
        lin_all.append(lin[0])
        ang_all.append(ang[0])
    out["vel_actions"], out["vel_angles"] = vels, vangles
    out["vel_lin"], out["vel_ang"] = np.array(lin_all, dtype=np.float32), np.array(ang_all, dtype=np.float32)

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "with", len(out), "arrays")


if __name__ == "__main__":
    main()
