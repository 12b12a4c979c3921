"""Oracle + host code against golden vectors produced by EXECUTING the reference's Python
(tests/golden/make_golden.py -> reference_python.npz): lights, motor model, shape construction.
These pin the Python half of the hot path; the Box2D half has no
reference artefact to pin against (parity unpinned)."""
import os

import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scene as S

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_python.npz"))


def test_circular_light_value_and_gradients(oracle):
    spec = S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=float(G["light_radius"]))
    v, g = oracle.eval_light([spec], G["light_pos"], G["light_pts"])
    assert np.array_equal(v, G["light_value"])
    assert np.array_equal(g, G["light_grad"])


def test_composite_light(oracle):
    specs = [S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=.2), S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=.3)]
    v, g = oracle.eval_light(specs, np.array([-0.1, 0.0, 0.15, 0.05]), G["comp_pts"])
    assert np.array_equal(v, G["comp_value"])
    assert np.array_equal(g, G["comp_grad"])


def test_single_position_light_step(oracle):
    b = G["lstep_bounds"]
    spec = S.LightSpec(abi.KB_LIGHT_CIRCULAR, radius=.2, bounds=(b[0], b[1]), action_bounds=((-.01, -.01), (.01, .01)))
    state = G["lstep_init"].copy()
    traces = []
    for a in G["lstep_actions"]:
        state, tr = oracle.light_step(spec, state, a, 10)
        traces.append(tr)
    assert np.array_equal(np.concatenate(traces), G["lstep_trace"])


def test_momentum_light_step(oracle):
    b = G["lstep_bounds"]
    spec = S.LightSpec(abi.KB_LIGHT_MOMENTUM, radius=.2, bounds=(b[0], b[1]), action_bounds=((-.01, -.01), (.01, .01)),
                       max_velocity=.01)
    state = G["mstep_init"].copy()
    traces = []
    for a in G["mstep_actions"]:
        state, tr = oracle.light_step(spec, state, a, 10)
        traces.append(tr)
    assert np.array_equal(np.concatenate(traces), G["mstep_trace"])


def test_phototaxis_motor_model(oracle):
    """PhototaxisKilobot._loop + Kilobot.step: velocities handed to Box2D, 12 headings x 40 calls."""
    for th, vals, lin, ang in zip(G["photo_angles"], G["photo_light"], G["photo_lin"], G["photo_ang"]):
        feed = np.stack([vals, np.zeros_like(vals), np.zeros_like(vals)], axis=1)
        out = oracle.controller_trace(abi.KB_KILOBOT_PHOTOTAXIS, (0.0, 0.0, float(th)), feed)
        assert np.array_equal(out[:, :2], lin)
        assert np.array_equal(out[:, 2], ang)


def test_simple_phototaxis_velocity(oracle):
    for gvec, lin in zip(G["simple_grads"], G["simple_lin"]):
        out = oracle.controller_trace(abi.KB_KILOBOT_SIMPLE_PHOTOTAXIS, (0.0, 0.0, 0.0), [[1.0, gvec[0], gvec[1]]])
        assert np.array_equal(out[0, :2], lin)


def test_velocity_control_kilobot(oracle):
    """np.cos/np.sin (libm) in the reference vs the shared double sincos: equal after the float32 cast."""
    for th, vv, lin, ang in zip(G["vel_angles"], G["vel_actions"], G["vel_lin"], G["vel_ang"]):
        out = oracle.controller_trace(abi.KB_KILOBOT_VELOCITY, (0.0, 0.0, float(th)), [[0.0, 0.0, 0.0]], velocity=vv)
        assert np.array_equal(out[0, :2], lin)
        assert out[0, 2] == ang


def test_light_sensor_position(oracle):
    x, y, a = G["sensor_pose"]
    assert np.array_equal(oracle.sensor_pos(x, y, a), G["sensor_pos"])


@pytest.mark.parametrize("name,template", [("Triangle", S.TRIANGLE_TEMPLATE), ("LForm", S.LFORM_TEMPLATE),
                                           ("TForm", S.TFORM_TEMPLATE), ("CForm", S.CFORM_TEMPLATE)])
def test_polygon_vertices_handed_to_box2d(name, template):
    local = S.polygon_local_vertices(template, .15, .15)
    assert np.array_equal(local, G["body_%s_local" % name])
    fixtures = S.polygon_fixtures(local)
    got = np.array([f.vertices for f in fixtures])
    assert np.array_equal(got, G["body_%s_verts" % name])
    assert [f.density for f in fixtures] == [G["body_%s_material" % name][0]] * len(fixtures)
    assert fixtures[0].friction == G["body_%s_material" % name][1]


def test_quad_circle_and_kilobot_fixtures():
    q = S.box_fixture(.15, .15)
    assert np.array_equal(np.array([q.hx, q.hy], np.float32), G["body_Quad_box"])
    c = S.circle_fixture(.075, S.OBJECT_DENSITY, S.OBJECT_FRICTION, S.OBJECT_RESTITUTION)
    assert np.float32(c.radius) == G["body_Circle_radius"]
    k = S.kilobot_body(abi.KB_KILOBOT_PHOTOTAXIS).fixtures[0]
    assert np.float32(k.radius) == G["kilobot_radius_b2"]
    assert np.array_equal([k.density, k.friction, k.restitution], G["kilobot_material"])
    assert np.array_equal(G["body_Quad_damping"], [S.LINEAR_DAMPING, S.ANGULAR_DAMPING])
    # Body.__init__ position scaling: float64 * 25 -> float32
    assert np.array_equal(np.array([25.0 * .3, 25.0 * -.2], np.float64).astype(np.float32), G["body_Quad_pos"])
