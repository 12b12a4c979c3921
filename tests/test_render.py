"""Off-screen rasteriser (include/kb_b200.h kb_render): the frame KilobotsEnv.render draws
through kb_rendering.KilobotsViewer (kilobots_env.py:221-275).  pygame is not installable here, so the oracle's
restatement is pinned on the draw semantics it cites (colours, radii, painter's order) at analytically known
pixels, and the CUDA rasteriser is compared with the oracle's byte for byte."""
import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC

W, H = 1200, 900         # the reference's screen_size (kilobots_env.py:22): 600 px per metre on the 2.0 x 1.5 m table


def _px(x, y):
    return int((x + 1.0) * 600), int((0.75 - y) * 600)      # (column, row) of world point (x, y) metres


def test_oracle_render_draw_semantics(oracle):
    sc = SC.c1_single_env(1, seed=0)
    sc.body_pose[0, 0] = (0.5, 0.3, 0.0)                    # Quad 0.15 x 0.15
    sc.body_pose[0, 1:, :2] = np.array([-0.8, -0.6]) + 0.05 * np.arange(10)[:, None] * np.array([1.0, 0.0])
    sc.body_pose[0, 1:, 2] = 0.0
    sc.light_state[0] = (0.0, 0.0)
    ob = oracle.OracleBatch(sc.scenes, 1, sc.env_scene, sc.max_contacts)
    ob.reset(sc.body_pose, sc.light_state)
    img = ob.render((0,), W, H)[0]
    assert img.shape == (H, W, 3) and img.dtype == np.uint8
    at = lambda x, y: tuple(int(v) for v in img[_px(x, y)[1], _px(x, y)[0]])
    assert at(0.9, 0.6) == (255, 255, 255)                  # bare table (kilobots_env.py:254-255)
    assert tuple(img[0, 600]) == (0, 0, 0) and tuple(img[450, 0]) == (0, 0, 0)   # border polyline (:256-257)
    assert at(0.5, 0.3) == (93, 133, 195)                   # object colour (lib/body.py:21)
    assert at(0.5 + 0.08, 0.3) == (255, 255, 255)           # just outside the 0.075 half-width
    kx, ky = ob.bodies()[0, 1 + 3, 8:10].astype(np.float64) / 25.0   # kilobot 3, heading 0
    assert at(kx + 0.006, ky) == (255, 255, 255)            # heading line towards +x (lib/kilobot.py:137-145)
    assert at(kx, ky + 0.008) == (150, 150, 150)            # body disc (:132)
    assert at(kx, ky + 0.016) == (100, 100, 100)            # ring, width .005 inwards from r + .002 (:133-134)
    # translucent light disc (lib/light.py:194-195): (255,255,30) at alpha 150 over white
    blend = tuple(int(255 * 150 / 255 + 255 * (1 - 150 / 255) + 0.5) if c == 255 else
                  int(30 * 150 / 255 + 255 * (1 - 150 / 255) + 0.5) for c in (255, 255, 30))
    assert at(0.1, 0.1) == blend and at(0.25, 0.0) == (255, 255, 255)


@pytest.mark.gpu
def test_kernel_render_matches_oracle(oracle, native):
    sc = SC.pushing_yard(6, light="composite")              # every object shape, two lights
    ob = oracle.OracleBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts, threads=4)
    nb = native.NativeBatch(sc.scenes, sc.num_envs, sc.env_scene, sc.max_contacts)
    ob.reset(sc.body_pose, sc.light_state)
    nb.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, sc.num_envs, 5)
    for a in acts:
        ob.step(a)
        nb.step(a)
    ids = (0, 3, 5)
    io = ob.render(ids, 400, 300)
    im = nb.render(ids, 400, 300).cpu().numpy()
    assert io.shape == im.shape == (3, 300, 400, 3)
    assert np.array_equal(io, im), "pixels differ at %s" % (np.argwhere(io != im)[:3],)
    assert len(np.unique(io.reshape(-1, 3), axis=0)) >= 5   # table, border, object, kilobot, light blend ...


@pytest.mark.gpu
def test_env_render_surfaces(native):
    from gym_kilobots_b200.envs import KilobotsVecEnv, QuadAssemblyKilobotsEnv
    vec = KilobotsVecEnv(SC.c2_quad_assembly(8, seed=1))
    vec.reset()
    frames = vec.render(env_ids=(1, 2), width=320, height=240)
    assert frames.shape == (2, 240, 320, 3) and frames.dtype == np.uint8
    env = QuadAssemblyKilobotsEnv(seed=0)
    env.reset()
    frame = env.render(mode='rgb_array')
    assert frame.shape == (900, 1200, 3) and (frame == np.array([93, 133, 195], np.uint8)).all(-1).any()
    with pytest.raises(NotImplementedError):
        env.render(mode='human')
