"""Counter-based sampling (gym_kilobots_b200/philox.py) and the vectorised BASELINE scenarios built on it."""
import numpy as np

from gym_kilobots_b200 import philox as PX
from gym_kilobots_b200 import scenarios as SC


def _hex(t):
    return [int(x) for x in t]


def test_philox4x32_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    assert _hex(PX.philox4x32(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _hex(PX.philox4x32(f, f, f, f, f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _hex(PX.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draws_depend_on_env_id_only():
    a = PX.EnvRng(7, np.arange(100))
    b = PX.EnvRng(7, np.arange(40, 60))
    ua, ub = a.uniform2(PX.STREAM_LIGHT, 3), b.uniform2(PX.STREAM_LIGHT, 3)
    assert np.array_equal(ua[0][40:60], ub[0]) and np.array_equal(ua[1][40:60], ub[1])
    na = a.normal2(PX.STREAM_KILOBOT_POS, np.arange(5)[None, :])
    nb = b.normal2(PX.STREAM_KILOBOT_POS, np.arange(5)[None, :])
    assert np.array_equal(na[0][40:60], nb[0])
    u = a.uniform2(PX.STREAM_OBJECT, np.arange(2000)[None, :])[0]
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3
    z = a.normal2(PX.STREAM_SWARM, np.arange(2000)[None, :])[0]
    assert abs(z.mean()) < 1e-2 and abs(z.std() - 1.0) < 1e-2


def test_a_slice_of_a_batch_is_the_batch_of_the_slice():
    """What makes 1 GPU and 8 GPUs simulate the same environments: rank r builds envs [rE, (r+1)E) on its own."""
    for build, n in ((SC.c3_shapes, 48), (SC.c5_small, 4096), (SC.c4_swarm, 6)):
        whole = build(n)
        part = build(n // 3, env_offset=n // 3)
        sl = slice(n // 3, 2 * (n // 3))
        assert np.array_equal(whole.body_pose[sl], part.body_pose), build.__name__
        assert np.array_equal(whole.light_state[sl], part.light_state), build.__name__


def test_c5_spawn_follows_survey_8d():
    sc = SC.c5_small(20000)
    p = sc.body_pose[:, 1:, :2]
    d = np.hypot(p[:, :, None, 0] - p[:, None, :, 0], p[:, :, None, 1] - p[:, None, :, 1]) + np.eye(4)[None]
    assert d.min() >= 2 * 0.0165 + 1e-3 - 1e-12          # rejection-separated
    off = p - sc.light_state[:, None, :]
    inner = np.all(np.abs(sc.light_state) < np.array([0.8, 0.55]), axis=1)   # away from the clip at the table edge
    s = off[inner].std()
    assert 0.03 < s < 0.045                               # N(L0, 0.03^2), widened a little by the rejection


def test_c4_lattice():
    sc = SC.c4_swarm(3)
    assert sc.body_pose.shape == (3, 1024, 3) and sc.scenes[0].num_kilobots == 1024
    p = sc.body_pose[0, :, :2].reshape(32, 32, 2)
    assert np.allclose(np.diff(p[:, :, 0], axis=1), 0.036, atol=0.0021)
    assert np.allclose(np.diff(p[:, :, 1], axis=0), 0.036, atol=0.0021)
    assert np.abs(p).max() < 0.6
