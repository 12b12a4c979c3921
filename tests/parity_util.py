"""Shared helpers: drive the oracle and the CUDA path with identical inputs and compare."""
import numpy as np


def assert_same_state(ob, nb, where, atol=0.0):
    """Bit-level agreement (== on floats, so -0.0 == +0.0) of everything observable."""
    bo, bn = ob.bodies(), nb.bodies()
    if atol == 0.0:
        bad = np.argwhere(~((bo == bn) | (np.isnan(bo) & np.isnan(bn))))
    else:
        bad = np.argwhere(~(np.abs(bo - bn) <= atol))
    assert bad.size == 0, "%s: body state differs at (env, body, field) %s: oracle %r kernel %r" % (
        where, bad[0], bo[tuple(bad[0])], bn[tuple(bad[0])])
    (po, no), (pn, nn) = ob.contacts(), nb.contacts()
    assert np.array_equal(no, nn), "%s: contact counts differ, first env %s: oracle %s kernel %s" % (
        where, np.argwhere(no != nn)[0], no[no != nn][:4], nn[no != nn][:4])
    assert np.array_equal(po, pn), "%s: contact lists differ at %s" % (where, np.argwhere(po != pn)[0])
    fo, fn = ob.proxies(), nb.proxies()
    assert np.array_equal(fo, fn), "%s: fat AABBs differ at %s" % (where, np.argwhere(fo != fn)[0])
    io, im = ob.impulses(), nb.impulses()
    assert np.array_equal(io, im), "%s: stored impulses differ at %s" % (where, np.argwhere(io != im)[0])
    (co, lo), (cn, ln) = ob.controllers(), nb.controllers()
    assert np.array_equal(co, cn), "%s: controller state differs at %s" % (where, np.argwhere(co != cn)[0])
    assert np.array_equal(lo, ln), "%s: light state differs" % where


def assert_same_obs(oo, on, where):
    for k in ("kilobots", "objects", "light", "reward", "done", "status"):
        a, b = oo[k], on[k]
        assert a.shape == b.shape, (where, k, a.shape, b.shape)
        assert np.array_equal(a, b), "%s: %s differs at %s" % (where, k, np.argwhere(a != b)[0])


def run_parity(kbo, native, scenario, steps, actions=None, mode=None, check_every=1):
    from gym_kilobots_b200 import scenarios as SC
    ob = kbo.OracleBatch(scenario.scenes, scenario.num_envs, scenario.env_scene, scenario.max_contacts, threads=8)
    nb = native.NativeBatch(scenario.scenes, scenario.num_envs, scenario.env_scene, scenario.max_contacts)
    assert np.array_equal(ob.mass_data(), nb.mass_data()), "mass data differ"
    ob.reset(scenario.body_pose, scenario.light_state)
    nb.reset(scenario.body_pose, scenario.light_state)
    assert_same_state(ob, nb, "%s after reset" % scenario.name)
    if actions is None:
        actions = SC.random_actions(scenario, scenario.num_envs, steps)
    for t in range(steps):
        oo = ob.step(actions[t], mode)
        on = nb.step(actions[t], mode)
        if t % check_every == 0 or t == steps - 1:
            assert_same_obs(oo, on, "%s step %d" % (scenario.name, t))
            assert_same_state(ob, nb, "%s step %d" % (scenario.name, t))
    co, cn = ob.counters(), nb.counters()
    for idx, name in ((0, "substeps"), (1, "contacts"), (2, "points"), (4, "pos_iters"), (5, "toi_events"), (7, "islands")):
        assert np.array_equal(co[:, idx], cn[:, idx]), "%s: counter %s differs: %s vs %s" % (
            scenario.name, name, co[:, idx][:4], cn[:, idx][:4])
    return ob, nb
