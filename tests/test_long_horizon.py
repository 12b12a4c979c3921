"""Long-horizon agreement (BASELINE.json north_star: "statistical agreement of rewards over long horizons";
SURVEY.md 7.4: >= 1024 envs, >= 1000 env-steps -- the reference's own usage is 2000 steps, gym_kilobots/test.py:19-25).

(1) CUDA path vs the oracle with the SAME trigonometry: bit-identical observations, rewards and final state over 1000
    env-steps (10 000 world steps) of 1024 QuadAssembly envs -- the parity contract has no horizon.
(2) CUDA path vs the oracle built with libm sinf / cosf (what Box2D itself calls; 1 ulp different on ~1.3 % of the
    arguments): trajectories separate chaotically, so this measures the HORIZON of pose agreement (first env-step at
    which a pose differs by more than 1e-4 m / rad) and the STATISTICAL agreement of the episode returns (difference of
    means in standard errors, two-sample Kolmogorov-Smirnov test).
(3) soak: no status bit (capacity overflow, non-finite pose) in any env.
The report goes to gpurun_out/long_horizon.json (committed as profiles/long_horizon_r02.json)."""
import json
import os

import numpy as np
import pytest

from gym_kilobots_b200 import _abi as abi
from gym_kilobots_b200 import scenarios as SC
from gym_kilobots_b200 import scene as S

pytestmark = pytest.mark.gpu

E, T = 1024, 1000


def test_thousand_step_agreement_and_reward_statistics(oracle, native):
    from scipy import stats
    sc = SC.c2_quad_assembly(E, degenerate=False)
    task = S.TaskSpec(abi.KB_TASK_OBJECT_TO_TARGET, object=3, w_position=1.0, w_orientation=0.05, step_penalty=0.001,
                      position_tolerance=0.0)
    targets = sc.body_pose[:, 3].copy()
    targets[:, 0] -= 0.3
    targets[:, 1] += 0.2
    nb = native.NativeBatch(sc.scenes, E, sc.env_scene, sc.max_contacts)
    same = oracle.OracleBatch(sc.scenes, E, sc.env_scene, sc.max_contacts, threads=os.cpu_count() or 8)
    libm = oracle.OracleBatch(sc.scenes, E, sc.env_scene, sc.max_contacts, threads=os.cpu_count() or 8, libm_trig=True)
    for b in (nb, same, libm):
        b.set_task(task, targets)
        b.reset(sc.body_pose, sc.light_state)
    acts = SC.random_actions(sc, E, 64, seed=5)
    horizon = np.full(E, T, np.int64)
    ret_n = np.zeros(E)
    ret_l = np.zeros(E)
    flags = np.zeros(E, np.int64)
    for t in range(T):
        a = acts[t % 64] * (1.0 if (t // 64) % 2 == 0 else -1.0)
        on, os_, ol = nb.step(a), same.step(a), libm.step(a)
        for k in ("kilobots", "objects", "light", "reward", "done"):
            assert np.array_equal(on[k], os_[k]), "step %d: %s differs from the oracle (same trigonometry)" % (t, k)
        flags |= on["status"]
        ret_n += on["reward"]
        ret_l += ol["reward"]
        d = np.maximum(np.abs(on["kilobots"] - ol["kilobots"]).max(axis=(1, 2)),
                       np.abs(on["objects"] - ol["objects"]).max(axis=(1, 2)))
        first = (d > 1e-4) & (horizon == T)
        horizon[first] = t
    assert np.array_equal(nb.bodies(), same.bodies()) and np.array_equal(nb.impulses(), same.impulses())
    assert np.array_equal(nb.episode_stats(), same.episode_stats())
    assert not flags.any(), "status bits raised during the soak"
    se = np.sqrt(ret_n.var() / E + ret_l.var() / E)
    ks = stats.ks_2samp(ret_n, ret_l)
    report = {
        "envs": E, "env_steps": T, "world_steps": 10 * T, "scene": "C2' (QuadAssembly, non-degenerate spawn)",
        "same_trig": {"bit_identical_env_steps": T, "note": "observations, rewards, done, final body state, impulses, "
                      "episode statistics identical to the oracle at every one of the 1000 env-steps"},
        "libm_trig": {
            "pose_agreement_horizon_env_steps": {"tolerance": "1e-4 m / rad", "min": int(horizon.min()),
                                                 "p10": float(np.percentile(horizon, 10)),
                                                 "median": float(np.median(horizon)),
                                                 "p90": float(np.percentile(horizon, 90)),
                                                 "never_diverged_frac": float((horizon == T).mean())},
            "episode_return": {"mean_cuda": float(ret_n.mean()), "mean_libm_oracle": float(ret_l.mean()),
                               "std_cuda": float(ret_n.std()), "std_libm_oracle": float(ret_l.std()),
                               "difference_of_means_in_standard_errors": float(abs(ret_n.mean() - ret_l.mean()) / se),
                               "ks_statistic": float(ks.statistic), "ks_pvalue": float(ks.pvalue)}},
        "status_flags": int((flags != 0).sum()),
    }
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "long_horizon.json"), "w") as f:
            json.dump(report, f, indent=1)
    print(json.dumps(report))
    # the stated short horizon of the pose bar (5 env-steps = 50 world steps) holds for the bulk of the envs even
    # against the other trigonometry, and the reward statistics agree over the long one
    assert np.percentile(horizon, 10) >= 5
    assert abs(ret_n.mean() - ret_l.mean()) <= 4 * se and ks.pvalue > 1e-3
