"""The diagnostics that stand in for the reference's main.py / test_visualize.py scripts, against matrices and
trajectories recorded from the unmodified reference (tests/golden/traj_*.npz)."""
import numpy as np
import pytest
import torch

from helpers import config_from_fixture, load_fixture, scene_from_fixture

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["traj_default_s0", "traj_hard_s0"])
def test_difficulty_report_matches_the_reference_matrices(case):
    """main.py:38-78 on an injected reference scene: all-pairs means, best-UAV-per-target means, penetration ratio
    from the reference's own p_final / p_damage matrices (fp64, 1e-9)."""
    import uavenv_b200 as ub
    try:
        fx = load_fixture(case)
    except FileNotFoundError:
        pytest.skip("fixture %s not recorded" % case)
    env = ub.UAVEnvBatched(3, config=config_from_fixture(fx), auto_reset=False)
    env.load_scene(scene_from_fixture(fx, copies=3))
    rep = ub.analyze_environment_difficulty(env=env, verbose=False)
    pf, pd = np.asarray(fx["p_final"], np.float64), np.asarray(fx["p_damage"], np.float64)
    want = {"avg_p_dmg": pd.mean(), "avg_p_final": pf.mean(), "best_p_dmg": pd.max(axis=0).mean(), "best_p_final": pf.max(axis=0).mean()}
    for k, v in want.items():
        np.testing.assert_allclose(rep[k], np.full(3, v), rtol=1e-9)
    np.testing.assert_allclose(rep["pen_rate"], want["best_p_final"] / (want["best_p_dmg"] + 1e-6), rtol=1e-9)
    assert rep["summary"]["rounds"] == 3 and isinstance(rep["verdict"], str)
    env.close()


def test_difficulty_report_on_generated_scenes_runs_like_main_py():
    import uavenv_b200 as ub
    rep = ub.analyze_environment_difficulty(num_rounds=10, seed=1, verbose=False)
    assert rep["avg_p_dmg"].shape == (10,) and (rep["best_p_final"] >= rep["avg_p_final"]).all()
    assert 0.0 < rep["summary"]["penetration"] <= 1.0 + 1e-6


@pytest.mark.parametrize("case", ["traj_default_s0", "traj_omega20_s0"])
def test_recorded_decisions_match_the_reference_trajectory(case):
    """test_visualize.py:33-45 with the recorded action stream as the 'policy': the (uav id, target id) pair of every
    Assign decision, and the accept flags, equal what the reference env produced."""
    import uavenv_b200 as ub
    try:
        fx = load_fixture(case)
    except FileNotFoundError:
        pytest.skip("fixture %s not recorded" % case)
    first = np.asarray(fx["episode"]) == np.asarray(fx["episode"])[0]
    actions = np.asarray(fx["action"])[first]
    env = ub.UAVEnvBatched(2, config=config_from_fixture(fx), auto_reset=False)
    env.load_scene(scene_from_fixture(fx, copies=2))
    step = [0]

    def replay(obs):
        a = int(actions[min(step[0], len(actions) - 1)])
        step[0] += 1
        return torch.full((2,), a, dtype=torch.int64, device=obs.device)

    rec = ub.record_decisions(env, replay)
    T = rec["steps"]
    assert T == len(actions)
    tgt_id = np.asarray(fx["tgt_id"]).reshape(-1)
    k, m = np.asarray(fx["uav_idx"])[first], np.asarray(fx["target_idx"])[first]
    # the fixture stores the pointers AFTER each step; the pair decided at step t is the pointer before it
    k_before = np.concatenate([[0], k[:-1]]); m_before = np.concatenate([[0], m[:-1]])
    for b in range(2):
        assert np.array_equal(rec["uav_id"][:, b], k_before)
        assert np.array_equal(rec["target_id"][:, b], tgt_id[m_before])
        assert np.array_equal(rec["action"][:, b], actions)
        valid = np.asarray(fx["is_valid"])[first]
        assert np.array_equal(rec["accepted"][:, b], np.where(actions == 1, (valid == 1).astype(np.int64), 0))
        assert rec["assignments"][b] == [(int(u), int(tgt_id[t])) for u, t, a in zip(k_before, m_before, actions) if a == 1]
    assert rec["scene"]["uav_x"].shape == (2, env.N)
    env.close()
