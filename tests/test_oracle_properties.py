"""Property tests of the CPU oracle (hypothesis): invariants of the reference's episode semantics that the GPU
parity tests rely on at full size.  CPU only, small random configurations."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as orc


@st.composite
def episodes(draw):
    n = draw(st.integers(1, 9))
    m = draw(st.integers(1, 6))
    omega = draw(st.sampled_from([0.0, 0.0, 0.4, 1.5]))
    k1, k2 = draw(st.integers(0, 2)), draw(st.integers(0, 2))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    p = draw(st.sampled_from([0.2, 0.5, 0.9]))
    return n, m, omega, k1, k2, seed, p


@settings(max_examples=60, deadline=None)
@given(episodes())
def test_episode_invariants(ep):
    n, m, omega, k1, k2, seed, p = ep
    cfg = orc.make_cfg(NUM_UAVS=n, NUM_TARGETS=m, NUM_NFZ=k1, NUM_INTERCEPTORS=k2, COST_WEIGHT_OMEGA=omega)
    env = orc.OracleEnv(cfg)
    env.generate_scene(seed, 3, 0)
    obs = env.reset()
    assert obs.shape == (5, 14) and not obs[:4].any() and obs[4, 13] == 1.0
    rng = np.random.RandomState(seed % 1000)
    total, steps, done = 0.0, 0, False
    while not done:
        a = int(rng.rand() < p)
        before = (env.uav_idx, env.target_idx, env.assigned().copy(), env.paper_reward())
        obs, r, done, info = env.step(a)
        steps += 1
        total += r
        if a != 1:
            assert info["is_valid_action"] is None and (r == 0.0 or done)
        if a == 1 and omega == 0.0:
            assert env.uav_idx == before[0] + 1 or done            # every Assign is accepted (uav_env.py:317)
        if a == 1 and not done and env.uav_idx == before[0] and env.target_idx != 0:
            # rejected (Eq.21 rollback, uav_env.py:326-342): nothing but the target pointer moved
            assert np.array_equal(env.assigned(), before[2]) and abs(env.paper_reward() - before[3]) == 0.0
        assert 0 <= info["num_assigned"] <= m
    assert n <= steps <= n * m                                      # every UAV takes 1..M decisions
    r_final = env.paper_reward()
    assert abs(total - 2.0 * r_final) <= 1e-9 * max(1.0, abs(r_final))   # telescoping return (+ goal reward :361-363)
    assert obs.shape == (14,) and not obs.any()


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(0, 1000), st.integers(0, 50))
def test_scene_generator_is_a_pure_function_of_its_key(seed, env_id, scene):
    cfg = orc.make_cfg(NUM_UAVS=8, NUM_TARGETS=5, NUM_NFZ=1, NUM_INTERCEPTORS=1)
    a, b = orc.OracleEnv(cfg), orc.OracleEnv(cfg)
    a.generate_scene(seed, env_id, scene)
    b.generate_scene(seed, env_id, scene)
    sa, sb = a.export_scene(), b.export_scene()
    for k in sa:
        assert np.array_equal(sa[k], sb[k])
    assert sorted(sa["tgt_id"].tolist()) == list(range(5))
    assert (sa["uav_type"] == 2).sum() == 2 and sorted(sa["tgt_value"].tolist())[:2] == [4.0, 4.0]
    assert sa["tgt_value"].max() == 16.0 and (sa["tgt_value"] == 16.0).sum() == 1
