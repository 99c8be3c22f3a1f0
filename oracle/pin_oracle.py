#!/usr/bin/env python
"""Pin the C oracle against the LIVE reference (container only: needs /root/reference).

TEST INFRASTRUCTURE ONLY.  Complements tests/test_oracle_golden.py (which replays committed
fixtures): here fresh seeds and cfg overrides are drawn, the unmodified reference UAVEnv is
stepped next to the oracle on the scene it generated, and every step is compared
(integers exact, fp64 <= 1e-12 rel).

    python oracle/pin_oracle.py [--cases 12]
"""
import argparse
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_shim  # noqa: E402
import oracle as orc  # noqa: E402


def one_case(seed, overrides, p_assign):
    UAVEnv, _, cfg = ref_shim.load()
    saved = {k: getattr(cfg, k) for k in overrides}
    try:
        for k, v in overrides.items():
            setattr(cfg, k, v)
        random.seed(seed); np.random.seed(seed)
        env = UAVEnv()
        env.reset(full_reset=True)
        oe = orc.OracleEnv(orc.make_cfg(**overrides))
        oe.load_scene(ref_shim.export_scene(env))
        rng = np.random.RandomState(seed + 1)
        steps = 0
        for ep in range(2):
            obs = env.reset(full_reset=False)
            oobs = oe.reset()
            np.testing.assert_allclose(oobs, obs, rtol=3e-7, atol=1e-9)
            done = False
            while not done:
                a = int(rng.rand() < p_assign)
                obs, r, done, info = env.step(a)
                oobs, orr, odone, oinfo = oe.step(a)
                assert done == odone and env.uav_idx == oe.uav_idx and env.target_idx == oe.target_idx
                assert [u.assigned_target_id for u in env.uavs] == list(oe.assigned())
                np.testing.assert_allclose(orr, r, rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(oinfo["J_val"], info["J_val"], rtol=1e-12, atol=1e-12)
                assert oinfo["num_assigned"] == info["num_assigned"] and oinfo["is_valid_action"] == info["is_valid_action"]
                np.testing.assert_allclose(oobs, obs, rtol=3e-7, atol=1e-9)
                steps += 1
        return steps
    finally:
        for k, v in saved.items():
            setattr(cfg, k, v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=12)
    args = ap.parse_args()
    rs = np.random.RandomState(2024)
    total = 0
    for c in range(args.cases):
        ov = {"NUM_UAVS": int(rs.randint(1, 40)), "NUM_TARGETS": int(rs.randint(1, 16)),
              "NUM_NFZ": int(rs.randint(0, 3)), "NUM_INTERCEPTORS": int(rs.randint(0, 3)),
              "COST_WEIGHT_OMEGA": float(rs.choice([0.0, 0.3, 1.0])), "PARAM_K": float(rs.choice([1.2, 5.0]))}
        n = one_case(1000 + c, ov, float(rs.choice([0.3, 0.6, 0.9])))
        total += n
        print("case %2d %s: %d steps match" % (c, ov, n))
    print("oracle pinned against the live reference on %d steps" % total)


if __name__ == "__main__":
    main()
