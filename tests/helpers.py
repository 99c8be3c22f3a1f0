"""Shared helpers of the parity tests."""
import os

import numpy as np

from conftest import GOLDEN

CFG_INT = {"NUM_UAVS", "NUM_TARGETS", "NUM_NFZ", "NUM_INTERCEPTORS"}


def load_fixture(case):
    return np.load(os.path.join(GOLDEN, case + ".npz"))


def config_from_fixture(fx, **extra):
    """Product-side Config carrying the cfg overrides the reference ran the fixture with."""
    import uavenv_b200 as ub
    kw = {}
    for n, v in zip(fx["cfg_names"], fx["cfg_values"]):
        n = str(n)
        if n in ("UAV_GEN_X_RANGE", "TARGET_GEN_X_RANGE"):
            continue
        kw[n] = int(v) if n in CFG_INT else float(v)
    kw["UAV_GEN_X_RANGE"] = tuple(float(x) for x in fx["cfg_uav_gen_x"])
    kw["TARGET_GEN_X_RANGE"] = tuple(float(x) for x in fx["cfg_target_gen_x"])
    kw.update(extra)
    return ub.Config(**kw)


def oracle_cfg_from_config(cfg):
    from oracle import oracle as orc
    return orc.make_cfg(**{k: v for k, v in cfg.as_dict().items() if k in (
        "NUM_UAVS", "NUM_TARGETS", "NUM_NFZ", "NUM_INTERCEPTORS", "PARAM_ZETA_D", "PARAM_K", "PARAM_C1", "PARAM_C2",
        "PARAM_C3", "PARAM_C4", "COST_WEIGHT_OMEGA", "WEATHER_SPEED_FACTOR", "WEATHER_LOAD_FACTOR", "MAP_WIDTH",
        "MAP_HEIGHT", "INTERCEPT_RAD", "UAV_GEN_X_RANGE", "TARGET_GEN_X_RANGE")})


SCENE_KEYS = ["uav_x", "uav_y", "uav_vx", "uav_vy", "uav_load", "uav_cost", "uav_type", "tgt_x", "tgt_y", "tgt_vx",
              "tgt_vy", "tgt_value", "tgt_id", "nfz_x", "nfz_y", "nfz_radius", "int_x", "int_y", "int_vx", "int_vy"]


def scene_from_fixture(fx, copies=1):
    return {k: np.tile(np.asarray(fx[k]), (copies, 1)) for k in SCENE_KEYS}


def rel_close(got, want, rtol, atol=0.0):
    np.testing.assert_allclose(np.asarray(got, np.float64), np.asarray(want, np.float64), rtol=rtol, atol=atol)
