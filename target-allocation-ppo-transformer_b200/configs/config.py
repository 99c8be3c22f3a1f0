"""Configuration mirror of the reference's `cfg` singleton (configs/config.py:5-87).

Same attribute names, same live values, read at call time - so code written against the
reference (`cfg.NUM_UAVS = 64` before building an env, `cfg.COST_WEIGHT_OMEGA = 0.5`, ...) keeps
working.  `to_c()` lowers the constants the rollout path needs into the C-ABI struct.
The reference's unused weather tables (config.py:14-30) are not carried.
"""
import copy

_DEFAULTS = {
    # Eq.(1)-(6) score parameters                                   configs/config.py:7-12
    "PARAM_ZETA_D": 150.0, "PARAM_K": 1.2,
    "PARAM_C1": 0.75, "PARAM_C2": 0.25, "PARAM_C3": 0.75, "PARAM_C4": 0.25,
    # map and scene generation                                      configs/config.py:33-50
    "MAP_WIDTH": 180.0, "MAP_HEIGHT": 160.0,
    "UAV_GEN_X_RANGE": (60, 90), "TARGET_GEN_X_RANGE": (160, 180),
    "NUM_UAVS": 30, "NUM_TARGETS": 10, "NUM_NFZ": 1, "NUM_INTERCEPTORS": 1, "INTERCEPT_RAD": 2.0,
    # objective                                                     configs/config.py:53-58
    "COST_WEIGHT_OMEGA": 0.0, "WEATHER_SPEED_FACTOR": 1.0, "WEATHER_LOAD_FACTOR": 1.0,
    # observation / action layout                                   configs/config.py:61-63
    "STATE_DIM": 14, "SEQ_LEN": 5, "ACTION_DIM": 2,
    # policy network                                                configs/config.py:66-68
    "EMBED_DIM": 128, "NUM_HEADS": 8, "NUM_LAYERS": 2,
    # PPO                                                           configs/config.py:71-85
    "LR_ACTOR": 2e-4, "LR_CRITIC": 1e-3, "GAMMA": 0.998, "GAE_LAMBDA": 0.95,
    "K_EPOCHS": 5, "EPS_CLIP": 0.2, "BATCH_SIZE": 64, "GRAD_NORM_CLIP": 1.0,
    "MAX_EPISODES": 2000, "RESET_EPISODES": 200, "SEED": 42,
}

# configs/config0.py ("paper-faithful hard mode", dead in the reference): the fields that differ
HARD_MODE = {"PARAM_K": 5.0, "UAV_GEN_X_RANGE": (0, 30), "NUM_NFZ": 2, "NUM_INTERCEPTORS": 2,
             "INTERCEPT_RAD": 3.0, "WEATHER_SPEED_FACTOR": 0.85, "WEATHER_LOAD_FACTOR": 0.90}


class Config:
    def __init__(self, **overrides):
        for name, value in _DEFAULTS.items():
            setattr(self, name, value)
        self.update(**overrides)

    def update(self, **overrides):
        for name, value in overrides.items():
            if name not in _DEFAULTS:
                raise AttributeError("unknown config field %r" % name)
            setattr(self, name, value)
        return self

    def copy(self, **overrides):
        return copy.copy(self).update(**overrides)

    def as_dict(self):
        return {name: getattr(self, name) for name in _DEFAULTS}

    def to_c(self, auto_reset=True, tie_band=1e-12):
        """uavenv_cfg_t (include/uavenv_b200.h) holding the constants of the rollout path."""
        from .._capi import UavenvCfg
        if self.STATE_DIM != 14 or self.SEQ_LEN != 5:
            raise ValueError("the kernels are specialised for STATE_DIM=14, SEQ_LEN=5 (configs/config.py:61-62)")
        c = UavenvCfg()
        c.num_uavs, c.num_targets = int(self.NUM_UAVS), int(self.NUM_TARGETS)
        c.num_nfz, c.num_interceptors = int(self.NUM_NFZ), int(self.NUM_INTERCEPTORS)
        c.reset_episodes, c.auto_reset = int(self.RESET_EPISODES), int(bool(auto_reset))
        c.param_zeta_d, c.param_k = float(self.PARAM_ZETA_D), float(self.PARAM_K)
        c.param_c1, c.param_c2 = float(self.PARAM_C1), float(self.PARAM_C2)
        c.param_c3, c.param_c4 = float(self.PARAM_C3), float(self.PARAM_C4)
        c.cost_weight_omega = float(self.COST_WEIGHT_OMEGA)
        c.weather_speed_factor, c.weather_load_factor = float(self.WEATHER_SPEED_FACTOR), float(self.WEATHER_LOAD_FACTOR)
        c.map_width, c.map_height = float(self.MAP_WIDTH), float(self.MAP_HEIGHT)
        c.uav_gen_x_lo, c.uav_gen_x_hi = float(self.UAV_GEN_X_RANGE[0]), float(self.UAV_GEN_X_RANGE[1])
        c.target_gen_x_lo, c.target_gen_x_hi = float(self.TARGET_GEN_X_RANGE[0]), float(self.TARGET_GEN_X_RANGE[1])
        c.intercept_rad = float(self.INTERCEPT_RAD)
        c.tie_band = float(tie_band)
        return c


cfg = Config()
