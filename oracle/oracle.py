"""ctypes front-end of the CPU oracle (oracle/uavenv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py - never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libuavenv_oracle.so")

STATE_DIM = 14
SEQ_LEN = 5

SCENE_F64 = ["uav_x", "uav_y", "uav_vx", "uav_vy", "uav_load", "uav_cost"]
SCENE_ORDER = ["uav_x", "uav_y", "uav_vx", "uav_vy", "uav_load", "uav_cost", "uav_type", "tgt_x", "tgt_y",
               "tgt_vx", "tgt_vy", "tgt_value", "tgt_id", "nfz_x", "nfz_y", "nfz_radius", "int_x", "int_y",
               "int_vx", "int_vy"]
SCENE_I32 = {"uav_type", "tgt_id"}


class OrcCfg(C.Structure):
    _fields_ = [("num_uavs", C.c_int32), ("num_targets", C.c_int32), ("num_nfz", C.c_int32),
                ("num_interceptors", C.c_int32),
                ("zeta_d", C.c_double), ("k", C.c_double), ("c1", C.c_double), ("c2", C.c_double),
                ("c3", C.c_double), ("c4", C.c_double), ("omega", C.c_double),
                ("weather_speed", C.c_double), ("weather_load", C.c_double),
                ("map_w", C.c_double), ("map_h", C.c_double),
                ("uav_x_lo", C.c_double), ("uav_x_hi", C.c_double), ("tgt_x_lo", C.c_double),
                ("tgt_x_hi", C.c_double), ("intercept_rad", C.c_double)]


class OrcInfo(C.Structure):
    _fields_ = [("J_val", C.c_double), ("num_assigned", C.c_int32), ("is_valid_action", C.c_int32),
                ("avg_p_dmg", C.c_double), ("avg_p_final", C.c_double)]


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("uavenv_oracle.c", "uavenv_oracle.h")]
    if (not force and os.path.isfile(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.run(["make", "-C", HERE, "-s"], check=True, env={k: v for k, v in os.environ.items() if k != "CC"})
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_float)
        L.orc_default_cfg.argtypes = [C.POINTER(OrcCfg)]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcCfg)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_load_scene.argtypes = [C.c_void_p] + [ip if n in SCENE_I32 else dp for n in SCENE_ORDER]
        L.orc_export_scene.argtypes = [C.c_void_p] + [ip if n in SCENE_I32 else dp for n in SCENE_ORDER]
        L.orc_generate_scene.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_reset.argtypes = [C.c_void_p, fp]
        L.orc_step.restype = C.c_int
        L.orc_step.argtypes = [C.c_void_p, C.c_int64, fp, ip, dp, ip, C.POINTER(OrcInfo)]
        L.orc_score_matrix.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_angle_score.restype = C.c_double
        L.orc_angle_score.argtypes = [C.c_double] * 6
        L.orc_speed_score.restype = C.c_double
        L.orc_speed_score.argtypes = [C.POINTER(OrcCfg), C.c_double, C.c_double]
        L.orc_dist_score.restype = C.c_double
        L.orc_dist_score.argtypes = [C.POINTER(OrcCfg), C.c_double, C.c_int]
        L.orc_damage_prob.restype = C.c_double
        L.orc_damage_prob.argtypes = [C.POINTER(OrcCfg)] + [C.c_double] * 9
        L.orc_state_vector_raw.argtypes = [C.c_double] * 10 + [C.c_int, fp]
        L.orc_uav_idx.restype = C.c_int32
        L.orc_uav_idx.argtypes = [C.c_void_p]
        L.orc_target_idx.restype = C.c_int32
        L.orc_target_idx.argtypes = [C.c_void_p]
        L.orc_get_assigned.argtypes = [C.c_void_p, ip]
        L.orc_get_covered.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.orc_paper_reward.restype = C.c_double
        L.orc_paper_reward.argtypes = [C.c_void_p]
        L.orc_calc_J.restype = C.c_double
        L.orc_calc_J.argtypes = [C.c_void_p]
        L.orc_random_action.restype = C.c_int64
        L.orc_random_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.orc_rollout_random.restype = C.c_int64
        L.orc_rollout_random.argtypes = [C.POINTER(OrcCfg), C.c_int32, C.c_int64, C.c_uint64, C.c_uint64,
                                         C.c_int32, C.c_int32, dp]
        L.orc_batch_create.restype = C.c_void_p
        L.orc_batch_create.argtypes = [C.POINTER(OrcCfg), C.c_int32, C.c_uint64, C.c_int32, C.c_int32]
        L.orc_batch_destroy.argtypes = [C.c_void_p]
        L.orc_batch_step_random.restype = C.c_double
        L.orc_batch_step_random.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32]
        L.orc_batch_step_actions.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int32, dp, C.POINTER(C.c_uint8), ip, ip, ip, ip]
        L.orc_batch_get_assigned.argtypes = [C.c_void_p, ip]
        L.orc_max_threads.restype = C.c_int32
        _lib = L
    return _lib


def make_cfg(**kw):
    """orc_cfg with configs/config.py defaults, overridden by reference-style names
    (NUM_UAVS, PARAM_K, COST_WEIGHT_OMEGA, UAV_GEN_X_RANGE, ...) or field names."""
    c = OrcCfg()
    lib().orc_default_cfg(C.byref(c))
    alias = {"NUM_UAVS": "num_uavs", "NUM_TARGETS": "num_targets", "NUM_NFZ": "num_nfz",
             "NUM_INTERCEPTORS": "num_interceptors", "PARAM_ZETA_D": "zeta_d", "PARAM_K": "k",
             "PARAM_C1": "c1", "PARAM_C2": "c2", "PARAM_C3": "c3", "PARAM_C4": "c4",
             "COST_WEIGHT_OMEGA": "omega", "WEATHER_SPEED_FACTOR": "weather_speed",
             "WEATHER_LOAD_FACTOR": "weather_load", "MAP_WIDTH": "map_w", "MAP_HEIGHT": "map_h",
             "INTERCEPT_RAD": "intercept_rad"}
    for k, v in kw.items():
        if k == "UAV_GEN_X_RANGE":
            c.uav_x_lo, c.uav_x_hi = float(v[0]), float(v[1])
        elif k == "TARGET_GEN_X_RANGE":
            c.tgt_x_lo, c.tgt_x_hi = float(v[0]), float(v[1])
        else:
            f = alias.get(k, k)
            cur = getattr(c, f)
            setattr(c, f, int(v) if isinstance(cur, int) else float(v))
    return c


def cfg_from_fixture(fx):
    """orc_cfg from a tests/golden/traj_*.npz fixture."""
    kw = {str(n): float(v) for n, v in zip(fx["cfg_names"], fx["cfg_values"])
          if str(n) not in ("UAV_GEN_X_RANGE", "TARGET_GEN_X_RANGE")}
    kw["UAV_GEN_X_RANGE"] = tuple(fx["cfg_uav_gen_x"])
    kw["TARGET_GEN_X_RANGE"] = tuple(fx["cfg_target_gen_x"])
    return make_cfg(**kw)


def _ptr(a):
    if a.dtype == np.float64:
        return a.ctypes.data_as(C.POINTER(C.c_double))
    if a.dtype == np.int32:
        return a.ctypes.data_as(C.POINTER(C.c_int32))
    if a.dtype == np.float32:
        return a.ctypes.data_as(C.POINTER(C.c_float))
    if a.dtype == np.uint8:
        return a.ctypes.data_as(C.POINTER(C.c_uint8))
    raise TypeError(a.dtype)


class OracleEnv:
    """Single-env CPU oracle with the reference's reset/step contract (envs/uav_env.py:42,295)."""

    def __init__(self, cfg=None):
        self.cfg = cfg if cfg is not None else make_cfg()
        self.N, self.M = self.cfg.num_uavs, self.cfg.num_targets
        self.K1, self.K2 = self.cfg.num_nfz, self.cfg.num_interceptors
        self._h = C.c_void_p(lib().orc_create(C.byref(self.cfg)))
        self._obs = np.zeros((SEQ_LEN, STATE_DIM), np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def _sizes(self, name):
        return {"uav": self.N, "tgt": self.M, "nfz": self.K1, "int": self.K2}[name.split("_")[0]]

    def load_scene(self, scene):
        arrs = []
        for n in SCENE_ORDER:
            dt = np.int32 if n in SCENE_I32 else np.float64
            a = np.ascontiguousarray(np.asarray(scene[n], dt).reshape(-1))
            if a.size == 0:
                a = np.zeros(1, dt)
            arrs.append(a)
        lib().orc_load_scene(self._h, *[_ptr(a) for a in arrs])

    def generate_scene(self, seed, env_id, scene_idx):
        lib().orc_generate_scene(self._h, int(seed), int(env_id), int(scene_idx))

    def export_scene(self):
        out = {}
        for n in SCENE_ORDER:
            out[n] = np.zeros(max(self._sizes(n), 1), np.int32 if n in SCENE_I32 else np.float64)
        lib().orc_export_scene(self._h, *[_ptr(out[n]) for n in SCENE_ORDER])
        return {n: out[n][: self._sizes(n)] for n in SCENE_ORDER}

    def reset(self):
        """State-only part of reset (the scene is whatever was loaded/generated last)."""
        lib().orc_reset(self._h, _ptr(self._obs.reshape(-1)))
        return self._obs.copy()

    def step(self, action):
        rows, done = C.c_int32(0), C.c_int32(0)
        reward = C.c_double(0.0)
        info = OrcInfo()
        rc = lib().orc_step(self._h, int(action), _ptr(self._obs.reshape(-1)), C.byref(rows), C.byref(reward),
                            C.byref(done), C.byref(info))
        if rc != 0:
            raise IndexError("step() on a finished episode (envs/uav_env.py:296)")
        obs = self._obs.copy() if rows.value == SEQ_LEN else np.zeros(STATE_DIM, np.float32)
        valid = None if info.is_valid_action < 0 else bool(info.is_valid_action)
        return obs, reward.value, bool(done.value), {
            "J_val": info.J_val, "num_assigned": info.num_assigned, "is_valid_action": valid,
            "avg_p_dmg": info.avg_p_dmg, "avg_p_final": info.avg_p_final}

    def score_matrix(self):
        pf = np.zeros((self.N, self.M))
        pd = np.zeros((self.N, self.M))
        pp = np.zeros(max(self.N, 1))
        lib().orc_score_matrix(self._h, _ptr(pf.reshape(-1)), _ptr(pd.reshape(-1)), _ptr(pp))
        return pf, pd, pp[: self.N]

    @property
    def uav_idx(self):
        return lib().orc_uav_idx(self._h)

    @property
    def target_idx(self):
        return lib().orc_target_idx(self._h)

    def assigned(self):
        a = np.zeros(max(self.N, 1), np.int32)
        lib().orc_get_assigned(self._h, _ptr(a))
        return a[: self.N]

    def covered(self):
        a = np.zeros(max(self.M, 1), np.uint8)
        lib().orc_get_covered(self._h, _ptr(a))
        return a[: self.M]

    def paper_reward(self):
        return lib().orc_paper_reward(self._h)

    def calc_J(self):
        return lib().orc_calc_J(self._h)


def random_action(action_seed, step, env_id):
    return int(lib().orc_random_action(int(action_seed), int(step), int(env_id)))


def philox(k0, k1, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(k0, k1, c0, c1, c2, c3, out)
    return [int(x) for x in out]


def rollout_random(cfg, num_envs, steps, seed=42, action_seed=1, reset_episodes=200, threads=0):
    """CPU baseline loop; returns (transitions, reward checksum)."""
    chk = C.c_double(0.0)
    n = lib().orc_rollout_random(C.byref(cfg), num_envs, steps, seed, action_seed, reset_episodes, threads,
                                 C.byref(chk))
    return int(n), chk.value


class OracleBatch:
    """Persistent batch of independent oracle envs stepped with the Bernoulli action stream on
    OpenMP threads (the timed CPU arm of bench.py)."""

    def __init__(self, cfg, num_envs, seed=42, reset_episodes=200, threads=0):
        self.num_envs, self.threads = num_envs, threads
        self._h = C.c_void_p(lib().orc_batch_create(C.byref(cfg), num_envs, seed, reset_episodes, threads))

    def step_random(self, step, action_seed=1):
        return lib().orc_batch_step_random(self._h, action_seed, step, self.threads)

    def step_actions(self, actions):
        """One reference-algorithm step of every env (auto-restart as main_train.py:79).  Returns per-env arrays
        (reward f64, done, num_assigned, is_valid {-1,0,1}, uav_idx, target_idx) - pointers AFTER the restart."""
        E = self.num_envs
        a = np.ascontiguousarray(actions, np.int64)
        out = (np.zeros(E), np.zeros(E, np.uint8), np.zeros(E, np.int32), np.zeros(E, np.int32), np.zeros(E, np.int32),
               np.zeros(E, np.int32))
        lib().orc_batch_step_actions(self._h, a.ctypes.data_as(C.POINTER(C.c_int64)), self.threads, _ptr(out[0]), _ptr(out[1]),
                                     _ptr(out[2]), _ptr(out[3]), _ptr(out[4]), _ptr(out[5]))
        return out

    def assigned(self, num_uavs):
        a = np.zeros((self.num_envs, num_uavs), np.int32)
        lib().orc_batch_get_assigned(self._h, _ptr(a.reshape(-1)))
        return a

    def close(self):
        if self._h:
            lib().orc_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def max_threads():
    """Host threads available to this process (CPU affinity), regardless of OMP_NUM_THREADS (torchrun sets it to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)
