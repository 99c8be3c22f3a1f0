"""Import the UNMODIFIED reference (read-only at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py (to produce the
committed fixtures under tests/golden/) and by oracle/pin_oracle.py (to pin the
C restatement against the live reference).  Nothing on the product path, and
nothing that runs on the GPU box, imports this module: /root/reference does not
exist there.

Shims:
  * `gym` is not installed: a stub exposing gym.Env, gym.spaces.Discrete and
    gym.spaces.Box is registered before envs/uav_env.py is imported
    (envs/uav_env.py:2,6,13,18,21 are the only uses).
  * site-packages ships an unrelated regular package `agents`; a namespace
    module pointing at /root/reference/agents is pre-registered.
  * the reference tree is read-only, so byte-code writing is disabled.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("UAVENV_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "envs", "uav_env.py"))


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Env:  # noqa: D401 - minimal stand-in for gym.Env
        pass

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    gym.Env = Env
    spaces.Discrete = Discrete
    spaces.Box = Box
    gym.spaces = spaces
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


def load():
    """Return (UAVEnv class, mechanics module, cfg singleton) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    _install_gym_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own package mirrors the reference's sub-package names (envs/, configs/...)
    # but lives under a different top-level name, so there is no clash.
    for name in ("envs", "configs", "networks"):
        mod = sys.modules.get(name)
        if mod is not None and REFERENCE_ROOT not in str(getattr(mod, "__path__", "")):
            del sys.modules[name]
    import envs.uav_env as ref_env  # noqa: E402
    import envs.mechanics as ref_mech  # noqa: E402
    from configs.config import cfg as ref_cfg  # noqa: E402
    return ref_env.UAVEnv, ref_mech, ref_cfg


def load_agents():
    """Return (PPOAgent class, TransformerActorCritic class) of the reference."""
    load()
    if "agents" not in sys.modules or REFERENCE_ROOT not in str(
            getattr(sys.modules["agents"], "__path__", "")):
        ns = types.ModuleType("agents")
        ns.__path__ = [os.path.join(REFERENCE_ROOT, "agents")]
        sys.modules["agents"] = ns
    import agents.ppo as ref_ppo  # noqa: E402
    import networks.transformer_net as ref_net  # noqa: E402
    return ref_ppo.PPOAgent, ref_net.TransformerActorCritic


def export_scene(env):
    """SoA (fp64 / int32) dump of a reference UAVEnv's scene, in LIST order."""
    import numpy as np
    u, t = env.uavs, env.targets
    f64 = np.float64
    scene = {
        "uav_x": np.array([a.pos[0] for a in u], f64),
        "uav_y": np.array([a.pos[1] for a in u], f64),
        "uav_vx": np.array([a.velocity[0] for a in u], f64),
        "uav_vy": np.array([a.velocity[1] for a in u], f64),
        "uav_load": np.array([a.load for a in u], f64),
        "uav_cost": np.array([a.cost for a in u], f64),
        "uav_type": np.array([a.uav_type for a in u], np.int32),
        "tgt_x": np.array([a.pos[0] for a in t], f64),
        "tgt_y": np.array([a.pos[1] for a in t], f64),
        "tgt_vx": np.array([a.velocity[0] for a in t], f64),
        "tgt_vy": np.array([a.velocity[1] for a in t], f64),
        "tgt_value": np.array([a.value for a in t], f64),
        "tgt_id": np.array([a.id for a in t], np.int32),
        "nfz_x": np.array([a.pos[0] for a in env.nfz_list], f64),
        "nfz_y": np.array([a.pos[1] for a in env.nfz_list], f64),
        "nfz_radius": np.array([a.radius for a in env.nfz_list], f64),
        "int_x": np.array([a.pos[0] for a in env.interceptors], f64),
        "int_y": np.array([a.pos[1] for a in env.interceptors], f64),
        "int_vx": np.array([a.velocity[0] for a in env.interceptors], f64),
        "int_vy": np.array([a.velocity[1] for a in env.interceptors], f64),
    }
    return scene
