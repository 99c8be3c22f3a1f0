"""Entity records of one scene, as host-side views of the device state.

Field names follow the reference's dataclasses (envs/entities.py:13-61) so analysis code written
against `env.uavs[i].pos`, `env.targets[j].locked_by_uavs`, ... reads the same.  These objects are
snapshots produced by UAVEnvBatched.scene(b); the live state is the structure-of-arrays in HBM.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class UAV:                      # envs/entities.py:13-36
    id: int
    pos: np.ndarray
    velocity: np.ndarray
    max_speed: float
    load: float
    uav_type: int
    cost: float
    assigned_target_id: int = -1
    available: bool = True


@dataclass
class Target:                   # envs/entities.py:39-49 (+ velocity, uav_env.py:141)
    id: int
    pos: np.ndarray
    value: float
    velocity: np.ndarray
    locked_by_uavs: List[int] = field(default_factory=list)


@dataclass
class NoFlyZone:                # envs/entities.py:52-55
    id: int
    pos: np.ndarray
    radius: float


@dataclass
class Interceptor:              # envs/entities.py:58-61 (+ velocity, uav_env.py:168)
    id: int
    pos: np.ndarray
    radius: float
    velocity: np.ndarray


def build_entities(scene, assigned, intercept_rad):
    """Entity lists of ONE env from its SoA scene (list order) and assigned_target_id[N]."""
    n, m = len(scene["uav_x"]), len(scene["tgt_x"])
    uavs = []
    for i in range(n):
        vel = np.array([scene["uav_vx"][i], scene["uav_vy"][i]])
        a = int(assigned[i])
        uavs.append(UAV(id=i, pos=np.array([scene["uav_x"][i], scene["uav_y"][i]]), velocity=vel,
                        max_speed=float(np.linalg.norm(vel)), load=float(scene["uav_load"][i]),
                        uav_type=int(scene["uav_type"][i]), cost=float(scene["uav_cost"][i]),
                        assigned_target_id=a, available=a < 0))
    targets = []
    for j in range(m):
        tid = int(scene["tgt_id"][j])
        # UAVs decide in ascending order, so ascending ids == lock order (uav_env.py:310,324)
        locked = [i for i in range(n) if int(assigned[i]) == tid]
        targets.append(Target(id=tid, pos=np.array([scene["tgt_x"][j], scene["tgt_y"][j]]),
                              value=float(scene["tgt_value"][j]),
                              velocity=np.array([scene["tgt_vx"][j], scene["tgt_vy"][j]]), locked_by_uavs=locked))
    nfz = [NoFlyZone(id=i, pos=np.array([scene["nfz_x"][i], scene["nfz_y"][i]]), radius=float(scene["nfz_radius"][i]))
           for i in range(len(scene["nfz_x"]))]
    inter = [Interceptor(id=i, pos=np.array([scene["int_x"][i], scene["int_y"][i]]), radius=float(intercept_rad),
                         velocity=np.array([scene["int_vx"][i], scene["int_vy"][i]]))
             for i in range(len(scene["int_x"]))]
    return uavs, targets, nfz, inter
