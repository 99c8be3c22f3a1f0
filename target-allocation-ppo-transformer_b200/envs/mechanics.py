"""Batched counterparts of the reference's score API (envs/mechanics.py).

The scalar functions of the reference (calc_angle_score, calc_damage_prob, calc_advantage, ...) are
evaluated on the device for every (env, UAV, target) pair at once by score_matrix_kernel
(csrc/uavenv_kernels.cuh); inside the fused step only the pair under the decision pointer is scored.
"""
import torch


def calc_advantage_matrix(env, dtype=torch.float32):
    """(p_final, p_damage) [B,N,M] in list order == mechanics.calc_advantage (mechanics.py:167-181)
    for every pair, i.e. main.py:38-45's double loop."""
    return env.score_matrix(dtype=dtype)


def calc_penetration_prob(env):
    """p_pen [B,N] (mechanics.py:118-163; target-independent): p_final / p_damage where defined."""
    pf, pd = env.score_matrix(dtype=torch.float64)
    ratio = torch.where(pd[..., 0] > 0, pf[..., 0] / pd[..., 0].clamp_min(1e-300), torch.zeros_like(pf[..., 0]))
    return ratio
