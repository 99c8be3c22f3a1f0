"""PPO rollout post-processing on the device (agents/ppo.py:77-94 of the reference).

compute_gae runs the chunked warp-scan GAE kernel + advantage normalisation of csrc/ppo_gae.cu on a
time-major [T,B] rollout.  With a torch.distributed process group the three normalisation
statistics {count, sum, sum of squares} are all-reduced so every rank normalises with the global
mean / unbiased std, as one big single-GPU batch would.
"""
import ctypes as C

import torch

from .. import _capi


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def compute_gae(rewards, values, dones, last_value=None, gamma=0.998, lam=0.95, normalize=True, group=None):
    """rewards, values: f32 [T,B]; dones: bool/uint8 [T,B]; last_value: f32 [B] or None (episodic, V_T = 0).
    Returns (returns [T,B], advantages [T,B]).  Defaults are cfg.GAMMA / cfg.GAE_LAMBDA."""
    if rewards.dim() == 1:
        rewards, values, dones = rewards[:, None], values[:, None], dones[:, None]
        squeeze = True
    else:
        squeeze = False
    if not rewards.is_cuda:
        raise RuntimeError("compute_gae runs on the GPU only (no CPU fallback)")
    T, B = rewards.shape
    dev = rewards.device
    rewards = rewards.contiguous().float()
    values = values.contiguous().float()
    dones_u8 = dones.contiguous().view(torch.uint8) if dones.dtype == torch.bool else dones.contiguous().to(torch.uint8)
    if last_value is not None:
        last_value = last_value.contiguous().float()
    returns = torch.empty_like(rewards)
    adv = torch.empty_like(rewards)
    stats = torch.zeros(3, dtype=torch.float64, device=dev)
    lib = _capi.load()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    distributed = group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                        and torch.distributed.get_world_size() > 1)
    local_norm = int(bool(normalize) and not distributed)
    rc = lib.ppo_gae_advantages(_ptr(rewards), _ptr(values), _ptr(dones_u8), _ptr(last_value), T, B, float(gamma),
                                float(lam), _ptr(returns), _ptr(adv), local_norm, _ptr(stats), dev.index, stream)
    if rc != 0:
        raise _capi.UavenvError(rc, "ppo_gae_advantages failed")
    if normalize and distributed:
        torch.distributed.all_reduce(stats, group=group)
        rc = lib.ppo_normalize_advantages(_ptr(adv), T * B, _ptr(stats), dev.index, stream)
        if rc != 0:
            raise _capi.UavenvError(rc, "ppo_normalize_advantages failed")
    if squeeze:
        returns, adv = returns[:, 0], adv[:, 0]
    return returns, adv


def normalize_advantages(adv, stats):
    lib = _capi.load()
    stream = C.c_void_p(torch.cuda.current_stream(adv.device).cuda_stream)
    rc = lib.ppo_normalize_advantages(_ptr(adv), adv.numel(), _ptr(stats), adv.device.index, stream)
    if rc != 0:
        raise _capi.UavenvError(rc, "ppo_normalize_advantages failed")
    return adv
