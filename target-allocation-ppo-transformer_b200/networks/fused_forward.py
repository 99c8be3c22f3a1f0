"""Rollout forward of the policy on the 5th-gen tensor cores (csrc/policy_forward.cu + policy_gemm.cu) behind the
`get_action` contract of networks/transformer_net.py:96-122.

    fused = FusedPolicyForward(max_batch, device); fused.sync(policy_old)
    action, logp, value, entropy = fused.get_action(obs, step)

bf16 operands / fp32 accumulation: outputs agree with the fp32 network to ~1e-2 absolute (tested); gradients are
never taken through this path - PPO's update evaluates the fp32 PyTorch mirror."""
import ctypes as C

import torch

from .. import _capi


class FusedPolicyForward:
    def __init__(self, max_batch, device, seed=0, env_id_base=0, fused=True):
        self._h = None
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("the fused policy forward runs on a CUDA device (sm_100a) only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _capi.load_policy()
        h = C.c_void_p()
        rc = self._lib.uavpolicy_create(self.device.index, int(max_batch), C.byref(h))
        if rc != 0:
            raise _capi.UavenvError(rc, (self._lib.uavpolicy_last_error(None) or b"").decode())
        self._h = h
        self._lib.uavpolicy_set_fused(h, int(bool(fused)))   # False: one tcgen05 GEMM launch per layer (A/B reference)
        self.max_batch, self.seed, self.env_id_base = int(max_batch), int(seed), int(env_id_base)
        dev = self.device
        self.action = torch.zeros(max_batch, dtype=torch.int64, device=dev)
        self.logp = torch.zeros(max_batch, device=dev)
        self.value = torch.zeros(max_batch, 1, device=dev)
        self.entropy = torch.zeros(max_batch, device=dev)
        self.logits = torch.zeros(max_batch, 2, device=dev)
        self._flat = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc):
        if rc != 0:
            raise _capi.UavenvError(rc, (self._lib.uavpolicy_last_error(self._h) or b"").decode())

    def sync(self, module):
        """Copy the weights of a TransformerActorCritic (named_parameters order == the reference's state_dict order)."""
        flat = torch.cat([p.detach().reshape(-1).float() for p in module.parameters()]).contiguous()
        if flat.numel() != 419267 or flat.device != self.device:
            raise ValueError("expected the 419267 parameters of TransformerActorCritic on %s" % self.device)
        self._flat = flat
        self._chk(self._lib.uavpolicy_set_weights(self._h, C.c_void_p(flat.data_ptr()), self._stream()))

    def get_action(self, obs, step):
        """(action int64 [B], log_prob [B], value [B,1], entropy [B]); tensors are reused by the next call."""
        B = obs.shape[0]
        obs = obs.contiguous()
        self._chk(self._lib.uavpolicy_get_action(
            self._h, C.c_void_p(obs.data_ptr()), B, self.seed, int(step), self.env_id_base,
            C.c_void_p(self.action.data_ptr()), C.c_void_p(self.logp.data_ptr()), C.c_void_p(self.value.data_ptr()),
            C.c_void_p(self.entropy.data_ptr()), C.c_void_p(self.logits.data_ptr()), self._stream()))
        return self.action[:B], self.logp[:B], self.value[:B], self.entropy[:B]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.uavpolicy_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
