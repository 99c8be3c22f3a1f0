"""The hand-written dense-layer kernel (csrc/policy_dense.cu: persistent TMA-fed tcgen05 GEMM with fused bias / ReLU /
ReLU-backward epilogues) against a plain PyTorch fp32 reference of the same op on the same bf16 inputs: the shapes the
network uses (networks/transformer_net.py:24-91), ragged M, strided A."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 128, 128), (130, 384, 128), (128 * 150 + 7, 128, 384), (70000, 256, 128), (5000, 128, 256),
          (777, 64, 128), (777, 128, 64), (1, 128, 128), (128, 384, 128), (300001, 128, 128)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_dense_kernel_matches_fp32_reference(M, N, K, act):
    import uavenv_b200  # noqa: F401
    from target_allocation_ppo_transformer_b200 import _capi
    L = _capi.load_policy()
    g = torch.Generator(device="cuda").manual_seed(M + 7 * N + 13 * K + act)
    lda = K + 64                                                     # A is a strided view, as the last-token Q projection is
    a_full = (torch.randn(M, lda, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
    a = a_full[:, :K]
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g) if act != 2 else None
    aux = None
    if act == 2:
        aux = torch.relu(torch.randn(M, N, device="cuda", generator=g)).to(torch.bfloat16)
    out = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    rc = L.uavpolicy_selftest_dense(C.c_void_p(a_full.data_ptr()), lda, C.c_void_p(w.data_ptr()),
                                    C.c_void_p(bias.data_ptr()) if bias is not None else None,
                                    C.c_void_p(aux.data_ptr()) if aux is not None else None, N, C.c_void_p(out.data_ptr()), M, N, K,
                                    act, None)
    assert rc == 0
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if act != 2:
        ref = ref + bias
    if act == 1:
        ref = torch.relu(ref)
    if act == 2:
        ref = torch.where(aux.float() > 0, ref, torch.zeros_like(ref))
    # bf16 output rounding (2^-8 relative) on top of an exact-product fp32 accumulation
    err = (out.float() - ref).abs()
    tol = 1e-2 * ref.abs() + 2e-2
    assert bool((err <= tol).all()), (float(err.max()), int((err > tol).sum()))
    if act == 2:
        assert bool((out[aux.float() <= 0] == 0).all())
