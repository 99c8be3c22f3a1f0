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


@pytest.mark.parametrize("M,K,ldx", [(1000, 128, 128), (128 * 9 + 5, 256, 128), (70001, 128, 640), (1, 128, 128)])
def test_dense_kernel_with_residual_layernorm_epilogue(M, K, ldx):
    """out = LayerNorm(x + a W^T + b) * gamma + beta fused into the GEMM epilogue (post-LN encoder layer,
    networks/transformer_net.py:34-43), with the normalised rows and 1/sigma the backward keeps."""
    import uavenv_b200  # noqa: F401
    from target_allocation_ppo_transformer_b200 import _capi
    L = _capi.load_policy()
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.7).to(torch.bfloat16)
    w = (torch.randn(128, K, device="cuda", generator=g) * 0.15).to(torch.bfloat16)
    bias = torch.randn(128, device="cuda", generator=g) * 0.3
    x_full = torch.randn(M, ldx, device="cuda", generator=g).to(torch.bfloat16)       # the residual is a strided view
    gamma = 1.0 + 0.2 * torch.randn(128, device="cuda", generator=g)
    beta = 0.3 * torch.randn(128, device="cuda", generator=g)
    out = torch.full((M, 128), 7.0, device="cuda", dtype=torch.bfloat16)
    xhat = torch.full((M, 128), 7.0, device="cuda", dtype=torch.bfloat16)
    rstd = torch.full((M,), 7.0, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = L.uavpolicy_selftest_dense_ln(p(a), K, p(w), p(bias), p(x_full), ldx, p(gamma), p(beta), p(out), p(xhat), p(rstd), M, K, None)
    assert rc == 0
    torch.cuda.synchronize()
    v = a.float() @ w.float().t() + bias + x_full[:, :128].float()
    mean = v.mean(-1, keepdim=True)
    var = ((v - mean) ** 2).mean(-1, keepdim=True)
    r = torch.rsqrt(var + 1e-5)
    xh = (v - mean) * r
    ref = xh * gamma + beta
    assert torch.allclose(rstd, r.squeeze(-1), rtol=2e-4, atol=1e-6)
    assert bool(((xhat.float() - xh).abs() <= 1e-2 * xh.abs() + 1e-2).all())
    assert bool(((out.float() - ref).abs() <= 1e-2 * ref.abs() + 2e-2).all())
