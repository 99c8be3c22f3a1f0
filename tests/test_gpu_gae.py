"""GAE + advantage-normalisation kernel (csrc/ppo_gae.cu) against a plain PyTorch restatement of the
reference recurrence (agents/ppo.py:77-94), which is the fp32 reference for this floating-point kernel.
Tolerance: 1e-5 relative to the rollout's advantage scale (fp32 outputs; the kernel scans in fp64)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_gae(rewards, values, dones, last_value, gamma, lam):
    """agents/ppo.py:77-94 on a [T,B] rollout, evaluated in fp64."""
    T, B = rewards.shape
    r, v, d = rewards.double(), values.double(), dones.bool()
    nxt = torch.zeros(B, dtype=torch.float64, device=r.device) if last_value is None else last_value.double()
    gae = torch.zeros(B, dtype=torch.float64, device=r.device)
    ret = torch.empty_like(r)
    for t in reversed(range(T)):
        nd = (~d[t]).double()
        delta = r[t] + gamma * nxt * nd - v[t]                   # ppo.py:86 (v_next = 0 when done)
        gae = delta + gamma * lam * gae * nd                      # ppo.py:87
        ret[t] = gae + v[t]                                       # ppo.py:88
        nxt = v[t]
    adv = ret.float() - values                                    # ppo.py:92 (f32 tensors in the reference)
    return ret.float(), adv


@pytest.mark.parametrize("T,B", [(1, 1), (3, 5), (37, 65), (128, 4096), (257, 33), (2048, 64), (300, 1)])
@pytest.mark.parametrize("bootstrap", [False, True])
def test_gae_matches_reference_recurrence(T, B, bootstrap):
    import uavenv_b200 as ub
    g = torch.Generator(device="cuda").manual_seed(T * 131 + B)
    r = torch.randn(T, B, device="cuda", generator=g) * 2.0
    v = torch.randn(T, B, device="cuda", generator=g) * 3.0
    d = torch.rand(T, B, device="cuda", generator=g) < 0.03
    lv = torch.randn(B, device="cuda", generator=g) if bootstrap else None
    gamma, lam = 0.998, 0.95                                       # configs/config.py:73-74
    ret, adv = ub.compute_gae(r, v, d, lv, gamma, lam, normalize=False)
    ret_ref, adv_ref = _reference_gae(r, v, d, lv, float(torch.tensor(gamma).float()), float(torch.tensor(lam).float()))
    scale = max(1.0, float(adv_ref.abs().max()))
    assert torch.allclose(ret, ret_ref, rtol=1e-5, atol=1e-5 * scale)
    assert torch.allclose(adv, adv_ref, rtol=1e-5, atol=1e-5 * scale)
    if T * B > 1:
        _, adv_n = ub.compute_gae(r, v, d, lv, gamma, lam, normalize=True)
        want = (adv_ref - adv_ref.mean()) / (adv_ref.std() + 1e-7)   # ppo.py:94 (unbiased std)
        assert torch.allclose(adv_n, want, rtol=1e-4, atol=2e-5)


def test_gae_episodic_buffer_like_the_reference():
    """The reference's own use: one env, a buffer of whole episodes, last next_value = 0 (ppo.py:77)."""
    import uavenv_b200 as ub
    T = 300
    g = torch.Generator(device="cuda").manual_seed(5)
    r = torch.rand(T, device="cuda", generator=g)
    v = torch.randn(T, device="cuda", generator=g)
    d = torch.zeros(T, dtype=torch.bool, device="cuda")
    d[[99, 180, 299]] = True
    ret, adv = ub.compute_gae(r, v, d)
    ret_ref, adv_ref = _reference_gae(r[:, None], v[:, None], d[:, None], None, 0.998, 0.95)
    adv_ref = adv_ref[:, 0]
    assert ret.shape == (T,)
    assert torch.allclose(ret, ret_ref[:, 0], rtol=1e-5, atol=1e-4)
    assert torch.allclose(adv, (adv_ref - adv_ref.mean()) / (adv_ref.std() + 1e-7), rtol=1e-4, atol=5e-5)
    # terminal steps do not bootstrap: return == reward there up to the GAE tail, i.e. delta = r - v
    assert torch.allclose(ret[299], r[299], atol=1e-6)
