#!/usr/bin/env python
"""Golden vectors of the reference policy network (networks/transformer_net.py) and PPO maths (agents/ppo.py).
TEST INFRASTRUCTURE ONLY; run in the build container:  python oracle/gen_golden_policy.py
Writes tests/golden/policy_net.npz: the reference TransformerActorCritic's state_dict (seeded init, then perturbed
so that biases / LayerNorm affine terms are non-trivial), a batch of observation windows with realistic padding
patterns, and the reference outputs (logits via evaluate's log-probs, values, entropy); plus the GAE returns /
normalised advantages the reference PPOAgent.update computes on an episodic buffer (ppo.py:77-94)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    _, RefNet = ref_shim.load_agents()
    torch.manual_seed(1234)
    net = RefNet()
    with torch.no_grad():
        for name, p in net.named_parameters():
            if p.dim() == 1:                               # biases, LayerNorm weight/bias: make them non-trivial
                p.add_(0.05 * torch.randn_like(p))
    net.eval()
    g = torch.Generator().manual_seed(7)
    B = 96
    x = torch.rand(B, 5, 14, generator=g)
    x[:, :, 13] = 1.0
    for b in range(B):                                      # episode starts: 0..4 leading zero (padding) rows
        x[b, : (b % 5)] = 0.0
    x[-1] = 0.0                                             # an all-zero window (only the last row is unmasked)
    actions = torch.randint(0, 2, (B,), generator=g)
    with torch.no_grad():
        logp, value, ent = net.evaluate(x, actions)
        xa = net.actor_net(x)[:, -1]
        logits = net.actor_head(xa)
    sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, "policy_net.npz"), obs=x.numpy(), actions=actions.numpy(),
                        logp=logp.numpy(), value=value.numpy(), entropy=ent.numpy(), logits=logits.numpy(),
                        keys=np.array(list(sd.keys())), **{"p::" + k: v for k, v in sd.items()})
    print("policy_net.npz: %d tensors, %d parameters" % (len(sd), sum(v.size for v in sd.values())))


if __name__ == "__main__":
    main()
