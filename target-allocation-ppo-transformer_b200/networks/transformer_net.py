"""Policy / value network with the reference's parameters and numerics (networks/transformer_net.py).

`TransformerActorCritic` here is NOT built from torch.nn.TransformerEncoder: the forward pass is written out
explicitly (embedding + learned positions, post-LN encoder layers with a key-padding mask, last-token readout,
MLP heads), which is what the fused sm_100a forward (csrc/policy_forward.cu) implements as well.  The module
tree carries the SAME parameter names and shapes as the reference, so `state_dict()`s are interchangeable in
both directions (checkpoints of the reference load here and vice versa; tested against golden vectors
recorded from the reference network).

Architecture (configs/config.py:61-68, transformer_net.py:21-91): STATE_DIM 14 -> EMBED_DIM 128, SEQ_LEN 5,
8 heads, FFN 256, ReLU, dropout 0; actor = 1 layer, critic = 2 layers; heads 128 -> 64 -> {2, 1}.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..configs.config import cfg as _cfg


def _orthogonal_(linear, std=math.sqrt(2.0), bias=0.0):  # init_layer of transformer_net.py:9-12
    nn.init.orthogonal_(linear.weight, std)
    nn.init.constant_(linear.bias, bias)
    return linear


class _SelfAttention(nn.Module):
    """Parameter layout of nn.MultiheadAttention: packed in_proj_weight [3D, D] / in_proj_bias [3D] + out_proj."""

    def __init__(self, dim, heads):
        super().__init__()
        self.dim, self.heads = dim, heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * dim, dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * dim))
        self.out_proj = nn.Linear(dim, dim)
        self.fused_attention = True        # use csrc/ppo_attn.cu on CUDA fp32 tensors (same math, one kernel each way)
        nn.init.xavier_uniform_(self.in_proj_weight)           # torch's MultiheadAttention._reset_parameters
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, x, pad_mask):
        b, s, d = x.shape
        h, dh = self.heads, d // self.heads
        qkv = F.linear(x, self.in_proj_weight, self.in_proj_bias)
        if self.fused_attention and d == 128 and h == 8:
            from . import attn_op                                     # fused CUDA forward/backward (csrc/ppo_attn.cu)
            q3, k3, v3 = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
            if attn_op.usable(q3, k3, v3):
                return self.out_proj(attn_op.attention5(q3, k3, v3, pad_mask))
        qkv = qkv.view(b, s, 3, h, dh)
        q, k, v = qkv[:, :, 0].transpose(1, 2), qkv[:, :, 1].transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
        scores = (q @ k.transpose(-1, -2)) / math.sqrt(dh)                       # [b, h, s, s]
        scores = scores.masked_fill(pad_mask[:, None, None, :], float("-inf"))   # keys that are padding rows
        out = torch.softmax(scores, dim=-1) @ v                                  # [b, h, s, dh]
        return self.out_proj(out.transpose(1, 2).reshape(b, s, d))

    def forward_last(self, x, pad_mask):
        """Attention output of the NEWEST token only (all that the last encoder layer contributes downstream):
        K/V of the five tokens, Q of one - the formulation csrc/policy_forward.cu executes."""
        b, s, d = x.shape
        h, dh = self.heads, d // self.heads
        q = F.linear(x[:, -1], self.in_proj_weight[:d], self.in_proj_bias[:d])
        kv = F.linear(x, self.in_proj_weight[d:], self.in_proj_bias[d:])
        if self.fused_attention and d == 128 and h == 8:
            from . import attn_op
            q3, k3, v3 = q.view(b, 1, d), kv[..., :d], kv[..., d:]
            if attn_op.usable(q3, k3, v3):
                return self.out_proj(attn_op.attention5(q3, k3, v3, pad_mask).view(b, d))
        q = q.view(b, h, 1, dh)
        kv = kv.view(b, s, 2, h, dh)
        k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
        scores = (q @ k.transpose(-1, -2)) / math.sqrt(dh)                       # [b, h, 1, s]
        scores = scores.masked_fill(pad_mask[:, None, None, :], float("-inf"))
        return self.out_proj((torch.softmax(scores, dim=-1) @ v).reshape(b, d))


class _EncoderLayer(nn.Module):
    """Post-LN layer: x = LN1(x + attn(x)); x = LN2(x + W2 relu(W1 x))  (nn.TransformerEncoderLayer defaults)."""

    def __init__(self, dim, heads, ffn):
        super().__init__()
        self.self_attn = _SelfAttention(dim, heads)
        self.linear1 = nn.Linear(dim, ffn)
        self.linear2 = nn.Linear(ffn, dim)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)

    def forward(self, x, pad_mask):
        x = self.norm1(x + self.self_attn(x, pad_mask))
        return self.norm2(x + self.linear2(torch.relu(self.linear1(x))))

    def forward_last(self, x, pad_mask):
        y = self.norm1(x[:, -1] + self.self_attn.forward_last(x, pad_mask))
        return self.norm2(y + self.linear2(torch.relu(self.linear1(y))))


class _Encoder(nn.Module):
    def __init__(self, dim, heads, ffn, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(dim, heads, ffn) for _ in range(num_layers)])

    def forward(self, x, pad_mask):
        for layer in self.layers:
            x = layer(x, pad_mask)
        return x

    def forward_last(self, x, pad_mask):
        for layer in self.layers[:-1]:
            x = layer(x, pad_mask)
        return self.layers[-1].forward_last(x, pad_mask)


class TransformerBlock(nn.Module):
    """transformer_net.py:15-65: Linear(14,128)+ReLU embedding, learned positions, encoder; zero rows are padding."""

    def __init__(self, num_layers=None, cfg=None):
        super().__init__()
        c = cfg or _cfg
        num_layers = c.NUM_LAYERS if num_layers is None else num_layers
        self.embedding = nn.Sequential(_orthogonal_(nn.Linear(c.STATE_DIM, c.EMBED_DIM)), nn.ReLU())
        self.pos_embedding = nn.Parameter(torch.randn(1, c.SEQ_LEN, c.EMBED_DIM) * 0.02)
        self.transformer = _Encoder(c.EMBED_DIM, c.NUM_HEADS, 256, num_layers)

    def forward(self, x):
        pad = x.abs().sum(dim=-1) == 0          # transformer_net.py:52
        pad[:, -1] = False                      # :54 the newest row is never masked
        h = self.embedding(x) + self.pos_embedding[:, : x.size(1)]
        return self.transformer(h, pad)

    def forward_last(self, x):
        """== forward(x)[:, -1] (the only row the heads read, transformer_net.py:106,114) at ~half the work."""
        pad = x.abs().sum(dim=-1) == 0
        pad[:, -1] = False
        h = self.embedding(x) + self.pos_embedding[:, : x.size(1)]
        return self.transformer.forward_last(h, pad)


class TransformerActorCritic(nn.Module):
    """transformer_net.py:68-143: get_action(state) -> (action, logp, value, entropy); evaluate(state, action)."""

    def __init__(self, cfg=None):
        super().__init__()
        c = cfg or _cfg
        self.hidden_dim = c.EMBED_DIM
        self.actor_net = TransformerBlock(1, c)
        self.actor_head = nn.Sequential(_orthogonal_(nn.Linear(c.EMBED_DIM, 64)), nn.ReLU(),
                                        _orthogonal_(nn.Linear(64, c.ACTION_DIM), std=0.01))
        self.critic_net = TransformerBlock(2, c)
        self.critic_head = nn.Sequential(_orthogonal_(nn.Linear(c.EMBED_DIM, 64)), nn.ReLU(),
                                         _orthogonal_(nn.Linear(64, 1), std=1.0))

    def forward(self, state):
        raise NotImplementedError("use get_action or evaluate")    # as the reference (:93-94)

    def logits_and_value(self, state):
        if state.dim() == 2:
            state = state.unsqueeze(0)
        logits = self.actor_head(self.actor_net.forward_last(state))
        value = self.critic_head(self.critic_net.forward_last(state))
        return logits, value

    def get_action(self, state, generator=None):
        logits, value = self.logits_and_value(state)
        logp_all = torch.log_softmax(logits, dim=-1)
        probs = logp_all.exp()
        action = torch.multinomial(probs, 1, generator=generator).squeeze(-1)
        entropy = -(probs * logp_all).sum(-1)
        return action, logp_all.gather(-1, action[:, None]).squeeze(-1), value, entropy

    def evaluate(self, state, action):
        logits, value = self.logits_and_value(state)
        logp_all = torch.log_softmax(logits, dim=-1)
        entropy = -(logp_all.exp() * logp_all).sum(-1)
        return logp_all.gather(-1, action[:, None]).squeeze(-1), value, entropy
