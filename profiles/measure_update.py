#!/usr/bin/env python
"""Where one PPO minibatch step (evaluate + loss + backward) of the update spends its time: kernel table from
torch.profiler plus a CUDA-event total.      python profiles/measure_update.py [minibatch] [tf32|fp32|bf16|fused]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uavenv_b200 as ub  # noqa: E402


def main():
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(pos[0]) if pos else 131072
    prec = pos[1] if len(pos) > 1 else "tf32"
    torch.backends.cuda.matmul.allow_tf32 = prec != "fp32"
    net = ub.TransformerActorCritic().cuda()
    obs = torch.rand(n, 5, 14, device="cuda")
    obs[: n // 4, :3] = 0                      # some padded windows, as early in an episode
    act = torch.randint(0, 2, (n,), device="cuda")

    trunks = None
    if prec == "fused":
        from target_allocation_ppo_transformer_b200.networks.fused_train import FusedTrunks
        trunks = FusedTrunks(n, "cuda")

    def step():
        for p in net.parameters():
            p.grad = None
        if trunks is not None:
            lp, v, e = trunks.evaluate(net, obs, act)
        else:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=prec == "bf16"):
                lp, v, e = net.evaluate(obs, act)
        (lp.float().mean() + v.float().mean() + e.float().mean()).backward()

    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("minibatch %d, %s: %.2f ms per evaluate+backward" % (n, prec, e0.elapsed_time(e1) / 5))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))
    if "--timeline" in sys.argv:                 # every kernel launch of the step, in launch order
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        for e in evs:
            print("%9.1f us  %s" % (e.time_range.elapsed_us(), e.name[:110]))


if __name__ == "__main__":
    main()
