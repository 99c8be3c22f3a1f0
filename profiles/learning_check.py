#!/usr/bin/env python
"""Does the hand-written bf16 update train like the fp32/TF32 PyTorch update?  Same seed, 4096 envs x horizon 32, 40
iterations each; prints (iteration, Avg_Reward, Avg_J_Val, Entropy, Loss_Critic) every 4 iterations.
    python profiles/learning_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uavenv_b200 as ub  # noqa: E402

for prec, graph in (("fused", True), ("tf32", False)):
    torch.manual_seed(0)
    hist = ub.train(num_envs=4096, horizon=32, iterations=40, verbose=False, seed=7, update_precision=prec, graph_update=graph)
    print(prec, "samples/s in the last iteration: %.3g" % hist[-1]["samples_per_sec"])
    rows = [(h["iteration"], round(h["Avg_Reward"], 3), round(h["Avg_J_Value"], 3), round(h["Entropy"], 3), round(h["Loss_Critic"], 3)) for h in hist]
    for r in rows[::4] + [rows[-1]]:
        print("  ", r)
