#!/usr/bin/env python
"""Multi-GPU consistency of the graph-replayed PPO update: after every update the parameters of all ranks must be
bit-identical (same all-reduced gradient, same Adam step).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 profiles/ddp_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import uavenv_b200 as ub
from target_allocation_ppo_transformer_b200 import parallel
rank, local_rank, world = parallel.init("nccl")
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
B, T = 2048, 16
env = ub.UAVEnvBatched(B, device=dev, seed=3, env_id_base=rank * B)
agent = ub.PPOAgent(B, T, dev, fused_rollout=True, env_id_base=rank * B, seed=3, minibatch_size=4096, update_precision="fused", graph_update=True)
obs = env.reset()
for it in range(3):
    while not agent.full():
        a = agent.select_action(obs); obs, r, d, _ = env.step(a); agent.store_transition(r, d)
    stats = agent.update(obs)
    w = torch.cat([p.detach().flatten() for p in agent.policy.parameters()])
    chk = torch.stack([w.double().sum(), (w.double() ** 2).sum()])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    if rank == 0:
        same = all(torch.equal(allc[0], c) for c in allc)
        print("iter", it, "graph" if agent._graph is not None else "eager", "params identical across ranks:", same, stats)
agent.close(); env.close()
dist.destroy_process_group()
