"""Worker of tests/test_gpu_ddp.py (one process per GPU, NCCL): after every graph-replayed PPO update the parameters
of all ranks must be bit-identical - same all-reduced gradient, same fixed-order norm, same Adam step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/ddp_worker.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import uavenv_b200 as ub
from target_allocation_ppo_transformer_b200 import parallel


def main():
    rank, local_rank, world = parallel.init("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, T = 2048, 16
    env = ub.UAVEnvBatched(B, device=dev, seed=3, env_id_base=rank * B)
    agent = ub.PPOAgent(B, T, dev, fused_rollout=True, env_id_base=rank * B, seed=3, minibatch_size=4096,
                        update_precision="fused", graph_update=True)
    obs = env.reset()
    w_init = agent._flat_params.clone()
    ok = True
    for it in range(3):
        while not agent.full():
            a = agent.select_action(obs)
            obs, r, d, _ = env.step(a)
            agent.store_transition(r, d)
        stats = agent.update(obs)
        w = agent._flat_params
        gathered = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        same = all(torch.equal(gathered[0], g) for g in gathered)          # bit-identical replicas
        # the ranks saw DIFFERENT data (env shards differ), so identical weights prove the all-reduce happened
        obs_sum = torch.stack([obs.double().sum()])
        sums = [torch.empty_like(obs_sum) for _ in range(world)]
        dist.all_gather(sums, obs_sum)
        differ = len({float(s) for s in sums}) == world
        moved = not torch.equal(w, w_init)
        ok = ok and same and differ and moved
        if rank == 0:
            print("iter", it, "graph" if agent._graph is not None else "eager", "identical:", same, "shards differ:", differ,
                  "stats", stats, flush=True)
    ok = ok and agent._graph is not None        # the updates really ran as graph replays (NCCL kernel captured)
    ar_us = agent.time_gradient_allreduce(20)
    agent.close()
    env.close()
    dist.destroy_process_group()
    if rank == 0:
        print("allreduce_us %.1f" % ar_us)
    print("rank %d %s" % (rank, "ok" if ok else "FAILED"), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
