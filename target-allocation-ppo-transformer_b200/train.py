"""Batched training loop: the rollout path of main_train.py:35-232 on B device-resident envs.

    python -m torch.distributed.run --nproc-per-node 8 -m ... (or plain python) :  see main()

Per iteration: T fused env steps with actions sampled from policy_old (main_train.py:109-119), then one PPO
update (main_train.py:145-146 -> agents/ppo.py:68-181).  The reset schedule of main_train.py:79 lives inside
the env.  Logging keeps the reference's CSV header (main_train.py:57-63, same names and order; one row per iteration
instead of one per 10 episodes; Avg_Q0 = mean V(s_0) over the episodes that STARTED in the iteration, main_train.py:87-96;
Max_Coverage is the maximum over all ranks); checkpoints are plain
`policy.state_dict()` files (main_train.py:209-212, :231-232) that the reference network loads unchanged.
"""
import csv
import os
import sys
import time

import torch

from . import parallel
from .agents.ppo import PPOAgent
from .configs.config import cfg as global_cfg
from .envs.uav_env import UAVEnvBatched

CSV_COLUMNS = ["Episode", "Avg_Reward", "Avg_Q0", "Avg_J_Value", "Max_Coverage", "Action1_Ratio", "Valid_Assign_Rate",
               "Avg_P_Dmg", "Avg_P_Final", "Loss_Critic", "Loss_Actor", "Entropy"]      # main_train.py:57-63, same names and order


def train(num_envs=16384, horizon=32, iterations=10, cfg=None, log_dir=None, seed=None, minibatch_size=None,
          save_every=0, verbose=True, fused_rollout=True, update_precision="fused", graph_update=True):
    """Returns a list of per-iteration stat dicts (rank 0 also writes training_stats.csv / checkpoints)."""
    cfg = cfg or global_cfg
    owns_group = not (torch.distributed.is_available() and torch.distributed.is_initialized())
    rank, local_rank, world = parallel.init()
    owns_group = owns_group and world > 1       # this call created the process group, so it also destroys it
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    seed = cfg.SEED if seed is None else seed
    env = UAVEnvBatched(num_envs, device=device, seed=seed, env_id_base=rank * num_envs, config=cfg)
    agent = PPOAgent(num_envs, horizon, device, cfg=cfg, minibatch_size=minibatch_size, seed=seed,
                     fused_rollout=fused_rollout, env_id_base=rank * num_envs, update_precision=update_precision,
                     graph_update=graph_update)
    writer = fh = None
    if rank == 0 and log_dir:
        os.makedirs(log_dir, exist_ok=True)
        fh = open(os.path.join(log_dir, "training_stats.csv"), "w", newline="")
        writer = csv.writer(fh)
        writer.writerow(CSV_COLUMNS)
    history, episodes_done = [], 0
    try:
        obs = env.reset(full_reset=True)
        ep_return = torch.zeros(num_envs, dtype=torch.float64, device=device)
        for it in range(1, iterations + 1):
            t0 = time.perf_counter()
            acc = torch.zeros(10, dtype=torch.float64, device=device)   # see the unpacking below
            max_cov = torch.zeros((), dtype=torch.int32, device=device)
            while not agent.full():
                first = obs[:, -2].abs().sum(-1) == 0                 # window holds one real row: s_0 of an episode
                action = agent.select_action(obs)
                q0 = agent.buf_value[agent.t]                         # V(s) the rollout forward just stored
                obs, reward, done, info = env.step(action)
                agent.store_transition(reward, done)
                ep_return += reward
                assign = action == 1
                valid = info["is_valid_action"] == 1
                covered = info["num_assigned"] > 0
                acc += torch.stack([(ep_return * done).sum(), done.sum().double(), info["J_val"].double().sum(),
                                    assign.sum().double(), valid.sum().double(),
                                    (info["avg_p_dmg"].double() * covered).sum(), (info["avg_p_final"].double() * covered).sum(),
                                    covered.sum().double(), (q0.double() * first).sum(), first.sum().double()])
                max_cov = torch.maximum(max_cov, info["num_assigned"].max())
                ep_return.masked_fill_(done, 0.0)
            stats = agent.update(obs)
            torch.cuda.synchronize(device)
            dt = time.perf_counter() - t0
            a = acc.clone()
            if world > 1:
                torch.distributed.all_reduce(a)
                torch.distributed.all_reduce(max_cov, op=torch.distributed.ReduceOp.MAX)
            a = a.tolist()
            steps = num_envs * horizon * world
            episodes_done += int(a[1])
            row = {"Episode": episodes_done, "Avg_Reward": a[0] / max(a[1], 1.0), "Avg_Q0": a[8] / max(a[9], 1.0),
                   "Avg_J_Value": a[2] / steps, "Max_Coverage": int(max_cov.item()),
                   "Action1_Ratio": a[3] / steps, "Valid_Assign_Rate": a[4] / max(a[3], 1.0),
                   "Avg_P_Dmg": a[5] / max(a[7], 1.0), "Avg_P_Final": a[6] / max(a[7], 1.0),
                   "Loss_Critic": stats["loss_critic"], "Loss_Actor": stats["loss_actor"], "Entropy": stats["entropy"],
                   "samples_per_sec": steps / dt, "iteration": it}
            history.append(row)
            if rank == 0:
                if writer:
                    writer.writerow([row[c] for c in CSV_COLUMNS]); fh.flush()
                if verbose:
                    print("iter %d  episodes %d  avg_reward %.3f  J %.3f  entropy %.3f  %.3g samples/s" % (
                        it, episodes_done, row["Avg_Reward"], row["Avg_J_Value"], row["Entropy"], row["samples_per_sec"]))
                if log_dir and save_every and it % save_every == 0:
                    torch.save(agent.policy.state_dict(), os.path.join(log_dir, "model_iter%d.pth" % it))
        if rank == 0 and log_dir:
            torch.save(agent.policy.state_dict(), os.path.join(log_dir, "final_model.pth"))   # main_train.py:231-232
    finally:
        if fh:
            fh.close()
        agent.close()           # the captured update graph holds NCCL kernels: drop it before the process group
        env.close()
        if owns_group and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    return history


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--horizon", type=int, default=32)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--log-dir", default=None)
    args = ap.parse_args()
    train(args.envs, args.horizon, args.iterations, log_dir=args.log_dir)


class _CallableModule(sys.modules[__name__].__class__):
    """`package.train` names both this module and the function in it: calling the module runs train()."""

    def __call__(self, *args, **kwargs):
        return train(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule

if __name__ == "__main__":
    main()
