"""Batched training loop: the rollout path of main_train.py:35-232 on B device-resident envs.

    python -m torch.distributed.run --nproc-per-node 8 -m ... (or plain python) :  see main()

Per iteration: T fused env steps with actions sampled from policy_old (main_train.py:109-119), then one PPO
update (main_train.py:145-146 -> agents/ppo.py:68-181).  The reset schedule of main_train.py:79 lives inside
the env.  Logging keeps the reference's CSV columns (main_train.py:52-63); checkpoints are plain
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

CSV_COLUMNS = ["Episode", "Avg_Reward", "Avg_J_Val", "Max_Coverage", "Q0_Value", "Action1_Ratio", "Valid_Assign_Rate",
               "Avg_P_Dmg", "Avg_P_Final", "Loss_Critic", "Loss_Actor", "Entropy"]      # main_train.py:52-63


def train(num_envs=16384, horizon=32, iterations=10, cfg=None, log_dir=None, seed=None, minibatch_size=None,
          save_every=0, verbose=True, fused_rollout=True, update_precision="fused", graph_update=True):
    """Returns a list of per-iteration stat dicts (rank 0 also writes training_stats.csv / checkpoints)."""
    cfg = cfg or global_cfg
    rank, local_rank, world = parallel.init()
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
    obs = env.reset(full_reset=True)
    ep_return = torch.zeros(num_envs, dtype=torch.float64, device=device)
    history, episodes_done = [], 0
    for it in range(1, iterations + 1):
        t0 = time.perf_counter()
        acc = torch.zeros(8, dtype=torch.float64, device=device)   # see the unpacking below
        max_cov = torch.zeros((), dtype=torch.int32, device=device)
        with torch.no_grad():
            _, q0 = agent.policy_old.logits_and_value(obs)          # Q0 of the current states (main_train.py:87-96)
        while not agent.full():
            action = agent.select_action(obs)
            obs, reward, done, info = env.step(action)
            agent.store_transition(reward, done)
            ep_return += reward
            assign = action == 1
            valid = info["is_valid_action"] == 1
            covered = info["num_assigned"] > 0
            acc += torch.stack([(ep_return * done).sum(), done.sum().double(), info["J_val"].double().sum(),
                                assign.sum().double(), valid.sum().double(),
                                (info["avg_p_dmg"].double() * covered).sum(), (info["avg_p_final"].double() * covered).sum(),
                                covered.sum().double()])
            max_cov = torch.maximum(max_cov, info["num_assigned"].max())
            ep_return.masked_fill_(done, 0.0)
        stats = agent.update(obs)
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        a = acc.clone()
        if world > 1:
            torch.distributed.all_reduce(a)
        a = a.tolist()
        steps = num_envs * horizon * world
        episodes_done += int(a[1])
        row = {"Episode": episodes_done, "Avg_Reward": a[0] / max(a[1], 1.0), "Avg_J_Val": a[2] / steps,
               "Max_Coverage": int(max_cov.item()), "Q0_Value": float(q0.mean().item()),
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
                    it, episodes_done, row["Avg_Reward"], row["Avg_J_Val"], row["Entropy"], row["samples_per_sec"]))
            if log_dir and save_every and it % save_every == 0:
                torch.save(agent.policy.state_dict(), os.path.join(log_dir, "model_iter%d.pth" % it))
    if rank == 0 and log_dir:
        torch.save(agent.policy.state_dict(), os.path.join(log_dir, "final_model.pth"))   # main_train.py:231-232
        fh.close()
    agent.close()
    env.close()
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
