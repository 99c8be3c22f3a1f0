"""Scene / decision diagnostics with the content of the reference's print-and-plot scripts, as data.

analyze_environment_difficulty  <- main.py:9-103: `num_rounds` freshly generated scenes, the full [N,M] matrices of
    mechanics.calc_advantage (p_final, p_damage), the mean over all pairs ("random matching"), the mean over targets of
    the best UAV per target ("optimal matching" = the ceiling a perfect policy can reach), the penetration ratio and
    the script's three-way verdict.  Here all rounds are one batch and the matrices come from `score_matrix_kernel`.
record_decisions                <- test_visualize.py:15-48: one episode per env under a policy, every Assign decision as
    (uav id, target id) plus whether the accept rule kept it, and the scene geometry the script plots.
No plotting (matplotlib is not a dependency); the returned arrays are what the plots are drawn from.
"""
import numpy as np
import torch


def analyze_environment_difficulty(num_rounds=10, device="cuda", seed=None, config=None, env=None, verbose=True):
    """Returns {"avg_p_dmg", "avg_p_final", "best_p_dmg", "best_p_final", "pen_rate": [num_rounds] per scene,
    "hard_scenes": indices with best_p_final < 0.2 (main.py:78), "summary": {...}, "verdict": str}."""
    from .envs.uav_env import UAVEnvBatched
    own = env is None
    if own:
        env = UAVEnvBatched(num_rounds, device=device, seed=seed, config=config)
        env.reset(full_reset=True)                                    # main.py:31: a new random scene per round
    pf, pd = env.score_matrix(torch.float64)                          # [B,N,M]  main.py:38-45
    avg_dmg, avg_final = pd.mean(dim=(1, 2)), pf.mean(dim=(1, 2))     # :49-50 all possible pairs
    best_dmg = pd.max(dim=1).values.mean(dim=1)                       # :59-63 best UAV per target, averaged over targets
    best_final = pf.max(dim=1).values.mean(dim=1)
    pen = best_final / (best_dmg + 1e-6)                              # :73
    out = {"avg_p_dmg": avg_dmg.cpu().numpy(), "avg_p_final": avg_final.cpu().numpy(), "best_p_dmg": best_dmg.cpu().numpy(),
           "best_p_final": best_final.cpu().numpy(), "pen_rate": pen.cpu().numpy()}
    out["hard_scenes"] = np.nonzero(out["best_p_final"] < 0.2)[0]
    s = {"rounds": int(pf.shape[0]), "best_p_dmg": float(out["best_p_dmg"].mean()), "avg_p_dmg": float(out["avg_p_dmg"].mean()),
         "best_p_final": float(out["best_p_final"].mean()), "avg_p_final": float(out["avg_p_final"].mean())}
    s["penetration"] = s["best_p_final"] / (s["best_p_dmg"] + 1e-6)   # :94
    out["summary"] = s
    if s["penetration"] < 0.5:                                        # :97-105
        out["verdict"] = "penetration below 50%: interceptors / no-fly zones too dense or too strong"
    elif s["best_p_dmg"] < 0.5:
        out["verdict"] = "base damage probability too low: weather loss too high or targets too far from the UAV spawn area"
    else:
        out["verdict"] = "environment parameters look healthy: a solution exists in theory"
    if verbose:
        c = env.cfg
        print("environment check | map %gx%g | UAVs %d | targets %d" % (c.MAP_WIDTH, c.MAP_HEIGHT, env.N, env.M))
        for i in range(s["rounds"]):
            print("[round %02d] random pairing: P_dmg=%.3f P_final=%.3f | best UAV per target: P_dmg=%.3f P_final=%.3f | penetration %.1f%%%s" % (
                i + 1, out["avg_p_dmg"][i], out["avg_p_final"][i], out["best_p_dmg"][i], out["best_p_final"][i], 100 * out["pen_rate"][i],
                "  [!] very hard scene" if out["best_p_final"][i] < 0.2 else ""))
        print("summary over %d rounds: P_dmg best %.3f / mean %.3f; P_final best %.3f / mean %.3f; penetration %.1f%%" % (
            s["rounds"], s["best_p_dmg"], s["avg_p_dmg"], s["best_p_final"], s["avg_p_final"], 100 * s["penetration"]))
        print("verdict:", out["verdict"])
    if own:
        env.close()
    return out


@torch.no_grad()
def record_decisions(env, act_fn, max_steps=None):
    """One episode per env of `env` (a freshly reset UAVEnvBatched) under `act_fn(obs) -> int64 actions [B]`.
    Returns {"steps": T, "uav_id", "target_id", "action", "accepted": int arrays [T,B] (entries after an env's episode
    ended are -1), "assignments": per env the list of (uav id, target id) of every Assign decision (test_visualize.py:37-43),
    "scene": get_scene() of all envs (positions, values, radii: what plot_results draws)}."""
    B, N, M = env.num_envs, env.N, env.M
    scene = env.get_scene()
    tgt_id = scene["tgt_id"]                                           # list order -> Target.id (the list is shuffled, uav_env.py:173)
    limit = N * M if max_steps is None else int(max_steps)
    alive = np.ones(B, dtype=bool)
    rows = {k: [] for k in ("uav_id", "target_id", "action", "accepted")}
    obs = env.obs
    for _ in range(limit):
        st = env.get_state()
        k, m = st["uav_idx"].astype(np.int64), st["target_idx"].astype(np.int64)
        a = act_fn(obs)
        obs, _, done, info = env.step(a)
        a_h, d_h, v_h = a.cpu().numpy(), done.cpu().numpy().astype(bool), info["is_valid_action"].cpu().numpy()
        rows["uav_id"].append(np.where(alive, k, -1))                  # UAV ids are list positions (uav_env.py:95)
        rows["target_id"].append(np.where(alive, tgt_id[np.arange(B), np.minimum(m, M - 1)], -1))
        rows["action"].append(np.where(alive, a_h, -1))
        rows["accepted"].append(np.where(alive & (a_h == 1), (v_h == 1).astype(np.int64), np.where(alive, 0, -1)))
        alive &= ~d_h
        if not alive.any():
            break
    out = {key: np.stack(v).astype(np.int64) for key, v in rows.items()}
    out["steps"] = len(rows["action"])
    out["assignments"] = [[(int(u), int(t)) for u, t, a in zip(out["uav_id"][:, b], out["target_id"][:, b], out["action"][:, b]) if a == 1]
                          for b in range(B)]
    out["scene"] = scene
    return out
