/*
 * uavenv_b200.h - C ABI of the B200-native batched UAV->target allocation environment.
 *
 * This is the drop-in boundary for the rollout path of
 * Dingyf717/target-allocation-ppo-transformer.  Each entry point names the
 * reference interface it replaces (file:line under the reference tree).  The reference is pure
 * Python, so the binding a maintainer adds is a ctypes stub; the same symbols can
 * be bound from any FFI: plain pointers, sizes and C structs only - no torch / C++ types.
 *
 * Conventions
 *   - every function returns 0 on success and a negative UAVENV_E* code on failure; the message is
 *     available from uavenv_last_error().  No C++ exception crosses the boundary.
 *   - pointers prefixed d_ are DEVICE pointers on the handle's device, h_ are HOST pointers.
 *   - all device work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as
 *     void*; NULL = the legacy default stream).  No call synchronises unless stated.
 *   - a handle is not thread-safe; use one handle per device / per host thread.
 *   - B = num_envs, N = num_uavs, M = num_targets, K1 = num_nfz, K2 = num_interceptors.
 *   - there is no CPU fallback: creation fails when no CUDA device is usable.
 */
#ifndef UAVENV_B200_H
#define UAVENV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAVENV_ABI_VERSION 2
#define UAVENV_STATE_DIM 14 /* configs/config.py:61 STATE_DIM */
#define UAVENV_SEQ_LEN 5    /* configs/config.py:62 SEQ_LEN   */

enum {
    UAVENV_OK = 0,
    UAVENV_EINVAL = -1,  /* bad argument                         */
    UAVENV_ECUDA = -2,   /* a CUDA runtime call / launch failed  */
    UAVENV_ENOMEM = -3,  /* host or device allocation failed     */
    UAVENV_ESTATE = -4   /* call not valid in the current state  */
};

/* The constants the path reads from the reference's `cfg` singleton (configs/config.py:7-12,33-58,
 * 61-63,83) plus the batching knobs the reference has no need for. */
typedef struct uavenv_cfg {
    int32_t num_uavs;         /* NUM_UAVS            :42 */
    int32_t num_targets;      /* NUM_TARGETS         :43 */
    int32_t num_nfz;          /* NUM_NFZ             :48 */
    int32_t num_interceptors; /* NUM_INTERCEPTORS    :49 */
    int32_t reset_episodes;   /* RESET_EPISODES      :83 ; main_train.py:79 schedule. <=0: never regenerate */
    int32_t auto_reset;       /* 1: finished envs restart inside the same step launch (batched use);
                                 0: they stay finished until uavenv_reset (reference single-env contract) */
    double param_zeta_d;      /* PARAM_ZETA_D        :7  */
    double param_k;           /* PARAM_K             :8  */
    double param_c1, param_c2, param_c3, param_c4; /* :9-12 */
    double cost_weight_omega; /* COST_WEIGHT_OMEGA   :53 */
    double weather_speed_factor, weather_load_factor; /* :57-58 */
    double map_width, map_height;                  /* :33-34 */
    double uav_gen_x_lo, uav_gen_x_hi;             /* UAV_GEN_X_RANGE    :39 */
    double target_gen_x_lo, target_gen_x_hi;       /* TARGET_GEN_X_RANGE :40 */
    double intercept_rad;                          /* INTERCEPT_RAD      :50 */
    double tie_band;          /* Eq.21 `new_r >= prev_r` (envs/uav_env.py:317) is decided from carried sums; when the two
                                 rewards agree to within tie_band * max(|new_r|, |prev_r|) both are re-summed over the targets
                                 in the reference's list order (uav_env.py:244-269) and THAT decides.  Default 1e-12 (the carried
                                 sums are good to ~1e-14); a huge value takes the exact path always (tests) */
} uavenv_cfg_t;

/* Per-step diagnostics = the `info` dict of envs/uav_env.py:426-433, one value per env.
 * Every pointer is a device array of length B and may be NULL (not wanted). */
typedef struct uavenv_info {
    float *d_J_val;           /* "J_val"           */
    int32_t *d_num_assigned;  /* "num_assigned" = covered-target count N0 */
    int8_t *d_is_valid_action;/* "is_valid_action": -1 = None, 0 = False, 1 = True */
    float *d_avg_p_dmg;       /* "avg_p_dmg"       */
    float *d_avg_p_final;     /* "avg_p_final"     */
    double *d_reward_f64;     /* extension: the step reward before the cast to float */
} uavenv_info_t;

/* Scene of `count` consecutive envs as HOST structure-of-arrays in the reference's list order
 * (targets AFTER the shuffle of envs/uav_env.py:173), env-major: field[e*N + i].
 * = the public attributes env.uavs / env.targets / env.nfz_list / env.interceptors
 * (envs/uav_env.py:27-30, envs/entities.py:13-61).  On load, uav_type / nfz_radius may be NULL. */
typedef struct uavenv_scene {
    double *uav_x, *uav_y, *uav_vx, *uav_vy, *uav_load, *uav_cost; /* [count*N]  */
    int32_t *uav_type;                                               /* [count*N]  */
    double *tgt_x, *tgt_y, *tgt_vx, *tgt_vy, *tgt_value;            /* [count*M]  */
    int32_t *tgt_id;                                                 /* [count*M]  */
    double *nfz_x, *nfz_y, *nfz_radius;                              /* [count*K1] */
    double *int_x, *int_y, *int_vx, *int_vy;                         /* [count*K2] */
} uavenv_scene_t;

/* Mutable allocation state of `count` consecutive envs (HOST arrays; any pointer may be NULL).
 * = uav_idx / target_idx (envs/uav_env.py:33-34), UAV.assigned_target_id (entities.py:30),
 * len(Target.locked_by_uavs) (entities.py:46) and the running objective. */
typedef struct uavenv_state {
    int32_t *uav_idx, *target_idx;     /* [count]    */
    int32_t *assigned_target_id;       /* [count*N], -1 = unassigned */
    int32_t *lock_count;               /* [count*M], list order      */
    double *not_hit, *not_hit_pure;    /* [count*M]  prod(1-p_final), prod(1-p_damage) over the lock list */
    double *J_val;                     /* [count]    envs/uav_env.py:244-269 */
    int32_t *episode, *scene_index;    /* [count]    1-based episode counter, scenes generated so far */
    uint8_t *finished;                 /* [count]    only ever 1 with auto_reset = 0 */
} uavenv_state_t;

typedef struct uavenv uavenv_t;

/* configs/config.py defaults (the live "easy mode" file) with reset_episodes = 200, auto_reset = 1 */
void uavenv_default_cfg(uavenv_cfg_t *cfg);

/* replaces UAVEnv.__init__ (envs/uav_env.py:14-40).  Allocates every device array once;
 * `seed` keys the counter-based scene generator and `env_id_base` is the GLOBAL id of env 0 of
 * this handle, so a shard [base, base+B) reproduces the same envs on any GPU count. */
int uavenv_create(const uavenv_cfg_t *cfg, int32_t num_envs, int32_t device, uint64_t seed,
                  uint64_t env_id_base, uavenv_t **out);
int uavenv_destroy(uavenv_t *h);
/* message of the last failure on this handle (h may be NULL: last creation failure) */
const char *uavenv_last_error(const uavenv_t *h);
int uavenv_abi_version(void);
int32_t uavenv_num_envs(const uavenv_t *h);

/* replaces UAVEnv.reset(full_reset) (envs/uav_env.py:42-63) for the envs whose d_env_mask byte is
 * non-zero (NULL = all).  full_reset != 0: _generate_scene (uav_env.py:65-173) with the counter
 * RNG; 0: _reset_state_only (:175-182).  Writes the first observation window (4 zero rows + the
 * row of pointer pair (0,0)) to d_obs[B,5,14] for the reset envs; d_obs may be NULL. */
int uavenv_reset(uavenv_t *h, int32_t full_reset, const uint8_t *d_env_mask, float *d_obs, void *stream);

/* replaces UAVEnv.step(action) (envs/uav_env.py:295-435) for all B envs in ONE kernel launch.
 * d_actions[B] int64 (the dtype torch.distributions.Categorical.sample() yields): 1 = Assign,
 * anything else = Skip (:344).  Outputs: d_obs[B,5,14] f32, d_reward[B] f32, d_done[B] u8, info.
 * With auto_reset the observation of a finished env is the first window of its next episode
 * (reference: a zero row the training loop never reads, :188-189); without it the window is zero. */
int uavenv_step(uavenv_t *h, const int64_t *d_actions, float *d_obs, float *d_reward, uint8_t *d_done,
                const uavenv_info_t *info, void *stream);

/* same step through HOST buffers (the reference's caller lives on the host: main_train.py:111-113):
 * moves h_actions host->device, steps, moves reward/done back and synchronises the stream.  Pinned,
 * device-mapped host buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are read and written
 * in place by the fused kernel over PCIe; pageable buffers go through staging copies.
 * The observation window stays on the device (d_obs, or the handle's own buffer when NULL -
 * see uavenv_obs_buffer) where the policy consumes it. */
int uavenv_step_host(uavenv_t *h, const int64_t *h_actions, float *h_reward, uint8_t *h_done, float *d_obs,
                     void *stream);
/* the same with one byte per action (the action space is {0,1}, envs/uav_env.py:18): 8x less PCIe traffic */
int uavenv_step_host_i8(uavenv_t *h, const int8_t *h_actions, float *h_reward, uint8_t *h_done, float *d_obs,
                        void *stream);
float *uavenv_obs_buffer(uavenv_t *h); /* device [B,5,14] owned by the handle */

/* replaces assigning env.uavs / env.targets / env.nfz_list / env.interceptors by hand: injects
 * scenes (e.g. exported from the reference) into envs [first_env, first_env+count) and leaves them
 * in the reset(full_reset=False) state, episode 1; their first observation windows are written to
 * d_obs[B,5,14] (rows of the loaded envs only; NULL = the handle's own buffer).  Synchronous. */
int uavenv_load_scene(uavenv_t *h, const uavenv_scene_t *scene, int32_t first_env, int32_t count, float *d_obs);
/* replaces reading env.uavs / env.targets / ... (main.py:35-42, test_visualize.py:61-94). Synchronous. */
int uavenv_get_scene(uavenv_t *h, uavenv_scene_t *scene, int32_t first_env, int32_t count);
int uavenv_get_state(uavenv_t *h, uavenv_state_t *state, int32_t first_env, int32_t count);

/* restores the 1-based per-env episode counters that drive the main_train.py:79 regeneration
 * schedule (resume of a run; staggered steady state for benchmarks).  h_episode[count] is a HOST
 * array.  Synchronous. */
int uavenv_set_episode_counters(uavenv_t *h, const int32_t *h_episode, int32_t first_env, int32_t count);

/* replaces the double loop over mechanics.calc_advantage of main.py:38-45 (mechanics.py:167-181):
 * p_final / p_damage for every (env, UAV, target), [B,N,M] list order.  Either pair may be NULL. */
int uavenv_score_matrix(uavenv_t *h, float *d_p_final, float *d_p_damage, void *stream);
int uavenv_score_matrix_f64(uavenv_t *h, double *d_p_final, double *d_p_damage, void *stream);

/* re-derives every running aggregate (J, N0, cost / value sums) from the per-target products with
 * a warp per env, exactly as envs/uav_env.py:244-293 sums them; writes the largest |carried-fresh|
 * J difference seen to *h_max_abs_diff (may be NULL).  Synchronous.  Used to bound drift. */
int uavenv_recompute_objective(uavenv_t *h, double *h_max_abs_diff, void *stream);

/* Bernoulli(1/2) actions keyed (action_seed, step, global env id): the synthetic action stream of
 * the env-only benchmarks */
int uavenv_random_actions(uavenv_t *h, uint64_t action_seed, uint64_t step, int64_t *d_actions, void *stream);

/* ---- PPO rollout post-processing (agents/ppo.py:77-94) ------------------------------------------
 * GAE(gamma, lambda) over a [T,B] rollout (time-major) followed by the global advantage
 * normalisation (mean, unbiased std, +1e-7).  d_last_value[B] is V(s_T) for bootstrapping (NULL = 0,
 * the reference's episodic case, ppo.py:77).  d_adv_stats (optional, device double[3]) receives
 * {count, sum, sum of squares} BEFORE normalisation so multi-GPU callers can all-reduce them and
 * call ppo_normalize_advantages themselves (normalize = 0). */
int ppo_gae_advantages(const float *d_rewards, const float *d_values, const uint8_t *d_dones,
                       const float *d_last_value, int32_t T, int32_t B, float gamma, float lam,
                       float *d_returns, float *d_advantages, int32_t normalize, double *d_adv_stats,
                       int32_t device, void *stream);
int ppo_normalize_advantages(float *d_advantages, int64_t n, const double *d_adv_stats, int32_t device,
                             void *stream);

/* ---- PPO update: fused attention over the 5-token window (agents/ppo.py:126 -> the self-attention of
 * nn.TransformerEncoderLayer, networks/transformer_net.py:34-43,63), forward and backward in fp32 ---------------------
 * d_q: [n, num_queries, 128] rows of stride q_stride floats; d_k, d_v: [n, 5, 128] rows of stride kv_stride (views into a
 * packed in_proj output are fine); d_pad [n,5]: 1 = the key is a padding row.  num_queries = 5 (every token attends) or 1
 * (only the newest token, the last encoder layer).  d_out / d_grad_out: dense [n, num_queries, 128].  The gradients are
 * written with the strides of their inputs (d_grad_k / d_grad_v: every row; d_grad_q: every query row). */
int ppo_attn5_forward(const float *d_q, int64_t q_stride, const float *d_k, const float *d_v, int64_t kv_stride,
                      const uint8_t *d_pad, int64_t n, int32_t num_queries, float *d_out, int32_t device, void *stream);
int ppo_attn5_backward(const float *d_q, int64_t q_stride, const float *d_k, const float *d_v, int64_t kv_stride,
                       const uint8_t *d_pad, int64_t n, int32_t num_queries, const float *d_grad_out, float *d_grad_q,
                       float *d_grad_k, float *d_grad_v, int32_t device, void *stream);

/* ---- PPO update: the post-backward chain of one minibatch step on FLAT fp32 buffers -------------------------------
 * Replaces torch.nn.utils.clip_grad_norm_(policy.parameters(), GRAD_NORM_CLIP) + optimizer.step() of
 * agents/ppo.py:160-162 (Adam with the four parameter groups of agents/ppo.py:17-22) - and the division by the
 * world size that follows the NCCL gradient sum - with two launches:
 *   g = d_grad * grad_scale;  g *= min(1, max_norm / (||g||_2 + 1e-6));  Adam(beta1, beta2, eps, lr of the segment).
 * d_params / d_grad / d_exp_avg / d_exp_avg_sq: [n] fp32, same element order (d_grad is left holding the clipped g).
 * seg_end / seg_lr (HOST arrays, num_segments <= PPO_OPTIM_MAX_GROUPS): exclusive end offsets (ascending, last == n)
 * and learning rates of the contiguous parameter ranges.  d_step: device int64 Adam step counter (incremented by the
 * call, so CUDA-graph replays advance it).  d_partials: device double[ppo_optim_partials()] scratch.  d_norm_out
 * (optional): the pre-clip gradient norm.  The norm is reduced in a fixed order: bit-identical on every rank.
 * max_norm <= 0 disables clipping. */
#define PPO_OPTIM_MAX_GROUPS 8
int ppo_clip_adam_step(float *d_params, float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, int64_t n,
                       const int64_t *seg_end, const float *seg_lr, int32_t num_segments, float grad_scale,
                       float max_norm, float beta1, float beta2, float eps, int64_t *d_step, double *d_partials,
                       float *d_norm_out, int32_t device, void *stream);
int ppo_optim_partials(void);

#ifdef __cplusplus
}
#endif
#endif /* UAVENV_B200_H */
