/*
 * uavenv_oracle.h - CPU oracle for the UAV->target allocation environment.
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C (fp64) restatement of the reference's
 * algorithm for the rollout path, used as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under target-allocation-ppo-transformer_b200/ links, imports or
 * executes it.
 *
 * Parity pinning: tests/test_oracle_golden.py replays every fixture under
 * tests/golden/ (produced by oracle/gen_golden.py from the UNMODIFIED
 * reference) through this oracle: integers bit-exact, fp64 values <= 1e-12 rel
 * (libm vs numpy transcendental ulps).  oracle/pin_oracle.py repeats that
 * against the live reference whenever /root/reference is present.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef UAVENV_ORACLE_H
#define UAVENV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_STATE_DIM 14 /* configs/config.py:61 */
#define ORC_SEQ_LEN 5    /* configs/config.py:62 */

/* constants read from the reference's cfg singleton (configs/config.py:7-58) */
typedef struct orc_cfg {
    int32_t num_uavs, num_targets, num_nfz, num_interceptors; /* :42-49 */
    double zeta_d, k, c1, c2, c3, c4;                          /* :7-12  */
    double omega;                                              /* :53    */
    double weather_speed, weather_load;                        /* :57-58 */
    double map_w, map_h;                                       /* :33-34 */
    double uav_x_lo, uav_x_hi, tgt_x_lo, tgt_x_hi;             /* :39-40 */
    double intercept_rad;                                      /* :50    */
} orc_cfg;

typedef struct orc_info { /* envs/uav_env.py:426-433 */
    double J_val;
    int32_t num_assigned;    /* covered-target count N0 */
    int32_t is_valid_action; /* -1 = None, 0 = False, 1 = True */
    double avg_p_dmg, avg_p_final;
} orc_info;

typedef struct orc_env orc_env;

void orc_default_cfg(orc_cfg *c);
orc_env *orc_create(const orc_cfg *c);
void orc_destroy(orc_env *e);

/* scene in LIST order (targets after the uav_env.py:173 shuffle), fp64 */
void orc_load_scene(orc_env *e, const double *uav_x, const double *uav_y, const double *uav_vx,
                    const double *uav_vy, const double *uav_load, const double *uav_cost,
                    const int32_t *uav_type, const double *tgt_x, const double *tgt_y,
                    const double *tgt_vx, const double *tgt_vy, const double *tgt_value,
                    const int32_t *tgt_id, const double *nfz_x, const double *nfz_y,
                    const double *nfz_radius, const double *int_x, const double *int_y,
                    const double *int_vx, const double *int_vy);
void orc_export_scene(const orc_env *e, double *uav_x, double *uav_y, double *uav_vx, double *uav_vy,
                      double *uav_load, double *uav_cost, int32_t *uav_type, double *tgt_x,
                      double *tgt_y, double *tgt_vx, double *tgt_vy, double *tgt_value,
                      int32_t *tgt_id, double *nfz_x, double *nfz_y, double *nfz_radius,
                      double *int_x, double *int_y, double *int_vx, double *int_vy);

/* counter-based (Philox4x32-10) scene generator: the draw list of
 * envs/uav_env.py:65-173 keyed on (seed, global env id, scene index) */
void orc_generate_scene(orc_env *e, uint64_t seed, uint32_t env_id, uint32_t scene_idx);

/* envs/uav_env.py:42-63; full_reset != 0 keeps the scene already loaded/generated */
void orc_reset(orc_env *e, float *obs /* [5*14] */);
/* envs/uav_env.py:295-435; returns 0, or -1 when stepping a finished episode (IndexError, :296).
 * obs_rows is 5 normally and 1 (a single zero row) on done (:188-189). */
int orc_step(orc_env *e, int64_t action, float *obs, int32_t *obs_rows, double *reward, int32_t *done,
             orc_info *info);

/* main.py:38-45 style full matrices; p_pen may be NULL */
void orc_score_matrix(const orc_env *e, double *p_final, double *p_damage, double *p_pen);

/* score primitives, envs/mechanics.py:11-114 */
double orc_angle_score(double ux, double uy, double vx, double vy, double tx, double ty);
double orc_speed_score(const orc_cfg *c, double uav_speed, double target_speed);
double orc_dist_score(const orc_cfg *c, double dist, int is_obstacle);
double orc_damage_prob(const orc_cfg *c, double ux, double uy, double vx, double vy, double load,
                       double tx, double ty, double tvx, double tvy);
/* envs/mechanics.py:185-241 on explicit inputs (KAT entry point) */
void orc_state_vector_raw(double cost, double value, double chi_c, double chi_v, double chi_mc,
                          double p_km, double p_km_dmg, double prev_joint_p, double prev_revenue,
                          double prev_joint_p_pure, int available, float *out14);

/* state readback */
int32_t orc_uav_idx(const orc_env *e);
int32_t orc_target_idx(const orc_env *e);
void orc_get_assigned(const orc_env *e, int32_t *assigned /* [N] target ids, -1 = none */);
void orc_get_covered(const orc_env *e, uint8_t *covered /* [M] list order */);
double orc_paper_reward(const orc_env *e); /* envs/uav_env.py:271-293 */
double orc_calc_J(const orc_env *e);       /* envs/uav_env.py:244-269 */

/* Bernoulli(1/2) action keyed (action_seed, step, global env id): SURVEY.md §8d */
int64_t orc_random_action(uint64_t action_seed, uint64_t step, uint32_t env_id);

/* raw Philox4x32-10 block, for cross-checking the device generator */
void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                uint32_t out[4]);

/* CPU baseline: run `steps` reference-algorithm steps on each of `num_envs` independent envs with
 * the Bernoulli action stream, auto-resetting with the main_train.py:79 schedule (full reset every
 * reset_episodes episodes), on `threads` OpenMP threads.  Returns transitions executed;
 * *checksum accumulates rewards so the work cannot be optimised away. */
int64_t orc_rollout_random(const orc_cfg *c, int32_t num_envs, int64_t steps, uint64_t seed,
                           uint64_t action_seed, int32_t reset_episodes, int32_t threads,
                           double *checksum);

/* persistent batch of independent envs for the timed reference arm (bench.py --impl reference) */
typedef struct orc_batch orc_batch;
orc_batch *orc_batch_create(const orc_cfg *c, int32_t num_envs, uint64_t seed, int32_t reset_episodes, int32_t threads);
void orc_batch_destroy(orc_batch *b);
double orc_batch_step_random(orc_batch *b, uint64_t action_seed, uint64_t step, int32_t threads);
void orc_batch_step_actions(orc_batch *b, const int64_t *actions, int32_t threads, double *reward, uint8_t *done,
                            int32_t *num_assigned, int32_t *is_valid, int32_t *uav_idx, int32_t *target_idx);
void orc_batch_get_assigned(const orc_batch *b, int32_t *assigned);
int32_t orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
