/*
 * uavenv_oracle.c - CPU oracle (plain C, fp64) for the UAV->target allocation environment.
 *
 * TEST INFRASTRUCTURE ONLY - see uavenv_oracle.h.  This file restates the reference's algorithm
 * the way the reference runs it: array-of-records entities, per-target lock lists, and every
 * objective / diagnostic recomputed from scratch by calling the pair score again (the reference
 * makes ~5A+2 calc_advantage calls per step).  It is deliberately NOT the
 * incremental formulation the CUDA path uses, so that agreement between the two is evidence.
 *
 * Citations are file:line under /root/reference.
 */
#include "uavenv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {            /* envs/entities.py:13-36 (live fields only) */
    double x, y, vx, vy;
    double load, cost;
    int32_t type;
    int32_t assigned_target_id; /* -1 = none */
    int32_t available;
} o_uav;

typedef struct {            /* envs/entities.py:39-49 + velocity attached at uav_env.py:141 */
    double x, y, vx, vy, value;
    int32_t id;
    int32_t n_locked;
    int32_t *locked;        /* UAV ids in lock order */
} o_target;

typedef struct { double x, y, radius; } o_nfz;          /* envs/entities.py:52-55 */
typedef struct { double x, y, vx, vy, radius; } o_int;  /* envs/entities.py:58-61 */

struct orc_env {
    orc_cfg c;
    int32_t N, M, K1, K2;
    o_uav *uavs;
    o_target *targets;
    o_nfz *nfz;
    o_int *inter;
    int32_t *lock_pool;
    int32_t uav_idx, target_idx;           /* uav_env.py:33-34 */
    float window[ORC_SEQ_LEN][ORC_STATE_DIM]; /* deque(maxlen=5), oldest first: uav_env.py:37 */
    double total_swarm_cost;                /* uav_env.py:40 */
};

void orc_default_cfg(orc_cfg *c) { /* configs/config.py:7-58 */
    c->num_uavs = 30; c->num_targets = 10; c->num_nfz = 1; c->num_interceptors = 1;
    c->zeta_d = 150.0; c->k = 1.2; c->c1 = 0.75; c->c2 = 0.25; c->c3 = 0.75; c->c4 = 0.25;
    c->omega = 0.0; c->weather_speed = 1.0; c->weather_load = 1.0;
    c->map_w = 180.0; c->map_h = 160.0;
    c->uav_x_lo = 60.0; c->uav_x_hi = 90.0; c->tgt_x_lo = 160.0; c->tgt_x_hi = 180.0;
    c->intercept_rad = 2.0;
}

orc_env *orc_create(const orc_cfg *c) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->c = *c;
    e->N = c->num_uavs; e->M = c->num_targets; e->K1 = c->num_nfz; e->K2 = c->num_interceptors;
    e->uavs = (o_uav *)calloc((size_t)(e->N > 0 ? e->N : 1), sizeof(o_uav));
    e->targets = (o_target *)calloc((size_t)(e->M > 0 ? e->M : 1), sizeof(o_target));
    e->nfz = (o_nfz *)calloc((size_t)(e->K1 > 0 ? e->K1 : 1), sizeof(o_nfz));
    e->inter = (o_int *)calloc((size_t)(e->K2 > 0 ? e->K2 : 1), sizeof(o_int));
    e->lock_pool = (int32_t *)calloc((size_t)e->M * (size_t)e->N + 1, sizeof(int32_t));
    for (int j = 0; j < e->M; ++j) e->targets[j].locked = e->lock_pool + (size_t)j * e->N;
    return e;
}

void orc_destroy(orc_env *e) {
    if (!e) return;
    free(e->uavs); free(e->targets); free(e->nfz); free(e->inter); free(e->lock_pool); free(e);
}

/* ------------------------------------------------------------------ mechanics */

static double clip01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); } /* np.clip(.,0,1) */

/* envs/mechanics.py:6-7 get_distance == np.linalg.norm of a 2-vector */
static double norm2(double a, double b) { return sqrt(a * a + b * b); }

/* envs/mechanics.py:11-57 calc_angle_score (Eq.1) */
double orc_angle_score(double ux, double uy, double vx, double vy, double tx, double ty) {
    double dx = tx - ux, dy = ty - uy;          /* :19 */
    double dist = norm2(dx, dy);                /* :20 */
    if (dist < 1e-6) return 1.0;                /* :23-24 */
    double nx = dx / dist, ny = dy / dist;      /* :28 */
    double speed = norm2(vx, vy);               /* :31 */
    double wx, wy;
    if (speed < 1e-6) { wx = 1.0; wy = 0.0; }   /* :32-34 */
    else { wx = vx / speed; wy = vy / speed; }  /* :36 */
    double cos_theta = nx * wx + ny * wy;       /* :39 */
    if (cos_theta < -1.0) cos_theta = -1.0;
    if (cos_theta > 1.0) cos_theta = 1.0;
    double sigma = acos(cos_theta);             /* :40 */
    double b_val = 0.002 * dist;                /* :44 */
    if (b_val < 1e-6) b_val = 1e-6;             /* :51 */
    double q = sigma / (b_val * M_PI);
    double exponent = q * q;                    /* :55 */
    return exp(-exponent);                      /* :56 */
}

/* envs/mechanics.py:61-68 calc_speed_score (Eq.2) */
double orc_speed_score(const orc_cfg *c, double uav_speed, double target_speed) {
    if (uav_speed < 1e-6) return 0.0;
    return clip01(1.0 - (c->k * target_speed / uav_speed));
}

/* envs/mechanics.py:72-89 calc_dist_score (Eq.3); D_mid = 0 on both branches */
double orc_dist_score(const orc_cfg *c, double dist, int is_obstacle) {
    double zeta = is_obstacle ? 10.0 : c->zeta_d;
    double q = (dist - 0.0) / zeta;
    return exp(-(q * q));
}

/* envs/mechanics.py:93-114 calc_damage_prob (Eq.4) */
double orc_damage_prob(const orc_cfg *c, double ux, double uy, double vx, double vy, double load,
                       double tx, double ty, double tvx, double tvy) {
    double dist = norm2(ux - tx, uy - ty);      /* :98 */
    double uav_speed = norm2(vx, vy);           /* :99 */
    double target_speed = norm2(tvx, tvy);      /* :100 */
    double E_angle = orc_angle_score(ux, uy, vx, vy, tx, ty);
    double E_dist = orc_dist_score(c, dist, 0);
    double E_speed = orc_speed_score(c, uav_speed, target_speed);
    double term = c->c1 * E_dist + c->c2 * E_speed; /* :111 */
    double p_hat = E_angle * term * load;           /* :112 */
    return clip01(p_hat);
}

/* envs/mechanics.py:118-163 calc_penetration_prob (Eq.5-6); the target argument is unused there */
static double penetration_prob(const orc_env *e, const o_uav *u) {
    const orc_cfg *c = &e->c;
    double p = 1.0;
    double uav_speed = norm2(u->vx, u->vy);     /* :127 */
    for (int i = 0; i < e->K1; ++i) {           /* :130-141 */
        const o_nfz *z = &e->nfz[i];
        double E_angle = orc_angle_score(u->x, u->y, u->vx, u->vy, z->x, z->y);
        double dist = norm2(u->x - z->x, u->y - z->y);
        double E_dist = orc_dist_score(c, dist, 1);
        double p_nfz = (1.0 - E_angle) * (1.0 - E_dist);
        p *= clip01(p_nfz);
    }
    for (int i = 0; i < e->K2; ++i) {           /* :144-161 */
        const o_int *it = &e->inter[i];
        double E_angle = orc_angle_score(u->x, u->y, u->vx, u->vy, it->x, it->y);
        double dist = norm2(u->x - it->x, u->y - it->y);
        double E_dist = orc_dist_score(c, dist, 1);
        double inter_speed = norm2(it->vx, it->vy);
        double E_speed = orc_speed_score(c, uav_speed, inter_speed);
        double term = c->c3 * (1.0 - E_dist) + c->c4 * E_speed;
        double p_int = (1.0 - E_angle) * term;
        p *= clip01(p_int);
    }
    return p;
}

/* envs/mechanics.py:167-181 calc_advantage -> (p_final, p_damage) */
static void advantage(const orc_env *e, const o_uav *u, const o_target *t, double *p_final, double *p_damage) {
    double pd = orc_damage_prob(&e->c, u->x, u->y, u->vx, u->vy, u->load, t->x, t->y, t->vx, t->vy);
    double pp = penetration_prob(e, u);
    *p_final = pd * pp;
    *p_damage = pd;
}

/* envs/mechanics.py:185-241 get_state_vector, on explicit inputs */
void orc_state_vector_raw(double cost, double value, double chi_c, double chi_v, double chi_mc,
                          double p_km, double p_km_dmg, double prev_joint_p, double prev_revenue,
                          double prev_joint_p_pure, int available, float *out) {
    double hat_p_m = 1.0 - (1.0 - prev_joint_p) * (1.0 - p_km);                /* :196 */
    double hat_p_m_pure = 1.0 - (1.0 - prev_joint_p_pure) * (1.0 - p_km_dmg);  /* :199 */
    double hat_G_m = hat_p_m * value;                                          /* :201 */
    double delta_p_km = p_km_dmg - p_km;                                       /* :204 */
    double delta_p_m = hat_p_m_pure - hat_p_m;                                 /* :205 */
    double delta_G_m = (hat_p_m_pure * value) - hat_G_m;                       /* :206 */
    out[0] = (float)cost;  out[1] = (float)value; out[2] = (float)chi_c; out[3] = (float)chi_v;
    out[4] = (float)chi_mc; out[5] = (float)p_km; out[6] = (float)prev_joint_p; out[7] = (float)hat_p_m;
    out[8] = (float)prev_revenue; out[9] = (float)hat_G_m; out[10] = (float)delta_p_km;
    out[11] = (float)delta_p_m; out[12] = (float)delta_G_m; out[13] = available ? 1.0f : 0.0f;
    out[0] /= 2.0f; out[1] /= 16.0f; out[8] /= 16.0f; out[9] /= 16.0f; out[12] /= 16.0f; /* :235-239 */
}

/* ------------------------------------------------------------------ env */

static o_uav *uav_by_id(orc_env *e, int32_t uid) {
    /* next(u for u in self.uavs if u.id == uid): UAV ids equal list positions (uav_env.py:86,114) */
    return (uid >= 0 && uid < e->N) ? &e->uavs[uid] : NULL;
}

static void window_push(orc_env *e, const float *row) { /* deque.append with maxlen 5 */
    memmove(e->window[0], e->window[1], sizeof(float) * ORC_STATE_DIM * (ORC_SEQ_LEN - 1));
    memcpy(e->window[ORC_SEQ_LEN - 1], row, sizeof(float) * ORC_STATE_DIM);
}

/* envs/uav_env.py:184-242 _get_obs; returns rows written (5, or 1 zero row when finished) */
static int32_t get_obs(orc_env *e, float *obs) {
    if (e->uav_idx >= e->N) {                    /* :188-189 */
        memset(obs, 0, sizeof(float) * ORC_STATE_DIM);
        return 1;
    }
    o_uav *cu = &e->uavs[e->uav_idx];
    o_target *ct = &e->targets[e->target_idx];
    double assigned_cost = 0.0;                  /* :195 */
    for (int i = 0; i < e->N; ++i) if (!e->uavs[i].available) assigned_cost += e->uavs[i].cost;
    double chi_c = assigned_cost / (e->total_swarm_cost + 1e-6);
    double total_val = 0.0, covered_val = 0.0;   /* :198-200 */
    for (int j = 0; j < e->M; ++j) total_val += e->targets[j].value;
    for (int j = 0; j < e->M; ++j) if (e->targets[j].n_locked > 0) covered_val += e->targets[j].value;
    double chi_v = covered_val / (total_val + 1e-6);
    double tgt_cost = 0.0;                       /* :202-206 */
    for (int q = 0; q < ct->n_locked; ++q) { o_uav *u = uav_by_id(e, ct->locked[q]); if (u) tgt_cost += u->cost; }
    double chi_mc = tgt_cost / (e->total_swarm_cost + 1e-6);
    double not_hit = 1.0, not_hit_pure = 1.0;    /* :215-224 */
    for (int q = 0; q < ct->n_locked; ++q) {
        o_uav *u = uav_by_id(e, ct->locked[q]);
        if (u) { double pa, pp; advantage(e, u, ct, &pa, &pp); not_hit *= (1.0 - pa); not_hit_pure *= (1.0 - pp); }
    }
    double prev_joint_p = 1.0 - not_hit;
    double prev_joint_p_pure = 1.0 - not_hit_pure;
    double prev_revenue = prev_joint_p * ct->value;
    double p_km, p_km_dmg;                       /* mechanics.py:192 */
    advantage(e, cu, ct, &p_km, &p_km_dmg);
    float row[ORC_STATE_DIM];
    orc_state_vector_raw(cu->cost, ct->value, chi_c, chi_v, chi_mc, p_km, p_km_dmg, prev_joint_p,
                         prev_revenue, prev_joint_p_pure, cu->available, row);
    window_push(e, row);                         /* :241 */
    memcpy(obs, e->window, sizeof(e->window));   /* :242 */
    return ORC_SEQ_LEN;
}

/* envs/uav_env.py:244-269 _calc_J_X */
static double calc_J(orc_env *e) {
    double total_revenue = 0.0, total_cost = 0.0;
    for (int j = 0; j < e->M; ++j) {
        o_target *t = &e->targets[j];
        double not_hit = 1.0;
        for (int q = 0; q < t->n_locked; ++q) {
            o_uav *u = uav_by_id(e, t->locked[q]);
            if (u) { double pa, pd; advantage(e, u, t, &pa, &pd); not_hit *= (1.0 - pa); total_cost += u->cost; }
        }
        double joint_p = 1.0 - not_hit;
        total_revenue += joint_p * t->value;
    }
    return total_revenue - (e->c.omega * total_cost);
}
double orc_calc_J(const orc_env *e) { return calc_J((orc_env *)e); }

/* envs/uav_env.py:271-293 _calculate_paper_reward (Eq.19) */
static double paper_reward(orc_env *e) {
    double J = calc_J(e);
    int N0 = 0;
    for (int j = 0; j < e->M; ++j) if (e->targets[j].n_locked > 0) ++N0;
    int M = e->M;
    if (N0 == M) return 2.0 * J;
    return J * ((double)N0 / (double)M);
}
double orc_paper_reward(const orc_env *e) { return paper_reward((orc_env *)e); }

static void reset_state_only(orc_env *e) {  /* envs/uav_env.py:175-182 */
    for (int i = 0; i < e->N; ++i) { e->uavs[i].available = 1; e->uavs[i].assigned_target_id = -1; }
    for (int j = 0; j < e->M; ++j) e->targets[j].n_locked = 0;
}

/* envs/uav_env.py:42-63 reset: the caller has either loaded/generated a scene (full reset) or not */
void orc_reset(orc_env *e, float *obs) {
    reset_state_only(e);
    e->uav_idx = 0; e->target_idx = 0;
    memset(e->window, 0, sizeof(e->window));
    get_obs(e, obs);
}

/* envs/uav_env.py:295-435 step */
int orc_step(orc_env *e, int64_t action, float *obs, int32_t *obs_rows, double *reward_out, int32_t *done_out,
             orc_info *info) {
    if (e->uav_idx >= e->N) return -1;               /* :296 IndexError */
    o_uav *cu = &e->uavs[e->uav_idx];
    o_target *ct = &e->targets[e->target_idx];
    int done = 0;
    double prev_r = paper_reward(e);                 /* :301 */
    double reward = 0.0;
    if (action == 1) {                               /* :306 */
        cu->assigned_target_id = ct->id;
        cu->available = 0;
        ct->locked[ct->n_locked++] = (int32_t)(cu - e->uavs);
        double new_r = paper_reward(e);              /* :313 */
        if (new_r >= prev_r) {                       /* :317 */
            reward = new_r - prev_r;
            e->uav_idx += 1; e->target_idx = 0;
        } else {                                     /* :326-342 */
            cu->assigned_target_id = -1; cu->available = 1; ct->n_locked--;
            reward = 0.0;
            e->target_idx += 1;
            if (e->target_idx >= e->M) { e->uav_idx += 1; e->target_idx = 0; }
        }
    } else {                                         /* :344-352 */
        reward = 0.0;
        e->target_idx += 1;
        if (e->target_idx >= e->M) { e->uav_idx += 1; e->target_idx = 0; }
    }
    if (e->uav_idx >= e->N) done = 1;                /* :355-356 */
    if (done) reward += paper_reward(e);             /* :361-363 */
    *obs_rows = get_obs(e, obs);                     /* :367 */
    double total_dmg = 0.0, total_final = 0.0;       /* :370-408 */
    int count = 0;
    for (int j = 0; j < e->M; ++j) {
        o_target *t = &e->targets[j];
        for (int q = 0; q < t->n_locked; ++q) {
            o_uav *u = uav_by_id(e, t->locked[q]);
            if (u) { double pf, pd; advantage(e, u, t, &pf, &pd); total_dmg += pd; ++count; }
        }
    }
    for (int j = 0; j < e->M; ++j) {                 /* :403-406 second pass for p_final */
        o_target *t = &e->targets[j];
        for (int q = 0; q < t->n_locked; ++q) {
            double pf, pd; advantage(e, uav_by_id(e, t->locked[q]), t, &pf, &pd); total_final += pf;
        }
    }
    if (done) (void)calc_J(e);                       /* :414-418 (overwritten below, but evaluated) */
    info->J_val = calc_J(e);                         /* :427 */
    int n0 = 0;
    for (int j = 0; j < e->M; ++j) if (e->targets[j].n_locked > 0) ++n0;
    info->num_assigned = n0;                         /* :428 */
    info->is_valid_action = (action == 1) ? (reward != 0.0 ? 1 : 0) : -1; /* :429 */
    info->avg_p_dmg = count > 0 ? total_dmg / count : 0.0;
    info->avg_p_final = count > 0 ? total_final / count : 0.0;
    *reward_out = reward;
    *done_out = done;
    return 0;
}

void orc_score_matrix(const orc_env *e, double *p_final, double *p_damage, double *p_pen) {
    for (int i = 0; i < e->N; ++i) {                 /* main.py:38-45 */
        if (p_pen) p_pen[i] = penetration_prob(e, &e->uavs[i]);
        for (int j = 0; j < e->M; ++j)
            advantage(e, &e->uavs[i], &e->targets[j], &p_final[(size_t)i * e->M + j], &p_damage[(size_t)i * e->M + j]);
    }
}

int32_t orc_uav_idx(const orc_env *e) { return e->uav_idx; }
int32_t orc_target_idx(const orc_env *e) { return e->target_idx; }
void orc_get_assigned(const orc_env *e, int32_t *a) { for (int i = 0; i < e->N; ++i) a[i] = e->uavs[i].assigned_target_id; }
void orc_get_covered(const orc_env *e, uint8_t *c) { for (int j = 0; j < e->M; ++j) c[j] = e->targets[j].n_locked > 0; }

/* ------------------------------------------------------------------ scene I/O */

void orc_load_scene(orc_env *e, const double *uav_x, const double *uav_y, const double *uav_vx,
                    const double *uav_vy, const double *uav_load, const double *uav_cost,
                    const int32_t *uav_type, const double *tgt_x, const double *tgt_y,
                    const double *tgt_vx, const double *tgt_vy, const double *tgt_value,
                    const int32_t *tgt_id, const double *nfz_x, const double *nfz_y,
                    const double *nfz_radius, const double *int_x, const double *int_y,
                    const double *int_vx, const double *int_vy) {
    e->total_swarm_cost = 0.0;
    for (int i = 0; i < e->N; ++i) {
        o_uav *u = &e->uavs[i];
        u->x = uav_x[i]; u->y = uav_y[i]; u->vx = uav_vx[i]; u->vy = uav_vy[i];
        u->load = uav_load[i]; u->cost = uav_cost[i]; u->type = uav_type ? uav_type[i] : 1;
        u->available = 1; u->assigned_target_id = -1;
        e->total_swarm_cost += u->cost;              /* uav_env.py:118 */
    }
    for (int j = 0; j < e->M; ++j) {
        o_target *t = &e->targets[j];
        t->x = tgt_x[j]; t->y = tgt_y[j]; t->vx = tgt_vx[j]; t->vy = tgt_vy[j];
        t->value = tgt_value[j]; t->id = tgt_id[j]; t->n_locked = 0;
    }
    for (int i = 0; i < e->K1; ++i) { e->nfz[i].x = nfz_x[i]; e->nfz[i].y = nfz_y[i]; e->nfz[i].radius = nfz_radius ? nfz_radius[i] : 0.0; }
    for (int i = 0; i < e->K2; ++i) {
        e->inter[i].x = int_x[i]; e->inter[i].y = int_y[i]; e->inter[i].vx = int_vx[i]; e->inter[i].vy = int_vy[i];
        e->inter[i].radius = e->c.intercept_rad;
    }
}

void orc_export_scene(const orc_env *e, double *uav_x, double *uav_y, double *uav_vx, double *uav_vy,
                      double *uav_load, double *uav_cost, int32_t *uav_type, double *tgt_x,
                      double *tgt_y, double *tgt_vx, double *tgt_vy, double *tgt_value,
                      int32_t *tgt_id, double *nfz_x, double *nfz_y, double *nfz_radius,
                      double *int_x, double *int_y, double *int_vx, double *int_vy) {
    for (int i = 0; i < e->N; ++i) {
        const o_uav *u = &e->uavs[i];
        uav_x[i] = u->x; uav_y[i] = u->y; uav_vx[i] = u->vx; uav_vy[i] = u->vy;
        uav_load[i] = u->load; uav_cost[i] = u->cost; uav_type[i] = u->type;
    }
    for (int j = 0; j < e->M; ++j) {
        const o_target *t = &e->targets[j];
        tgt_x[j] = t->x; tgt_y[j] = t->y; tgt_vx[j] = t->vx; tgt_vy[j] = t->vy; tgt_value[j] = t->value; tgt_id[j] = t->id;
    }
    for (int i = 0; i < e->K1; ++i) { nfz_x[i] = e->nfz[i].x; nfz_y[i] = e->nfz[i].y; nfz_radius[i] = e->nfz[i].radius; }
    for (int i = 0; i < e->K2; ++i) { int_x[i] = e->inter[i].x; int_y[i] = e->inter[i].y; int_vx[i] = e->inter[i].vx; int_vy[i] = e->inter[i].vy; }
}

/* ------------------------------------------------------------------ counter-based RNG */

void orc_philox(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    /* Philox4x32-10 (Salmon et al., SC'11) */
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u53(uint32_t hi, uint32_t lo) { /* uniform in [0,1) on a 2^-53 grid */
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

/* draw streams of the scene generator (order of envs/uav_env.py:65-173) */
enum { S_UAV_TYPE = 1, S_UAV_POS = 2, S_UAV_DYN = 3, S_N2 = 4, S_TGT_VAL = 5, S_TGT_POS = 6, S_TGT_VEL = 7,
       S_NFZ_A = 8, S_NFZ_B = 9, S_INT_A = 10, S_INT_B = 11, S_TGT_LIST = 12 };

static int32_t rank_of(const uint32_t *keys, int n, int i) { /* position of i in a sort by (key, index) */
    int32_t r = 0;
    for (int j = 0; j < n; ++j) r += (keys[j] < keys[i]) || (keys[j] == keys[i] && j < i);
    return r;
}

void orc_generate_scene(orc_env *e, uint64_t seed, uint32_t env_id, uint32_t scene_idx) {
    const orc_cfg *c = &e->c;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int N = e->N, M = e->M;
    uint32_t r[4];
    int nmax = N > M ? N : M;
    uint32_t *keys = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nmax > 0 ? nmax : 1));
    /* 1. UAV types: N//4 type-2, the rest type-1, uniformly permuted (uav_env.py:81-84) */
    int num_type2 = N / 4, num_type1 = N - num_type2;
    for (int i = 0; i < N; ++i) { orc_philox(k0, k1, (uint32_t)i, S_UAV_TYPE, scene_idx, env_id, r); keys[i] = r[0]; }
    e->total_swarm_cost = 0.0;
    for (int i = 0; i < N; ++i) {
        o_uav *u = &e->uavs[i];
        int type = rank_of(keys, N, i) >= num_type1 ? 2 : 1;
        orc_philox(k0, k1, (uint32_t)i, S_UAV_POS, scene_idx, env_id, r);
        u->x = c->uav_x_lo + (c->uav_x_hi - c->uav_x_lo) * u53(r[0], r[1]);   /* :88 */
        u->y = 0.0 + (c->map_h - 0.0) * u53(r[2], r[3]);                        /* :89 */
        orc_philox(k0, k1, (uint32_t)i, S_UAV_DYN, scene_idx, env_id, r);
        double base_speed, cost, base_load;
        if (type == 1) { base_speed = 0.35 + (0.50 - 0.35) * u53(r[0], r[1]); cost = 1.0; base_load = 0.95; }  /* :93-97 */
        else { base_speed = 0.75 + (0.90 - 0.75) * u53(r[0], r[1]); cost = 1.25; base_load = 1.0; }             /* :98-102 */
        double real_speed = base_speed * c->weather_speed;                      /* :106 */
        double real_load = base_load * c->weather_load;                         /* :107 */
        double deg = -15.0 + (15.0 - (-15.0)) * u53(r[2], r[3]);                /* :110 */
        double angle = deg * (M_PI / 180.0);
        u->vx = cos(angle) * real_speed; u->vy = sin(angle) * real_speed;       /* :111 */
        u->load = real_load; u->cost = cost; u->type = type;
        u->available = 1; u->assigned_target_id = -1;
        e->total_swarm_cost += cost;                                            /* :118 */
    }
    /* 2. target values (uav_env.py:121-129) */
    int n1 = M / 2, n4 = 1, n_remain = M - n1 - n4, n2 = 0;
    if (n_remain >= 1) {
        orc_philox(k0, k1, 0u, S_N2, scene_idx, env_id, r);
        n2 = 1 + (int)(((uint64_t)r[0] * (uint64_t)n_remain) >> 32);            /* randint(1, n_remain+1) */
    }
    for (int i = 0; i < M; ++i) { orc_philox(k0, k1, (uint32_t)i, S_TGT_VAL, scene_idx, env_id, r); keys[i] = r[0]; }
    double *vals = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    for (int i = 0; i < M; ++i) {
        int q = rank_of(keys, M, i);
        vals[i] = q < n1 ? 4.0 : (q < n1 + n2 ? 6.0 : (q < n1 + n_remain ? 8.0 : 16.0));
    }
    /* 3. final list permutation (uav_env.py:173): target id i sits at list position rank_i */
    for (int i = 0; i < M; ++i) { orc_philox(k0, k1, (uint32_t)i, S_TGT_LIST, scene_idx, env_id, r); keys[i] = r[0]; }
    for (int i = 0; i < M; ++i) {
        o_target *t = &e->targets[rank_of(keys, M, i)];
        orc_philox(k0, k1, (uint32_t)i, S_TGT_POS, scene_idx, env_id, r);
        t->x = c->tgt_x_lo + (c->tgt_x_hi - c->tgt_x_lo) * u53(r[0], r[1]);   /* :134 */
        t->y = 0.0 + (c->map_h - 0.0) * u53(r[2], r[3]);                        /* :135 */
        orc_philox(k0, k1, (uint32_t)i, S_TGT_VEL, scene_idx, env_id, r);
        t->vx = (u53(r[0], r[1]) - 0.5) * 0.03; t->vy = (u53(r[2], r[3]) - 0.5) * 0.03; /* :139 */
        t->value = vals[i]; t->id = i; t->n_locked = 0;
    }
    for (int i = 0; i < e->K1; ++i) {                                           /* :146-153 */
        orc_philox(k0, k1, (uint32_t)i, S_NFZ_A, scene_idx, env_id, r);
        e->nfz[i].radius = 5.0 + (10.0 - 5.0) * u53(r[0], r[1]);
        e->nfz[i].x = 120.0 + (140.0 - 120.0) * u53(r[2], r[3]);
        orc_philox(k0, k1, (uint32_t)i, S_NFZ_B, scene_idx, env_id, r);
        e->nfz[i].y = 0.0 + (c->map_h - 0.0) * u53(r[0], r[1]);
    }
    for (int i = 0; i < e->K2; ++i) {                                           /* :157-170 */
        orc_philox(k0, k1, (uint32_t)i, S_INT_A, scene_idx, env_id, r);
        e->inter[i].x = 140.0 + (160.0 - 140.0) * u53(r[0], r[1]);
        e->inter[i].y = 0.0 + (c->map_h - 0.0) * u53(r[2], r[3]);
        orc_philox(k0, k1, (uint32_t)i, S_INT_B, scene_idx, env_id, r);
        double sp = 0.30 + (0.32 - 0.30) * u53(r[0], r[1]);
        double ang = 0.0 + (2.0 * M_PI - 0.0) * u53(r[2], r[3]);
        e->inter[i].vx = cos(ang) * sp; e->inter[i].vy = sin(ang) * sp;
        e->inter[i].radius = c->intercept_rad;
    }
    free(vals); free(keys);
}

int64_t orc_random_action(uint64_t action_seed, uint64_t step, uint32_t env_id) {
    uint32_t r[4];
    orc_philox((uint32_t)action_seed, (uint32_t)(action_seed >> 32), env_id, (uint32_t)step, (uint32_t)(step >> 32),
               0x00AC7101u, r);
    return (int64_t)(r[0] >> 31);
}

/* ------------------------------------------------------------------ CPU baseline loop */

int64_t orc_rollout_random(const orc_cfg *c, int32_t num_envs, int64_t steps, uint64_t seed,
                           uint64_t action_seed, int32_t reset_episodes, int32_t threads, double *checksum) {
    int64_t total = 0;
    double sum = 0.0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total, sum)
#endif
    for (int32_t b = 0; b < num_envs; ++b) {
        orc_env *e = orc_create(c);
        float obs[ORC_SEQ_LEN * ORC_STATE_DIM];
        uint32_t scene = 0, episode = 1;          /* main_train.py:77-79: i_episode starts at 1 */
        orc_generate_scene(e, seed, (uint32_t)b, scene);
        orc_reset(e, obs);
        for (int64_t s = 0; s < steps; ++s) {
            double reward; int32_t done, rows; orc_info info;
            orc_step(e, orc_random_action(action_seed, (uint64_t)s, (uint32_t)b), obs, &rows, &reward, &done, &info);
            sum += reward; ++total;
            if (done) {
                ++episode;
                if (reset_episodes > 0 && episode % (uint32_t)reset_episodes == 0) orc_generate_scene(e, seed, (uint32_t)b, ++scene);
                orc_reset(e, obs);
            }
        }
        orc_destroy(e);
    }
    if (checksum) *checksum = sum;
    return total;
}

/* ------------------------------------------------------------------ persistent batch (bench --impl reference) */

struct orc_batch {
    int32_t E, reset_episodes;
    uint64_t seed;
    orc_env **envs;
    uint32_t *episode, *scene;
};

orc_batch *orc_batch_create(const orc_cfg *c, int32_t num_envs, uint64_t seed, int32_t reset_episodes, int32_t threads) {
    orc_batch *B = (orc_batch *)calloc(1, sizeof(orc_batch));
    B->E = num_envs; B->seed = seed; B->reset_episodes = reset_episodes;
    B->envs = (orc_env **)calloc((size_t)num_envs, sizeof(orc_env *));
    B->episode = (uint32_t *)calloc((size_t)num_envs, sizeof(uint32_t));
    B->scene = (uint32_t *)calloc((size_t)num_envs, sizeof(uint32_t));
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static)
#endif
    for (int32_t b = 0; b < num_envs; ++b) {
        float obs[ORC_SEQ_LEN * ORC_STATE_DIM];
        B->envs[b] = orc_create(c);
        B->episode[b] = 1;
        orc_generate_scene(B->envs[b], seed, (uint32_t)b, 0);
        orc_reset(B->envs[b], obs);
    }
    (void)threads;
    return B;
}

void orc_batch_destroy(orc_batch *B) {
    if (!B) return;
    for (int32_t b = 0; b < B->E; ++b) orc_destroy(B->envs[b]);
    free(B->envs); free(B->episode); free(B->scene); free(B);
}

/* one reference-algorithm step of every env with the Bernoulli action stream; returns sum of rewards */
double orc_batch_step_random(orc_batch *B, uint64_t action_seed, uint64_t step, int32_t threads) {
    double sum = 0.0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static) reduction(+ : sum)
#endif
    for (int32_t b = 0; b < B->E; ++b) {
        float obs[ORC_SEQ_LEN * ORC_STATE_DIM];
        double reward; int32_t done, rows; orc_info info;
        orc_env *e = B->envs[b];
        orc_step(e, orc_random_action(action_seed, step, (uint32_t)b), obs, &rows, &reward, &done, &info);
        sum += reward;
        if (done) {                                   /* main_train.py:79 schedule */
            ++B->episode[b];
            if (B->reset_episodes > 0 && B->episode[b] % (uint32_t)B->reset_episodes == 0)
                orc_generate_scene(e, B->seed, (uint32_t)b, ++B->scene[b]);
            orc_reset(e, obs);
        }
    }
    (void)threads;
    return sum;
}

/* differential testing at scale: one reference-algorithm step of every env with GIVEN actions; per-env outputs.
 * The state reported (uav_idx, target_idx, and orc_batch_get_assigned) is the state AFTER the step and after the
 * main_train.py:79 restart of finished envs - what the batched GPU env holds after its step. */
void orc_batch_step_actions(orc_batch *B, const int64_t *actions, int32_t threads, double *reward, uint8_t *done,
                            int32_t *num_assigned, int32_t *is_valid, int32_t *uav_idx, int32_t *target_idx) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int32_t b = 0; b < B->E; ++b) {
        float obs[ORC_SEQ_LEN * ORC_STATE_DIM];
        int32_t dn, rows; orc_info info;
        orc_env *e = B->envs[b];
        orc_step(e, actions[b], obs, &rows, &reward[b], &dn, &info);
        done[b] = (uint8_t)dn; num_assigned[b] = info.num_assigned; is_valid[b] = info.is_valid_action;
        if (dn) {                                     /* main_train.py:79 schedule */
            ++B->episode[b];
            if (B->reset_episodes > 0 && B->episode[b] % (uint32_t)B->reset_episodes == 0)
                orc_generate_scene(e, B->seed, (uint32_t)b, ++B->scene[b]);
            orc_reset(e, obs);
        }
        uav_idx[b] = e->uav_idx; target_idx[b] = e->target_idx;
    }
    (void)threads;
}

void orc_batch_get_assigned(const orc_batch *B, int32_t *assigned /* [E][N] */) {
    for (int32_t b = 0; b < B->E; ++b) orc_get_assigned(B->envs[b], assigned + (size_t)b * B->envs[b]->N);
}

int32_t orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
