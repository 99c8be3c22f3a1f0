// uavenv_kernels.cuh - the kernels of the batched UAV->target allocation environment (sm_100a).
//
//   step_kernel          one fused launch per env.step for all B envs (envs/uav_env.py:295-435):
//                        action/accept rule -> target products, cover flags, UAV assignment ->
//                        reward / done / info -> auto-reset of finished envs (counter RNG scene
//                        generation, uav_env.py:65-182) -> pair score + observation row of the new
//                        pointer pair (mechanics.py:167-241) -> [B,5,14] window.
//   reset_kernel         UAVEnv.reset for a masked subset (uav_env.py:42-63), one CTA per env.
//   pack_scene_kernel    scene injection (host SoA -> device records) + per-scene derived values.
//   score_matrix_kernel  p_final / p_damage [B,N,M] (main.py:38-45 over mechanics.py:167-181).
//   recompute_kernel     fresh J / N0 / sums from the per-target products, a warp per env with
//                        shuffle reductions (uav_env.py:244-293) - drift bound for the running sums.
#pragma once

#include "uavenv_device.cuh"

namespace uavk {

constexpr int kStepThreads = 128;          // envs per CTA in step_kernel (thread-per-env main phases)
constexpr int kWarpsPerCta = kStepThreads / 32;
constexpr int kResetThreads = 128;

struct StepIO {
    const int64_t *actions;  // [B]
    float *obs;              // [B,5,14]
    float *reward;           // [B]
    uint8_t *done;           // [B]
    float *J_val;            // info (nullable each)
    int32_t *num_assigned;
    int8_t *is_valid;
    float *avg_p_dmg, *avg_p_final;
    double *reward_f64;
};

// ------------------------------------------------------------------------------------------------
// block-cooperative helpers (all threads of the CTA must call)

__device__ __forceinline__ double block_sum(double v, double *s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += s_red[i];
    return t;
}

// position of element i in a sort of keys[0..n) by (key, index): a uniformly random permutation
__device__ __forceinline__ int rank_of(const uint32_t *keys, int n, int i) {
    const uint32_t ki = keys[i];
    int r = 0;
    for (int j = 0; j < n; ++j) {
        const uint32_t kj = keys[j];
        r += (kj < ki) || (kj == ki && j < i);
    }
    return r;
}

// _reset_state_only (envs/uav_env.py:175-182) for env b: every UAV available again, every lock list empty
__device__ __forceinline__ void block_clear_allocation(const Params &P, int b) {
    int32_t *asg = P.assigned + (size_t)b * P.N;
    for (int i = threadIdx.x; i < P.N; i += blockDim.x) asg[i] = -1;
    TgtRec *T = P.tgt + (size_t)b * P.M;
    for (int j = threadIdx.x; j < P.M; j += blockDim.x) {
        T[j].nh = 1.0; T[j].nh_pure = 1.0; T[j].lock_cost = 0.0; T[j].lock_cnt = 0;
    }
}

// per-scene derived values of one UAV: speed and p_pen (obstacles of env b must be visible)
__device__ __forceinline__ void finish_uav(const Params &P, int b, UavRec &u) {
    u.speed = sqrt(u.vx * u.vx + u.vy * u.vy);
    u.p_pen = penetration_prob(P, b, u.x, u.y, u.vx, u.vy, u.speed);
}

// _generate_scene (envs/uav_env.py:65-173) for env b with the counter RNG; also clears the allocation.
// s_keys: >= max(N,M) uint32, s_vals: >= M doubles, s_red: >= 32 doubles.
__device__ void block_generate_scene(const Params &P, int b, uint32_t scene, uint32_t *s_keys, double *s_vals,
                                     double *s_red) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t k0 = P.seed_lo, k1 = P.seed_hi, env = P.env_id_base + (uint32_t)b;
    const int N = P.N, M = P.M;
    // obstacles first: the per-UAV penetration probability needs them   (uav_env.py:146-170)
    NfzRec *Z = P.nfz + (size_t)b * P.K1;
    for (int i = tid; i < P.K1; i += nt) {
        const uint4 a = philox4x32(k0, k1, i, S_NFZ_A, scene, env);
        const uint4 c = philox4x32(k0, k1, i, S_NFZ_B, scene, env);
        Z[i].radius = 5.0 + (10.0 - 5.0) * u53(a.x, a.y);
        Z[i].x = 120.0 + (140.0 - 120.0) * u53(a.z, a.w);
        Z[i].y = 0.0 + (P.map_h - 0.0) * u53(c.x, c.y);
    }
    IntRec *I = P.intc + (size_t)b * P.K2;
    for (int i = tid; i < P.K2; i += nt) {
        const uint4 a = philox4x32(k0, k1, i, S_INT_A, scene, env);
        const uint4 c = philox4x32(k0, k1, i, S_INT_B, scene, env);
        I[i].x = 140.0 + (160.0 - 140.0) * u53(a.x, a.y);
        I[i].y = 0.0 + (P.map_h - 0.0) * u53(a.z, a.w);
        const double sp = 0.30 + (0.32 - 0.30) * u53(c.x, c.y);
        const double ang = 0.0 + (2.0 * 3.141592653589793 - 0.0) * u53(c.z, c.w);
        I[i].vx = cos(ang) * sp;
        I[i].vy = sin(ang) * sp;
    }
    // 1. UAV types: N//4 of type 2, uniformly permuted   (uav_env.py:81-84)
    for (int i = tid; i < N; i += nt) s_keys[i] = philox4x32(k0, k1, i, S_UAV_TYPE, scene, env).x;
    __syncthreads();  // keys + obstacles visible
    const int num_type1 = N - N / 4;
    UavRec *U = P.uav + (size_t)b * N;
    int32_t *asg = P.assigned + (size_t)b * N;
    int32_t *typ = P.uav_type + (size_t)b * N;
    double cost_part = 0.0;
    for (int i = tid; i < N; i += nt) {
        const int type = rank_of(s_keys, N, i) >= num_type1 ? 2 : 1;
        const uint4 a = philox4x32(k0, k1, i, S_UAV_POS, scene, env);
        const uint4 d = philox4x32(k0, k1, i, S_UAV_DYN, scene, env);
        UavRec u;
        u.x = P.uav_x_lo + (P.uav_x_hi - P.uav_x_lo) * u53(a.x, a.y);            // :88
        u.y = 0.0 + (P.map_h - 0.0) * u53(a.z, a.w);                              // :89
        double base_speed, base_load;
        if (type == 1) { base_speed = 0.35 + (0.50 - 0.35) * u53(d.x, d.y); u.cost = 1.0; base_load = 0.95; }
        else { base_speed = 0.75 + (0.90 - 0.75) * u53(d.x, d.y); u.cost = 1.25; base_load = 1.0; }  // :93-102
        const double real_speed = base_speed * P.weather_speed;                   // :106
        u.load = base_load * P.weather_load;                                      // :107
        const double ang = (-15.0 + (15.0 - (-15.0)) * u53(d.z, d.w)) * (3.141592653589793 / 180.0);  // :110
        u.vx = cos(ang) * real_speed;
        u.vy = sin(ang) * real_speed;                                             // :111
        finish_uav(P, b, u);
        U[i] = u;
        asg[i] = -1;
        typ[i] = type;
        cost_part += u.cost;
    }
    const double total_cost = block_sum(cost_part, s_red);  // (contains __syncthreads: s_keys reusable)
    // 2. target values   (uav_env.py:121-129)
    const int n1 = M / 2, n_remain = M - n1 - 1;
    int n2 = 0;
    if (n_remain >= 1) n2 = 1 + (int)(((uint64_t)philox4x32(k0, k1, 0u, S_N2, scene, env).x * (uint64_t)n_remain) >> 32);
    for (int i = tid; i < M; i += nt) s_keys[i] = philox4x32(k0, k1, i, S_TGT_VAL, scene, env).x;
    __syncthreads();
    double val_part = 0.0;
    for (int i = tid; i < M; i += nt) {
        const int q = rank_of(s_keys, M, i);
        const double v = q < n1 ? 4.0 : (q < n1 + n2 ? 6.0 : (q < n1 + n_remain ? 8.0 : 16.0));
        s_vals[i] = v;
        val_part += v;
    }
    const double total_val = block_sum(val_part, s_red);
    // 3. list permutation (uav_env.py:173) + kinematics: target id i lands at list position rank_i
    for (int i = tid; i < M; i += nt) s_keys[i] = philox4x32(k0, k1, i, S_TGT_LIST, scene, env).x;
    __syncthreads();
    TgtRec *T = P.tgt + (size_t)b * M;
    double2 *TV = P.tgt_vel + (size_t)b * M;
    for (int i = tid; i < M; i += nt) {
        const int pos = rank_of(s_keys, M, i);
        const uint4 a = philox4x32(k0, k1, i, S_TGT_POS, scene, env);
        const uint4 v = philox4x32(k0, k1, i, S_TGT_VEL, scene, env);
        TgtRec t;
        t.x = P.tgt_x_lo + (P.tgt_x_hi - P.tgt_x_lo) * u53(a.x, a.y);            // :134
        t.y = 0.0 + (P.map_h - 0.0) * u53(a.z, a.w);                              // :135
        const double vx = (u53(v.x, v.y) - 0.5) * 0.03, vy = (u53(v.z, v.w) - 0.5) * 0.03;  // :139
        t.speed = sqrt(vx * vx + vy * vy);
        t.value = s_vals[i];
        t.nh = 1.0; t.nh_pure = 1.0; t.lock_cost = 0.0; t.lock_cnt = 0; t.id = i;
        T[pos] = t;
        TV[pos] = make_double2(vx, vy);
    }
    if (tid == 0) { P.hd.total_cost[b] = total_cost; P.hd.total_val[b] = total_val; }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// thread-level pieces

// calc_advantage of pointer pair (k,m) + its observation row (uav_env.py:184-242 -> mechanics.py:185-241)
__device__ __forceinline__ void eval_pointer_pair(const Params &P, int b, int k, int m, double cost_sum,
                                                  double covered_val, double total_cost, double total_val,
                                                  double &pf, double &pd, float *row) {
    const UavRec u = P.uav[(size_t)b * P.N + k];
    const TgtRec t = P.tgt[(size_t)b * P.M + m];
    pd = damage_prob(P, u.x, u.y, u.vx, u.vy, u.speed, u.load, t.x, t.y, t.speed);
    pf = pd * u.p_pen;                                                  // mechanics.py:179
    const double chi_c = cost_sum / (total_cost + 1e-6);                // uav_env.py:195-196
    const double chi_v = covered_val / (total_val + 1e-6);              // :198-200
    const double chi_mc = t.lock_cost / (total_cost + 1e-6);            // :202-206
    const double P_prev = 1.0 - t.nh, P_pure = 1.0 - t.nh_pure;         // :226-227
    state_vector(u.cost, t.value, chi_c, chi_v, chi_mc, pf, pd, P_prev, P_prev * t.value, P_pure, row);
}

// ------------------------------------------------------------------------------------------------
// The fused step.  Thread t of a CTA owns env b = blockIdx.x*128 + t for the O(1) state machine;
// finished envs are restarted by the whole CTA; the [5,14] windows leave through per-warp shared
// memory tiles so the 280 B/env rows are written with coalesced 16 B stores.

__global__ void __launch_bounds__(kStepThreads) step_kernel(const __grid_constant__ Params P, const StepIO io) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ __align__(16) float s_tile[kWarpsPerCta][32 * kObsFloats];
    __shared__ int32_t s_done_env[kStepThreads];
    __shared__ int32_t s_done_cnt;
    double *s_red = reinterpret_cast<double *>(s_dyn);                 // 32 doubles
    double *s_vals = s_red + 32;                                       // M doubles
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_vals + P.M);     // max(N,M) uint32

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x * kStepThreads + tid;
    const bool live = b < P.B;
    const int N = P.N, M = P.M;
    const uint32_t head_new = (P.step_ctr[0] + 1u) % (uint32_t)kSeqLen;  // ring slot of this step's row
    if (tid == 0) s_done_cnt = 0;
    __syncthreads();

    // ---- phase A: action -> accept rule -> state update -> reward / done / info --------------------
    int k = 0, m = 0, nA = 0, n0 = 0, age = 0;
    double rev = 0, cost_sum = 0, covered_val = 0, sum_pd = 0, sum_pf = 0, total_val = 0, total_cost = 0;
    bool was_finished = false, done = false, restarted = false;
    if (live) {
        k = P.hd.k[b]; m = P.hd.m[b]; nA = P.hd.n_assigned[b]; n0 = P.hd.n_covered[b]; age = P.hd.age[b];
        rev = P.hd.rev[b]; cost_sum = P.hd.cost_sum[b]; covered_val = P.hd.covered_val[b];
        sum_pd = P.hd.sum_pd[b]; sum_pf = P.hd.sum_pf[b];
        total_val = P.hd.total_val[b]; total_cost = P.hd.total_cost[b];
        was_finished = !P.auto_reset && P.hd.finished[b];
        const int64_t action = io.actions[b];
        double reward = 0.0;
        if (!was_finished) {
            const double prev_r = paper_reward(rev - (P.omega * cost_sum), n0, M);      // uav_env.py:301
            double cur_r = prev_r;
            bool advance_uav = false;
            if (action == 1) {                                                           // :306
                const double pf = P.hd.cur_pf[b], pd = P.hd.cur_pd[b];
                TgtRec *tp = P.tgt + (size_t)b * M + m;
                const double value = tp->value, nh = tp->nh, nh_pure = tp->nh_pure, lock_cost = tp->lock_cost;
                const int lock_cnt = tp->lock_cnt;
                const double ucost = P.uav[(size_t)b * N + k].cost;
                // tentative X' (:308-310): only target m's product, the cost sum and N0 change
                const double nh2 = nh * (1.0 - pf);
                const double rev2 = rev + ((1.0 - nh2) - (1.0 - nh)) * value;
                const double cost2 = cost_sum + ucost;
                const int n02 = n0 + (lock_cnt == 0);
                const double new_r = paper_reward(rev2 - (P.omega * cost2), n02, M);    // :313
                if (new_r >= prev_r) {                                                   // :317 (Eq.21)
                    reward = new_r - prev_r;                                             // :321
                    cur_r = new_r;
                    tp->nh = nh2; tp->nh_pure = nh_pure * (1.0 - pd);
                    tp->lock_cost = lock_cost + ucost; tp->lock_cnt = lock_cnt + 1;
                    P.assigned[(size_t)b * N + k] = tp->id;                              // :308
                    rev = rev2; cost_sum = cost2;
                    if (lock_cnt == 0) { covered_val += value; n0 = n02; }
                    sum_pd += pd; sum_pf += pf; nA += 1;
                    advance_uav = true;                                                  // :324-325
                }
            }
            if (advance_uav) { k += 1; m = 0; }
            else { m += 1; if (m >= M) { k += 1; m = 0; } }                              // :336-342, :347-352
            done = k >= N;                                                               // :355
            if (done) reward += cur_r;                                                   // :361-363
            if (io.is_valid) io.is_valid[b] = (action == 1) ? (reward != 0.0 ? 1 : 0) : -1;  // :429
        } else {
            done = true;  // a finished env without auto-reset is inert
            if (io.is_valid) io.is_valid[b] = -1;
        }
        io.reward[b] = (float)reward;
        io.done[b] = done ? 1 : 0;
        if (io.reward_f64) io.reward_f64[b] = reward;
        if (io.J_val) io.J_val[b] = (float)(rev - (P.omega * cost_sum));                 // :427
        if (io.num_assigned) io.num_assigned[b] = n0;                                    // :428
        if (io.avg_p_dmg) io.avg_p_dmg[b] = nA > 0 ? (float)(sum_pd / nA) : 0.0f;        // :409
        if (io.avg_p_final) io.avg_p_final[b] = nA > 0 ? (float)(sum_pf / nA) : 0.0f;    // :416
        if (done && !was_finished) {
            if (P.auto_reset) {
                restarted = true;
                s_done_env[atomicAdd(&s_done_cnt, 1)] = tid;
            } else {
                P.hd.finished[b] = 1;
            }
        }
    }
    __syncthreads();

    // ---- phase B: the CTA restarts its finished envs (main_train.py:79 schedule) -------------------
    const int ndone = s_done_cnt;
    for (int q = 0; q < ndone; ++q) {
        const int eb = blockIdx.x * kStepThreads + s_done_env[q];
        const int episode = P.hd.episode[eb] + 1;
        const bool full = P.reset_episodes > 0 && (episode % P.reset_episodes) == 0;
        if (full) {
            const int scene = P.hd.scene_idx[eb];
            block_generate_scene(P, eb, (uint32_t)scene, s_keys, s_vals, s_red);
            if (tid == 0) P.hd.scene_idx[eb] = scene + 1;
        } else {
            block_clear_allocation(P, eb);
        }
        __syncthreads();
        if (tid == 0) P.hd.episode[eb] = episode;
    }
    if (ndone > 0) __syncthreads();

    // ---- phase C: pair score + observation row of the new pointer pair, window out -----------------
    float *tile = s_tile[warp] + lane * kObsFloats;
    if (live) {
        const bool inert = done && !restarted;  // finished, no auto-reset: zero window
        if (restarted) {
            k = 0; m = 0; nA = 0; n0 = 0; age = 0;
            rev = 0.0; cost_sum = 0.0; covered_val = 0.0; sum_pd = 0.0; sum_pf = 0.0;
            total_val = P.hd.total_val[b]; total_cost = P.hd.total_cost[b];
        }
        const int nprev = inert ? 0 : (age < kSeqLen - 1 ? age : kSeqLen - 1);
        // rows of the previous steps from the ring ([slot][feature pair][B], coalesced)
#pragma unroll
        for (int a = kSeqLen - 1; a >= 1; --a) {
            float *dst = tile + (kSeqLen - 1 - a) * kStateDim;
            if (a <= nprev) {
                const uint32_t slot = (head_new + (uint32_t)(kSeqLen - a)) % (uint32_t)kSeqLen;
                const float2 *src = P.hist + (size_t)slot * (kStateDim / 2) * P.B + b;
#pragma unroll
                for (int f = 0; f < kStateDim / 2; ++f) {
                    const float2 v = __ldg(src + (size_t)f * P.B);
                    dst[2 * f] = v.x; dst[2 * f + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int f = 0; f < kStateDim; ++f) dst[f] = 0.0f;
            }
        }
        float row[kStateDim];
        if (!inert) {
            double pf, pd;
            eval_pointer_pair(P, b, k, m, cost_sum, covered_val, total_cost, total_val, pf, pd, row);
            P.hd.cur_pf[b] = pf; P.hd.cur_pd[b] = pd;
            float2 *dsth = P.hist + (size_t)head_new * (kStateDim / 2) * P.B + b;
#pragma unroll
            for (int f = 0; f < kStateDim / 2; ++f) dsth[(size_t)f * P.B] = make_float2(row[2 * f], row[2 * f + 1]);
            age = nprev + 1;
        } else {
#pragma unroll
            for (int f = 0; f < kStateDim; ++f) row[f] = 0.0f;
            age = 0;
        }
#pragma unroll
        for (int f = 0; f < kStateDim; ++f) tile[(kSeqLen - 1) * kStateDim + f] = row[f];
        if (!was_finished) {
            P.hd.k[b] = k; P.hd.m[b] = m; P.hd.n_assigned[b] = nA; P.hd.n_covered[b] = n0; P.hd.age[b] = age;
            P.hd.rev[b] = rev; P.hd.cost_sum[b] = cost_sum; P.hd.covered_val[b] = covered_val;
            P.hd.sum_pd[b] = sum_pd; P.hd.sum_pf[b] = sum_pf;
        }
    }
    __syncwarp();
    // coalesced write-out of the warp's 32 windows (contiguous 32*280 B in obs)
    {
        const int b0 = blockIdx.x * kStepThreads + warp * 32;
        const int nenv = min(32, P.B - b0);
        if (nenv > 0) {
            float *dst = io.obs + (size_t)b0 * kObsFloats;
            const float *src = s_tile[warp];
            const int nfl = nenv * kObsFloats;
            if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
                const int nv = nfl >> 2;  // kObsFloats*4 = 280 B is a multiple of 8, 32 envs of 16
                for (int i = lane; i < nv; i += 32)
                    reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(src)[i];
                for (int i = (nv << 2) + lane; i < nfl; i += 32) dst[i] = src[i];
            } else {
                for (int i = lane; i < nfl; i += 32) dst[i] = src[i];
            }
        }
    }
    // the last CTA to finish advances the ring head (kept on the device so graph replays stay correct)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&P.step_ctr[1], 1u) == gridDim.x - 1) {
            P.step_ctr[1] = 0u;
            P.step_ctr[0] = head_new;  // stored modulo kSeqLen
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// UAVEnv.reset (envs/uav_env.py:42-63) for the masked envs: one CTA per env.
// mode: 0 = state only, 1 = generate a new scene, 2 = scene already packed (load_scene)

__global__ void __launch_bounds__(kResetThreads) reset_kernel(const __grid_constant__ Params P, int mode,
                                                               const uint8_t *mask, int first_env, int count,
                                                               float *obs) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    double *s_red = reinterpret_cast<double *>(s_dyn);
    double *s_vals = s_red + 32;
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_vals + P.M);
    __shared__ float s_row[kStateDim];
    const int tid = threadIdx.x;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int b = first_env + e;
        if (mask && !mask[b]) continue;  // uniform across the CTA
        if (mode == 1) {
            const int scene = P.hd.scene_idx[b];
            __syncthreads();
            block_generate_scene(P, b, (uint32_t)scene, s_keys, s_vals, s_red);
            if (tid == 0) P.hd.scene_idx[b] = scene + 1;
        } else {
            block_clear_allocation(P, b);
        }
        __syncthreads();
        if (tid == 0) {
            const double total_cost = P.hd.total_cost[b], total_val = P.hd.total_val[b];
            double pf, pd;
            float row[kStateDim];
            eval_pointer_pair(P, b, 0, 0, 0.0, 0.0, total_cost, total_val, pf, pd, row);
            P.hd.k[b] = 0; P.hd.m[b] = 0; P.hd.n_assigned[b] = 0; P.hd.n_covered[b] = 0; P.hd.age[b] = 1;
            P.hd.rev[b] = 0.0; P.hd.cost_sum[b] = 0.0; P.hd.covered_val[b] = 0.0;
            P.hd.sum_pd[b] = 0.0; P.hd.sum_pf[b] = 0.0; P.hd.cur_pf[b] = pf; P.hd.cur_pd[b] = pd;
            P.hd.finished[b] = 0;
            P.hd.episode[b] = (mode == 0) ? P.hd.episode[b] + 1 : 1;
            const uint32_t head = P.step_ctr[0] % (uint32_t)kSeqLen;
            float2 *dsth = P.hist + (size_t)head * (kStateDim / 2) * P.B + b;
            for (int f = 0; f < kStateDim / 2; ++f) dsth[(size_t)f * P.B] = make_float2(row[2 * f], row[2 * f + 1]);
            for (int f = 0; f < kStateDim; ++f) s_row[f] = row[f];
        }
        __syncthreads();
        if (obs) {
            float *o = obs + (size_t)b * kObsFloats;
            for (int i = tid; i < kObsFloats; i += blockDim.x)
                o[i] = i < (kSeqLen - 1) * kStateDim ? 0.0f : s_row[i - (kSeqLen - 1) * kStateDim];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Scene injection: staged SoA (device copies of the host arrays, env-major) -> records.

struct SceneSoA {
    const double *uav_x, *uav_y, *uav_vx, *uav_vy, *uav_load, *uav_cost;
    const int32_t *uav_type;
    const double *tgt_x, *tgt_y, *tgt_vx, *tgt_vy, *tgt_value;
    const int32_t *tgt_id;
    const double *nfz_x, *nfz_y, *nfz_radius, *int_x, *int_y, *int_vx, *int_vy;
};

__global__ void __launch_bounds__(kResetThreads) pack_scene_kernel(const __grid_constant__ Params P,
                                                                    const SceneSoA s, int first_env, int count) {
    __shared__ double s_red[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const int b = first_env + e;
        for (int i = tid; i < P.K1; i += nt) {
            NfzRec z; z.x = s.nfz_x[(size_t)e * P.K1 + i]; z.y = s.nfz_y[(size_t)e * P.K1 + i];
            z.radius = s.nfz_radius ? s.nfz_radius[(size_t)e * P.K1 + i] : 0.0;
            P.nfz[(size_t)b * P.K1 + i] = z;
        }
        for (int i = tid; i < P.K2; i += nt) {
            IntRec r; r.x = s.int_x[(size_t)e * P.K2 + i]; r.y = s.int_y[(size_t)e * P.K2 + i];
            r.vx = s.int_vx[(size_t)e * P.K2 + i]; r.vy = s.int_vy[(size_t)e * P.K2 + i];
            P.intc[(size_t)b * P.K2 + i] = r;
        }
        __syncthreads();
        double cost_part = 0.0, val_part = 0.0;
        for (int i = tid; i < P.N; i += nt) {
            const size_t g = (size_t)e * P.N + i;
            UavRec u;
            u.x = s.uav_x[g]; u.y = s.uav_y[g]; u.vx = s.uav_vx[g]; u.vy = s.uav_vy[g];
            u.load = s.uav_load[g]; u.cost = s.uav_cost[g];
            finish_uav(P, b, u);
            P.uav[(size_t)b * P.N + i] = u;
            P.uav_type[(size_t)b * P.N + i] = s.uav_type ? s.uav_type[g] : 1;
            cost_part += u.cost;
        }
        for (int j = tid; j < P.M; j += nt) {
            const size_t g = (size_t)e * P.M + j;
            TgtRec t;
            t.x = s.tgt_x[g]; t.y = s.tgt_y[g];
            const double vx = s.tgt_vx[g], vy = s.tgt_vy[g];
            t.speed = sqrt(vx * vx + vy * vy);
            t.value = s.tgt_value[g];
            t.nh = 1.0; t.nh_pure = 1.0; t.lock_cost = 0.0; t.lock_cnt = 0; t.id = s.tgt_id[g];
            P.tgt[(size_t)b * P.M + j] = t;
            P.tgt_vel[(size_t)b * P.M + j] = make_double2(vx, vy);
            val_part += t.value;
        }
        const double total_cost = block_sum(cost_part, s_red);
        const double total_val = block_sum(val_part, s_red);
        if (tid == 0) { P.hd.total_cost[b] = total_cost; P.hd.total_val[b] = total_val; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Score matrix: one thread per (env, UAV, target) pair, target index fastest (coalesced stores).

template <typename OutT>
__global__ void __launch_bounds__(256) score_matrix_kernel(const __grid_constant__ Params P, OutT *p_final,
                                                            OutT *p_damage) {
    const size_t total = (size_t)P.B * P.N * P.M;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx % P.M);
        const size_t bk = idx / P.M;  // b*N + k
        const size_t b = bk / P.N;
        const UavRec u = P.uav[bk];
        const TgtRec *t = P.tgt + b * P.M + m;
        const double pd = damage_prob(P, u.x, u.y, u.vx, u.vy, u.speed, u.load, t->x, t->y, t->speed);
        if (p_damage) p_damage[idx] = (OutT)pd;
        if (p_final) p_final[idx] = (OutT)(pd * u.p_pen);
    }
}

// ------------------------------------------------------------------------------------------------
// Fresh objective from the per-target products, a warp per env (uav_env.py:244-293): shuffle
// reductions over targets (revenue, covered value, N0) and UAVs (cost of assigned).  With fix != 0
// the running aggregates are re-anchored to the fresh values.

__global__ void __launch_bounds__(256) recompute_kernel(const __grid_constant__ Params P, int fix,
                                                         double *max_abs_diff) {
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double worst = 0.0;
    for (int b = wglobal; b < P.B; b += nwarps) {
        double rev = 0.0, cval = 0.0, cost = 0.0;
        int n0 = 0;
        const TgtRec *T = P.tgt + (size_t)b * P.M;
        for (int j = lane; j < P.M; j += 32) {
            rev += (1.0 - T[j].nh) * T[j].value;
            if (T[j].lock_cnt > 0) { cval += T[j].value; n0 += 1; }
        }
        const UavRec *U = P.uav + (size_t)b * P.N;
        const int32_t *asg = P.assigned + (size_t)b * P.N;
        for (int i = lane; i < P.N; i += 32) if (asg[i] >= 0) cost += U[i].cost;
        for (int o = 16; o > 0; o >>= 1) {
            rev += __shfl_xor_sync(0xffffffffu, rev, o);
            cval += __shfl_xor_sync(0xffffffffu, cval, o);
            cost += __shfl_xor_sync(0xffffffffu, cost, o);
            n0 += __shfl_xor_sync(0xffffffffu, n0, o);
        }
        if (lane == 0) {
            const double J_fresh = rev - (P.omega * cost);
            const double J_run = P.hd.rev[b] - (P.omega * P.hd.cost_sum[b]);
            worst = fmax(worst, fabs(J_fresh - J_run));
            if (n0 != P.hd.n_covered[b]) worst = fmax(worst, 1e30);  // integer state must agree exactly
            if (fix) { P.hd.rev[b] = rev; P.hd.cost_sum[b] = cost; P.hd.covered_val[b] = cval; }
        }
    }
    if (lane == 0 && max_abs_diff && worst > 0.0) {
        // doubles >= 0 order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(max_abs_diff), (unsigned long long)__double_as_longlong(worst));
    }
}

// Bernoulli(1/2) action stream keyed (seed, step, global env id)
__global__ void random_actions_kernel(const __grid_constant__ Params P, uint32_t s0, uint32_t s1, uint64_t step,
                                      int64_t *actions) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < P.B) {
        const uint4 r = philox4x32(s0, s1, P.env_id_base + (uint32_t)b, (uint32_t)step, (uint32_t)(step >> 32),
                                   0x00AC7101u);
        actions[b] = (int64_t)(r.x >> 31);
    }
}

}  // namespace uavk
