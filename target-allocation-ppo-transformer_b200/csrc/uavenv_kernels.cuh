// uavenv_kernels.cuh - the kernels of the batched UAV->target allocation environment (sm_100a).
//
//   step_kernel          one fused launch per env.step for all B envs (envs/uav_env.py:295-435):
//                        action/accept rule -> target products, cover flags, UAV assignment ->
//                        reward / done / info -> auto-reset of finished envs (counter RNG scene
//                        generation, uav_env.py:65-182) -> pair score + observation row of the new
//                        pointer pair (mechanics.py:167-241) -> [B,5,14] window.
//   reset_kernel         UAVEnv.reset for a masked subset (uav_env.py:42-63), one warp per env.
//   pack_scene_kernel    scene injection (host SoA -> device records) + per-scene derived values.
//   score_matrix_kernel  p_final / p_damage [B,N,M] (main.py:38-45 over mechanics.py:167-181).
//   recompute_kernel     fresh J / N0 / sums from the per-target products, a warp per env with
//                        shuffle reductions (uav_env.py:244-293) - drift bound for the running sums.
#pragma once

#include "uavenv_device.cuh"

namespace uavk {

#ifndef UAV_STEP_THREADS
#define UAV_STEP_THREADS 128
#endif
constexpr int kStepThreads = UAV_STEP_THREADS;   // envs per CTA in step_kernel (thread-per-env main phases)
constexpr int kWarpsPerCta = kStepThreads / 32;
constexpr int kResetThreads = 128;
constexpr int kServiceScratchPerWarp = 32 * kObsFloats * 4;   // a service warp's obstacle scratch = its idle window tile (8960 B)

struct StepIO {
    const void *actions;     // [B] int64 (action_bytes = 8) or int8 (action_bytes = 1)
    int32_t action_bytes;
    int32_t actions_bulk;    // 1: the actions sit in mapped host memory - one bulk copy per CTA brings its slice (16 B aligned)
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
// warp-cooperative helpers (all 32 lanes of the warp must call).  A finished env is restarted by the
// warp that owns it, so the step kernel needs no CTA-wide barrier on its hot path.

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
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

// _reset_state_only (envs/uav_env.py:175-182) for env b (scene in `slot`): every UAV available again, every
// lock list empty
__device__ __forceinline__ void warp_clear_allocation(const Params &P, int slot, int b, int lane) {
    int32_t *asg = P.assigned + (size_t)b * P.N;
    for (int i = lane; i < P.N; i += 32) asg[i] = -1;
    TgtRec *T = P.tgt + P.toff(slot, b);
    for (int j = lane; j < P.M; j += 32) {
        T[j].nh = 1.0; T[j].nh_pure = 1.0; T[j].lock_cost = 0.0; T[j].lock_tag = 0;
    }
}

// ---- _generate_scene (envs/uav_env.py:65-173) with the counter RNG, in independent chunks of 32 entities ----
// A scene is a pure function of (seed, global env id, scene index); it is produced chunk by chunk so that the
// pre-generation service can spread one scene over several launches.  Chunk order: target chunks first
// (chunk 0 also draws the obstacles), then UAV chunks (their p_pen needs the obstacles, already visible).

__device__ __forceinline__ int scene_chunks(const Params &P) { return (P.M + 31) / 32 + (P.N + 31) / 32; }

// number of value-6 targets, n2 = randint(1, n_remain+1)   (uav_env.py:121-126)
__device__ __forceinline__ int scene_n2(const Params &P, uint32_t env, uint32_t scene) {
    const int n_remain = P.M - P.M / 2 - 1;
    if (n_remain < 1) return 0;
    return 1 + (int)(((uint64_t)philox4x32(P.seed_lo, P.seed_hi, 0u, S_N2, scene, env).x * (uint64_t)n_remain) >> 32);
}

// total_swarm_cost (uav_env.py:118) and sum of target values (:198) of a generated scene, in closed form
// (type / value class counts are fixed by construction; all terms are exact in fp64)
__device__ __forceinline__ void scene_totals(const Params &P, int b, uint32_t scene, double &total_val, double &total_cost) {
    const int n1 = P.M / 2, n_remain = P.M - n1 - 1, n2 = scene_n2(P, P.env_id_base + (uint32_t)b, scene);
    total_val = 4.0 * n1 + 6.0 * n2 + 8.0 * (n_remain > 0 ? n_remain - n2 : 0) + 16.0;
    total_cost = 1.0 * (P.N - P.N / 4) + 1.25 * (P.N / 4);
}

// obstacles of a scene (uav_env.py:146-170) into Z[K1], I[K2] (global records, or a warp's shared-memory scratch)
__device__ __forceinline__ void warp_generate_obstacles(const Params &P, int b, uint32_t scene, NfzRec *Z, IntRec *I) {
    const int lane = threadIdx.x & 31;
    const uint32_t k0 = P.seed_lo, k1 = P.seed_hi, env = P.env_id_base + (uint32_t)b;
    for (int i = lane; i < P.K1; i += 32) {
        const uint4 a = philox4x32(k0, k1, i, S_NFZ_A, scene, env);
        const uint4 c = philox4x32(k0, k1, i, S_NFZ_B, scene, env);
        Z[i].radius = 5.0 + (10.0 - 5.0) * u53(a.x, a.y);
        Z[i].x = 120.0 + (140.0 - 120.0) * u53(a.z, a.w);
        Z[i].y = 0.0 + (P.map_h - 0.0) * u53(c.x, c.y);
    }
    for (int i = lane; i < P.K2; i += 32) {
        const uint4 a = philox4x32(k0, k1, i, S_INT_A, scene, env);
        const uint4 c = philox4x32(k0, k1, i, S_INT_B, scene, env);
        I[i].x = 140.0 + (160.0 - 140.0) * u53(a.x, a.y);
        I[i].y = 0.0 + (P.map_h - 0.0) * u53(a.z, a.w);
        const double sp = 0.30 + (0.32 - 0.30) * u53(c.x, c.y);
        const double ang = 0.0 + (2.0 * 3.141592653589793 - 0.0) * u53(c.z, c.w);
        I[i].vx = cos(ang) * sp;
        I[i].vy = sin(ang) * sp;
    }
}

// entities [32*c, 32*c+32) of chunk c: target chunks first, then UAV chunks; Z / I: the scene's obstacles, already
// visible to this warp (the UAV chunks need them for p_pen).  A UAV chunk can be produced in two stages (the pre-generation
// service runs them in different launches: p_pen - two acos, four exp - is as long as everything else of the chunk):
// stage 0 = the whole record, 1 = everything but p_pen, 2 = p_pen of the records stage 1 has written
__device__ void warp_generate_chunk(const Params &P, int slot, int b, uint32_t scene, int chunk, uint32_t *s_keys,
                                    const NfzRec *Z, const IntRec *I, int stage = 0) {
    const int lane = threadIdx.x & 31;
    const uint32_t k0 = P.seed_lo, k1 = P.seed_hi, env = P.env_id_base + (uint32_t)b;
    const int N = P.N, M = P.M, tchunks = (M + 31) / 32;
    __syncwarp();
    if (chunk < tchunks) {
        // targets [32*chunk, 32*chunk+32): value by rank among the value keys (uav_env.py:121-129), list position
        // by rank among the list keys (the shuffle of :173: target id i lands at list position rank_i)
        const int i = chunk * 32 + lane;
        const int n1 = M / 2, n_remain = M - n1 - 1, n2 = scene_n2(P, env, scene);
        for (int j = lane; j < M; j += 32) s_keys[j] = philox4x32(k0, k1, j, S_TGT_VAL, scene, env).x;
        __syncwarp();
        double value = 0.0;
        if (i < M) {
            const int q = rank_of(s_keys, M, i);
            value = q < n1 ? 4.0 : (q < n1 + n2 ? 6.0 : (q < n1 + n_remain ? 8.0 : 16.0));
        }
        __syncwarp();
        for (int j = lane; j < M; j += 32) s_keys[j] = philox4x32(k0, k1, j, S_TGT_LIST, scene, env).x;
        __syncwarp();
        if (i < M) {
            const int pos = rank_of(s_keys, M, i);
            const uint4 a = philox4x32(k0, k1, i, S_TGT_POS, scene, env);
            const uint4 v = philox4x32(k0, k1, i, S_TGT_VEL, scene, env);
            TgtRec t;
            t.x = P.tgt_x_lo + (P.tgt_x_hi - P.tgt_x_lo) * u53(a.x, a.y);            // :134
            t.y = 0.0 + (P.map_h - 0.0) * u53(a.z, a.w);                              // :135
            const double vx = (u53(v.x, v.y) - 0.5) * 0.03, vy = (u53(v.z, v.w) - 0.5) * 0.03;  // :139
            t.speed = sqrt(vx * vx + vy * vy);
            t.value = value;
            t.nh = 1.0; t.nh_pure = 1.0; t.lock_cost = 0.0; t.lock_tag = 0; t.id = i;
            P.tgt[P.toff(slot, b) + pos] = t;
            P.tgt_vel[P.toff(slot, b) + pos] = make_double2(vx, vy);
        }
    } else {
        // UAVs [32*c, 32*c+32): N//4 of type 2, uniformly permuted (uav_env.py:81-84); kinematics (:88-111)
        const int i = (chunk - tchunks) * 32 + lane;
        if (stage == 2) {
            if (i < N) {
                UavRec *up = P.uav + P.uoff(slot, b) + i;
                const UavRec u = *up;
                up->p_pen = penetration_prob(P, Z, I, u);
            }
            __syncwarp();
            return;
        }
        for (int j = lane; j < N; j += 32) s_keys[j] = philox4x32(k0, k1, j, S_UAV_TYPE, scene, env).x;
        __syncwarp();
        if (i < N) {
            const int type = rank_of(s_keys, N, i) >= N - N / 4 ? 2 : 1;
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
            const double vx = cos(ang) * real_speed, vy = sin(ang) * real_speed;      // :111
            if (stage == 0) finish_uav(P, Z, I, u, vx, vy);
            else { finish_uav_kinematics(u, vx, vy); u.p_pen = 1.0; }
            P.uav[P.uoff(slot, b) + i] = u;
            P.uav_vel[P.uoff(slot, b) + i] = make_double2(vx, vy);
            P.uav_type[P.uoff(slot, b) + i] = type;
        }
    }
    __syncwarp();
}

// whole scene in one go (reset(), and the fallback when the next scene was not pre-generated in time)
__device__ __noinline__ void warp_generate_scene(const Params &P, int slot, int b, uint32_t scene, uint32_t *s_keys) {
    const int nchunks = scene_chunks(P);
    NfzRec *Z = P.nfz + ((size_t)slot * P.B + b) * P.K1;
    IntRec *I = P.intc + ((size_t)slot * P.B + b) * P.K2;
    warp_generate_obstacles(P, b, scene, Z, I);
    __threadfence_block();  // obstacles visible to the UAV chunks of this warp
    __syncwarp();
    for (int c = 0; c < nchunks; ++c) warp_generate_chunk(P, slot, b, scene, c, s_keys, Z, I);
    warp_clear_allocation(P, slot, b, threadIdx.x & 31);
    __syncwarp();
}

// ---- pre-generation service: the first CTAs of the step grid (blockIdx < service CTAs) -----------------------------
// Every env always has its NEXT scene (index I_GEN >> 1) prepared in the slot it is not playing on, so that the
// scheduled regeneration of main_train.py:79 is a slot flip for its owner: three header stores, NO fence and NO extra
// load on the step's critical path.  That works because the service never acts on a request in the launch in which it
// first sees it, and never publishes in a launch in which it still writes records - kernel boundaries do the ordering:
//   a period = kServicePeriod launches, counted by every service CTA on its own (all counters are equal)
//   launch 0            SCAN     each CTA reads I_GEN / I_NEXT_TAG of its 1024 envs; an env whose other slot does not hold
//                                scene I_GEN >> 1 yet is appended to the period's queue (its I_GEN word is captured)
//   launches 1..P-2     PROCESS  job j = (queue entry, chunk of 32 entities) -> warp j of the round; jobs are independent, so a
//                                whole batch of scenes costs ONE job latency per stage, hidden behind the step's main CTAs;
//                                jobs whose env has moved on since the scan (I_GEN changed, or the scene was delivered by
//                                reset()) are dropped.  A UAV chunk is two jobs in different launches - the record without
//                                p_pen, then p_pen (from the obstacle records the scene's first job wrote): as one job it
//                                outlasted the main CTAs by ~9 us (every 16th step read 27-29 us instead of 18.5)
//   launch P-1          PUBLISH  entries with all chunks in: I_NEXT_TAG = scene index (I_GEN still unchanged)
// Single writer per word: owner -> I_GEN; service -> I_NEXT_TAG, the queues and the other slot's records.  An owner that
// sees the tag flips; its record reads are ordered after the service's writes by >= 1 kernel boundary.  Whatever is not
// ready when an env needs it (RESET_EPISODES of a few steps) is generated in place by the owner's warp - same result.
constexpr int kServicePeriod = 16;
constexpr int kServiceEnvsPerCta = kStepThreads * 8;  // every thread scans 8 envs

__device__ __noinline__ void pregen_service(const Params &P, int service_cta, int n_service, uint32_t *s_keys,
                                            unsigned char *s_scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t s_tick;
    __shared__ int s_n, s_base;
    if (tid == 0) {
        const uint32_t t = P.svc_ctr[service_cta];
        P.svc_ctr[service_cta] = t + 1u;
        s_tick = t;
        s_n = 0;
    }
    __syncthreads();
    const uint32_t phase = s_tick % (uint32_t)kServicePeriod;
    const int q = (int)((s_tick / (uint32_t)kServicePeriod) & 1u);
    int32_t *const q_env = P.q_env + (size_t)q * P.B, *const q_gen = P.q_gen + (size_t)q * P.B;
    int32_t *const q_done = P.q_done + (size_t)q * P.B;
    // jobs of a queue entry: its target chunks, its UAV chunks without p_pen (space A), then p_pen of its UAV chunks (space B,
    // in launches AFTER the last launch of space A: stage 2 reads what stage 1 wrote)
    const int tchunks = (P.M + 31) / 32, uchunks = (P.N + 31) / 32, chunks_a = tchunks + uchunks, nchunks = chunks_a + uchunks;
    if (phase == 0u) {
        // ---- SCAN: 8 consecutive envs per thread = one 32 B run of the I_GEN row and of the I_NEXT_TAG row of a tile
        if (service_cta == 0 && tid == 0) P.q_count[q ^ 1] = 0u;   // the other queue is idle during this whole period
        int *const s_list = reinterpret_cast<int *>(s_scratch);
        const int e0 = service_cta * kServiceEnvsPerCta + tid * 8;
        if (e0 < P.B) {
            const Hdr h = P.header(e0);
            const int4 g0 = *reinterpret_cast<const int4 *>(&h.n(I_GEN)), g1 = *reinterpret_cast<const int4 *>(&h.n(I_GEN) + 4);
            const int4 t0 = *reinterpret_cast<const int4 *>(&h.n(I_NEXT_TAG)), t1 = *reinterpret_cast<const int4 *>(&h.n(I_NEXT_TAG) + 4);
            const int gen[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const int tag[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (e0 + i < P.B && tag[i] != (gen[i] >> 1)) {
                    const int at = atomicAdd(&s_n, 1);
                    s_list[2 * at] = e0 + i;
                    s_list[2 * at + 1] = gen[i];
                }
        }
        __syncthreads();
        const int n = s_n;
        if (n == 0) return;
        if (tid == 0) s_base = (int)atomicAdd(&P.q_count[q], (uint32_t)n);
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            q_env[s_base + i] = s_list[2 * i];
            q_gen[s_base + i] = s_list[2 * i + 1];
            q_done[s_base + i] = 0;
        }
        return;
    }
    const int count = (int)P.q_count[q];
    if (count == 0) return;
    if (phase == (uint32_t)kServicePeriod - 1u) {
        // ---- PUBLISH
        for (int i = service_cta * blockDim.x + tid; i < count; i += n_service * blockDim.x) {
            const int eb = q_env[i], gen = q_gen[i];
            const Hdr h = P.header(eb);
            if (q_done[i] == nchunks && h.n(I_GEN) == gen) h.n(I_NEXT_TAG) = gen >> 1;
        }
        return;
    }
    // ---- PROCESS, round phase-1: one job per warp
    const long long cap = (long long)n_service * kWarpsPerCta, jobs_a = (long long)count * chunks_a, jobs_b = (long long)count * uchunks;
    const long long rounds_a = (jobs_a + cap - 1) / cap, widx = (long long)service_cta * kWarpsPerCta + warp;
    long long round = (long long)phase - 1;
    int entry, chunk, stage;
    if (round < rounds_a) {
        const long long job = round * cap + widx;
        if (job >= jobs_a) return;
        entry = (int)(job / chunks_a); chunk = (int)(job % chunks_a); stage = 1;
    } else {
        const long long job = (round - rounds_a) * cap + widx;
        if (job >= jobs_b) return;
        entry = (int)(job / uchunks); chunk = tchunks + (int)(job % uchunks); stage = 2;
    }
    const int eb = q_env[entry], gen = q_gen[entry];
    const Hdr h = P.header(eb);
    if (*(volatile int32_t *)&h.n(I_GEN) != gen || *(volatile int32_t *)&h.n(I_NEXT_TAG) == (gen >> 1)) return;  // stale
    const int scene = gen >> 1, slot = (gen & 1) ^ 1;
    // the obstacles: chunk 0 generates them (in its shared-memory scratch) and writes the records; the p_pen jobs run in a
    // later launch and read those records - nobody else needs them
    NfzRec *Zg = P.nfz + ((size_t)slot * P.B + eb) * P.K1;
    IntRec *Ig = P.intc + ((size_t)slot * P.B + eb) * P.K2;
    const NfzRec *Z = Zg;
    const IntRec *I = Ig;
    if (chunk == 0) {
        NfzRec *Zs = reinterpret_cast<NfzRec *>(s_scratch + (size_t)warp * kServiceScratchPerWarp);
        IntRec *Is = reinterpret_cast<IntRec *>(Zs + P.K1);
        warp_generate_obstacles(P, eb, (uint32_t)scene, Zs, Is);
        __syncwarp();
        for (int i = lane; i < P.K1; i += 32) Zg[i] = Zs[i];
        for (int i = lane; i < P.K2; i += 32) Ig[i] = Is[i];
    }
    warp_generate_chunk(P, slot, eb, (uint32_t)scene, chunk, s_keys, Z, I, stage);
    if (lane == 0) atomicAdd(&q_done[entry], 1);
}

// ------------------------------------------------------------------------------------------------
// thread-level pieces

// calc_advantage of the pointer pair + its observation row (uav_env.py:184-242 -> mechanics.py:185-241)
// from the two gathered records and the running aggregates.
__device__ __forceinline__ void eval_pointer_pair(const Params &P, const UavRec &u, const TgtRec &t, double cost_sum,
                                                  double covered_val, double total_cost, double total_val,
                                                  double &pf, double &pd, float *row) {
    pd = damage_prob(P, u, t.x, t.y, t.speed);
    pf = pd * u.p_pen;                                                  // mechanics.py:179
    const double inv_tc = 1.0 / (total_cost + 1e-6);
    const double chi_c = cost_sum * inv_tc;                             // uav_env.py:195-196
    const double chi_v = covered_val / (total_val + 1e-6);              // :198-200
    const double chi_mc = t.lock_cost * inv_tc;                         // :202-206
    const double P_prev = 1.0 - t.nh, P_pure = 1.0 - t.nh_pure;         // :226-227
    state_vector(u.cost, t.value, chi_c, chi_v, chi_mc, pf, pd, P_prev, P_prev * t.value, P_pure, row);
}

// capture the new current pair in the header (the next step's accept rule reads only these)
// (t: the record as seen in the current episode, lock_cnt = target_view(t, episode))
__device__ __forceinline__ void store_current_pair(const Hdr &h, const UavRec &u, const TgtRec &t, int lock_cnt, double pf, double pd) {
    h.f(F_CUR_PF) = pf; h.f(F_CUR_PD) = pd; h.f(F_CUR_VALUE) = t.value; h.f(F_CUR_NH) = t.nh; h.f(F_CUR_NHP) = t.nh_pure;
    h.f(F_CUR_LOCK_COST) = t.lock_cost; h.f(F_CUR_UCOST) = u.cost; h.n(I_CUR_LOCK_CNT) = lock_cnt; h.n(I_CUR_TID) = t.id;
}

// ---- async copy / TMA helpers (sm_100a) --------------------------------------------------------------
__device__ __forceinline__ void cp_async_8(void *smem_dst, const void *gmem_src) {   // LDGSTS, 8 bytes
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// TMA bulk store shared::cta -> global (UBLKCP): bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst),
                 "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
// wait until the bulk stores of this thread have READ their shared-memory source (the CTA may then exit)
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// The fused step.  Thread t of a CTA owns env b = blockIdx.x*128 + t for the O(1) state machine.
// Memory-level parallelism is what bounds this kernel (every env touches ~1 KB scattered over a few
// hundred MB), so the dependent chain is kept to TWO round trips and there is no CTA-wide barrier:
//   trip 1  scalar state (SoA, coalesced) + action + ring head; the four older window rows start
//           streaming into the warp's shared-memory tile with cp.async (LDGSTS) as soon as the head is known
//   trip 2  the two 64 B records of the NEW pointer pair - the accept rule itself needs no gather because the
//           current pair's values were captured in the header when its observation row was computed
// The [5,14] windows of a warp (32 x 280 B, contiguous in obs) leave through one TMA bulk store.
// A finished env restarts inside the launch.  The common case (same scene, uav_env.py:175-182) stays on the
// two-trip chain: its owner simply evaluates pointer pair (0,0) with the target's products taken as cleared,
// and the warp clears the allocation arrays afterwards with fire-and-forget stores (warp_soft_reset).
// The scheduled regeneration (every RESET_EPISODES-th episode, main_train.py:79) is a slot flip: the env's
// next scene was prepared in the other storage slot by the pre-generation service (the CTAs past the main
// grid, pregen_service above) - so it, too, stays on the two-trip chain.  Only when that scene is not ready
// (first episodes after a reset with a very short schedule) does the owner's warp generate it in place, in a
// second pass of the evaluation loop (out of line, so nothing but a few scalars is live across the call).

// Eq.21 near-tie (uav_env.py:317): r(X) and r(X') re-summed exactly as _calc_J_X does (uav_env.py:244-269): targets in
// list order, one running fp64 sum; X' differs from X in target m's product (nh2) only.  One thread, rare, out of line.
__device__ __noinline__ void exact_rewards(const Params &P, int slot, int b, int episode, int m, double nh2, double cost_sum,
                                           double cost2, int n0, int n02, double &prev_r, double &new_r) {
    const TgtRec *T = P.tgt + P.toff(slot, b);
    double rev = 0.0, rev2 = 0.0;
    for (int j = 0; j < P.M; ++j) {
        const double nh = tag_current(T[j].lock_tag, episode) ? T[j].nh : 1.0, value = T[j].value;
        rev += (1.0 - nh) * value;                                    // :264-265
        rev2 += (1.0 - (j == m ? nh2 : nh)) * value;
    }
    prev_r = paper_reward(rev - (P.omega * cost_sum), n0, P.M);       // :268, :287-291
    new_r = paper_reward(rev2 - (P.omega * cost2), n02, P.M);
}

// restart bookkeeping of the finished envs in mask (their scene is in slot_of[lane]): stores only
__device__ __noinline__ void warp_soft_reset(const Params &P, unsigned soft_mask, int b0, int my_slot) {
    const int lane = threadIdx.x & 31;
    __syncwarp();  // orders the owners' accept stores before the clears
    while (soft_mask) {
        const int src = __ffs(soft_mask) - 1;
        soft_mask &= soft_mask - 1;
        const int slot = __shfl_sync(kFullMask, my_slot, src);
        warp_clear_allocation(P, slot, b0 + src, lane);
    }
    __syncwarp();
}

// fallback regeneration in place (the next scene was not pre-generated in time) of the envs in mask
__device__ __noinline__ void warp_regen_inline(const Params &P, unsigned mask, int b0, int my_slot, int my_scene,
                                               uint32_t *s_keys) {
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const int slot = __shfl_sync(kFullMask, my_slot, src), scene = __shfl_sync(kFullMask, my_scene, src);
        __syncwarp();
        warp_generate_scene(P, slot, b0 + src, (uint32_t)scene, s_keys);
    }
    __threadfence_block();
}

__global__ void __launch_bounds__(kStepThreads, 512 / kStepThreads) step_kernel(const __grid_constant__ Params P, const StepIO io,
                                                               const int n_service) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ __align__(128) float s_tile[kWarpsPerCta][32 * kObsFloats];
    __shared__ __align__(16) unsigned char s_act[kStepThreads * 8];   // the CTA's actions when they come over PCIe (bulk copy)
    __shared__ uint64_t s_act_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *const s_keys = reinterpret_cast<uint32_t *>(s_dyn) + (size_t)warp * max(P.N, P.M);  // per-warp scratch
    if ((int)blockIdx.x < n_service) {  // service CTAs: prepare next scenes, off the step's critical path
        pregen_service(P, (int)blockIdx.x, n_service, s_keys, reinterpret_cast<unsigned char *>(&s_tile[0][0]));
        return;
    }
    const int n_main = (int)gridDim.x - n_service;
    const int b0 = ((int)blockIdx.x - n_service) * kStepThreads + warp * 32;   // first env of this warp
    const int b = b0 + lane;
    const bool live = b < P.B;
    const int bc = live ? b : P.B - 1;  // clamped index: idle tail lanes issue harmless loads
    const int N = P.N, M = P.M;
    // actions in mapped host memory: a 32 B read per warp is one PCIe read request per warp (2048 of them at c3, +5.7 us on
    // the chain); one bulk copy per CTA is a quarter of the requests and is in flight before anything else
    const int cta_first = b0 - warp * 32, cta_envs = min(kStepThreads, P.B - cta_first);
    const bool bulk_actions = io.actions_bulk && ((cta_envs * io.action_bytes) & 15) == 0;
    if (bulk_actions && tid == 0) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_act_bar), bytes = (uint32_t)(cta_envs * io.action_bytes);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_act)),
                     "l"(static_cast<const unsigned char *>(io.actions) + (size_t)cta_first * io.action_bytes), "r"(bytes), "r"(bar)
                     : "memory");
    }
    const Hdr H = P.header(bc);
    float *tile = s_tile[warp] + lane * kObsFloats;

    // ---- trip 1: everything that depends on b only (all loads issued before anything is consumed) ----
    const uint32_t head_old = P.step_ctr[0];
    int k = H.n(I_K), m = H.n(I_M), nA = H.n(I_NASSIGNED), n0 = H.n(I_NCOVERED), age = H.n(I_AGE);
    double rev = H.f(F_REV), cost_sum = H.f(F_COST_SUM), covered_val = H.f(F_COVERED_VAL);
    double sum_pd = H.f(F_SUM_PD), sum_pf = H.f(F_SUM_PF);
    double total_val = H.f(F_TOTAL_VAL), total_cost = H.f(F_TOTAL_COST);
    const double c_pf = H.f(F_CUR_PF), c_pd = H.f(F_CUR_PD), c_value = H.f(F_CUR_VALUE), c_nh = H.f(F_CUR_NH);
    const double c_nhp = H.f(F_CUR_NHP), c_lock_cost = H.f(F_CUR_LOCK_COST), c_ucost = H.f(F_CUR_UCOST);
    const int c_lock_cnt = H.n(I_CUR_LOCK_CNT), c_tid = H.n(I_CUR_TID);
    const int episode_new = H.n(I_EPISODE) + 1;
    const int gen = H.n(I_GEN), next_tag = H.n(I_NEXT_TAG);
    int slot = gen & 1;                                                  // storage slot of the current scene
    const bool was_finished = !P.auto_reset && H.n(I_FINISHED);
    int64_t action = 0;
    if (!bulk_actions)
        action = io.action_bytes == 8 ? static_cast<const int64_t *>(io.actions)[bc]
                                      : (int64_t) static_cast<const int8_t *>(io.actions)[bc];
    const uint32_t head_new = (head_old + 1u) % (uint32_t)kSeqLen;       // ring slot of this step's row
    // older rows of the window: ring tile [slot][feature pair][32 lanes] -> tile rows 0..3 (time order)
    float2 *const ring = P.ring(bc);
#pragma unroll
    for (int a = kSeqLen - 1; a >= 1; --a) {
        const uint32_t slot = (head_new + (uint32_t)(kSeqLen - a)) % (uint32_t)kSeqLen;
        const float2 *src = ring + slot * (kStateDim / 2) * 32;
        float *dst = tile + (kSeqLen - 1 - a) * kStateDim;
#pragma unroll
        for (int f = 0; f < kStateDim / 2; ++f) cp_async_8(dst + 2 * f, src + f * 32);
    }
    if (tid == 0) {
        // arrive; the last CTA to have READ the head publishes the new one (kept on the device so that
        // CUDA-graph replays stay correct).  The increment depends on head_old, so the read is ordered first.
        if (atomicAdd(&P.step_ctr[1], 1u + (head_old >> 31)) == (unsigned)n_main - 1u) {
            P.step_ctr[1] = 0u;
            P.step_ctr[0] = head_new;
        }
    }

    if (bulk_actions) {   // (uniform per CTA) the barrier publishes the mbarrier's initialisation; the loads above are in flight
        __syncthreads();
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_act_bar);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "ACT_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra ACT_DONE;\n\t"
            "bra ACT_WAIT;\n\t"
            "ACT_DONE:\n\t"
            "}\n" ::"r"(bar) : "memory");
        const int ia = live ? tid : 0;
        action = io.action_bytes == 8 ? reinterpret_cast<const int64_t *>(s_act)[ia] : (int64_t) reinterpret_cast<const int8_t *>(s_act)[ia];
    }
    // ---- accept rule -> state update -> reward / done / info (uav_env.py:295-363, :426-433) ----------
    bool done = false, restarted = false;
    if (live) {
        double reward = 0.0;
        if (!was_finished) {
            double prev_r = paper_reward(rev - (P.omega * cost_sum), n0, M);            // :301
            double cur_r = prev_r;
            bool advance_uav = false;
            if (action == 1) {                                                           // :306
                // tentative X' (:308-310): only target m's product, the cost sum and N0 change
                const double nh2 = c_nh * (1.0 - c_pf);
                const double rev2 = rev + ((1.0 - nh2) - (1.0 - c_nh)) * c_value;
                const double cost2 = cost_sum + c_ucost;
                const int n02 = n0 + (c_lock_cnt == 0);
                double new_r = paper_reward(rev2 - (P.omega * cost2), n02, M);          // :313
                // the carried sums differ from the reference's fresh sums by O(1e-14) relative: inside the tie band the
                // decision (and the reward) come from the exact re-summation.  Only with a cost term: at omega = 0 an Assign
                // can only grow one term of J and N0, so new_r >= prev_r holds under rounding in the reference's summation
                // and in the carried one alike (and near-ties are the NORM there once a target's product has underflowed)
                if (P.omega != 0.0 && fabs(new_r - prev_r) <= P.tie_band * fmax(fabs(new_r), fabs(prev_r)))
                    exact_rewards(P, slot, b, episode_new - 1, m, nh2, cost_sum, cost2, n0, n02, prev_r, new_r);
                if (new_r >= prev_r) {                                                   // :317 (Eq.21)
                    reward = new_r - prev_r;                                             // :321
                    cur_r = new_r;
                    TgtRec *tp = P.tgt + P.toff(slot, b) + m;
                    tp->nh = nh2; tp->nh_pure = c_nhp * (1.0 - c_pd);
                    tp->lock_cost = c_lock_cost + c_ucost; tp->lock_tag = make_tag(c_lock_cnt + 1, episode_new - 1);
                    tp->id = c_tid;   // (unchanged value: the record's second 32 B sector is then written WHOLE - no fill read)
                    P.assigned[(size_t)b * N + k] = c_tid;                               // :308
                    rev = rev2; cost_sum = cost2;
                    if (c_lock_cnt == 0) { covered_val += c_value; n0 = n02; }
                    sum_pd += c_pd; sum_pf += c_pf; nA += 1;
                    advance_uav = true;                                                  // :324-325
                }
            }
            if (advance_uav) { k += 1; m = 0; }
            else {                                                                       // :336-342, :347-352
                m += 1;
                // a UAV that passes every target stays unassigned: its slot may hold the previous episode's value
                if (m >= M) { P.assigned[(size_t)b * N + k] = -1; k += 1; m = 0; }
            }
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
            if (P.auto_reset) restarted = true;
            else H.n(I_FINISHED) = 1;
        }
    }
    const bool inert = live && done && !restarted;  // finished, no auto-reset: zero window, frozen state
    const bool regen = restarted && P.reset_episodes > 0 && (episode_new % P.reset_episodes) == 0;  // main_train.py:79
    bool inline_regen = false;
    if (restarted) {  // pointers and running sums of the next episode (uav_env.py:54-55, :175-182)
        k = 0; m = 0; nA = 0; n0 = 0; age = 0;
        rev = 0.0; cost_sum = 0.0; covered_val = 0.0; sum_pd = 0.0; sum_pf = 0.0;
        H.n(I_EPISODE) = episode_new;
        if (regen) {
            const int scene = gen >> 1;
            scene_totals(P, b, (uint32_t)scene, total_val, total_cost);
            H.f(F_TOTAL_VAL) = total_val; H.f(F_TOTAL_COST) = total_cost;
            // the pre-generated scene: flip the storage slot (its records were written >= 1 launch ago) - or generate in
            // place (second pass).  The service picks the new I_GEN up at its next scan: no fence, no request store.
            if (next_tag == scene) slot ^= 1;
            else inline_regen = true;
            H.n(I_GEN) = ((scene + 1) << 1) | slot;
        }
    }
    const bool soft = restarted && !inline_regen;
    // a restart leaves the allocation arrays alone: target records carry their episode (target_view) and UAV slots at or
    // past the pointer are read as unassigned.  Every 2^16-th episode of an env the arrays are wiped, so that a record's
    // 18-bit episode tag can never alias
    const bool wipe = soft && (episode_new & 0xffff) == 0;
    const int episode_eval = restarted ? episode_new : episode_new - 1;   // the episode the new pointer pair belongs to

    // ---- trip 2 + evaluation.  pass 0: every env but the (rare) inline-regenerated ones; pass 1: those ----
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            const unsigned soft_mask = __ballot_sync(kFullMask, wipe);
            if (soft_mask) warp_soft_reset(P, soft_mask, b0, slot);
            const unsigned regen_mask = __ballot_sync(kFullMask, inline_regen);
            if (regen_mask == 0u) break;
            warp_regen_inline(P, regen_mask, b0, slot, gen >> 1, s_keys);
        }
        if (live && (pass == 0 ? (!done || soft) : inline_regen)) {
            const UavRec u = P.uav[P.uoff(slot, b) + k];
            TgtRec t = P.tgt[P.toff(slot, b) + m];   // sees this thread's own accept store on target m
            const int t_cnt = target_view(t, episode_eval);   // (a restarted env sees every record as cleared)
            cp_async_wait_all();                        // ring rows have landed in the tile (gathers still in flight)
            const int nprev = age < kSeqLen - 1 ? age : kSeqLen - 1;
#pragma unroll
            for (int a = kSeqLen - 1; a >= 1; --a) {     // rows older than the episode are zero (uav_env.py:58-61)
                if (a > nprev) {
                    float *dst = tile + (kSeqLen - 1 - a) * kStateDim;
#pragma unroll
                    for (int f = 0; f < kStateDim; f += 2) *reinterpret_cast<float2 *>(dst + f) = make_float2(0.f, 0.f);
                }
            }
            double pf, pd;
            float row[kStateDim];
            eval_pointer_pair(P, u, t, cost_sum, covered_val, total_cost, total_val, pf, pd, row);
            store_current_pair(H, u, t, t_cnt, pf, pd);
            float2 *dsth = ring + head_new * (kStateDim / 2) * 32;
#pragma unroll
            for (int f = 0; f < kStateDim / 2; ++f) {
                const float2 v = make_float2(row[2 * f], row[2 * f + 1]);
                dsth[f * 32] = v;
                *reinterpret_cast<float2 *>(tile + (kSeqLen - 1) * kStateDim + 2 * f) = v;
            }
            H.n(I_K) = k; H.n(I_M) = m; H.n(I_NASSIGNED) = nA; H.n(I_NCOVERED) = n0; H.n(I_AGE) = nprev + 1;
            H.f(F_REV) = rev; H.f(F_COST_SUM) = cost_sum; H.f(F_COVERED_VAL) = covered_val;
            H.f(F_SUM_PD) = sum_pd; H.f(F_SUM_PF) = sum_pf;
        }
    }
    cp_async_wait_all();
    if (inert) {
#pragma unroll
        for (int f = 0; f < kObsFloats; f += 2) *reinterpret_cast<float2 *>(tile + f) = make_float2(0.f, 0.f);
        if (!was_finished) {  // the step that finished the episode: freeze the final pointers
            H.n(I_K) = k; H.n(I_M) = m; H.n(I_NASSIGNED) = nA; H.n(I_NCOVERED) = n0; H.n(I_AGE) = 0;
            H.f(F_REV) = rev; H.f(F_COST_SUM) = cost_sum; H.f(F_COVERED_VAL) = covered_val;
            H.f(F_SUM_PD) = sum_pd; H.f(F_SUM_PF) = sum_pf;
        }
    }
    // ---- window out: one TMA bulk store per warp (32 x 280 B contiguous in obs) ------------------------
    fence_proxy_async_smem();
    __syncwarp();
    {
        const int nenv = min(32, P.B - b0);
        if (nenv > 0) {
            float *dst = io.obs + (size_t)b0 * kObsFloats;
            const float *src = s_tile[warp];
            const uint32_t bytes = (uint32_t)nenv * kObsFloats * sizeof(float);
            if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (bytes & 15u) == 0) {
                if (lane == 0) { bulk_store(dst, src, bytes); bulk_store_wait_read(); }
            } else {  // ragged tail / unaligned caller buffer
                for (int i = lane; i < nenv * kObsFloats; i += 32) dst[i] = src[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// UAVEnv.reset (envs/uav_env.py:42-63) for the masked envs: one warp per env.
// mode: 0 = state only, 1 = generate a new scene, 2 = scene already packed (load_scene)

__global__ void __launch_bounds__(kResetThreads) reset_kernel(const __grid_constant__ Params P, int mode,
                                                               const uint8_t *mask, int first_env, int count,
                                                               float *obs) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ float s_row[kResetThreads / 32][kStateDim];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_dyn) + (size_t)warp * max(P.N, P.M);
    for (int e = blockIdx.x * wpb + warp; e < count; e += gridDim.x * wpb) {
        const int b = first_env + e;
        if (mask && !mask[b]) continue;  // uniform across the warp
        const Hdr H = P.header(b);
        const int gen = H.n(I_GEN), slot = gen & 1;
        __syncwarp();
        const bool ahead = P.auto_reset && P.reset_episodes > 0;  // keep the NEXT scene prepared in the other slot
        if (mode == 1) {
            // a new scene now, in place (index gen >> 1) ...
            const int scene = gen >> 1;
            warp_generate_scene(P, slot, b, (uint32_t)scene, s_keys);
            // ... and the one after it into the other slot, so that the first scheduled regeneration is a flip
            // (a reset is off the hot path; the step-time service only has to keep up with later flips)
            if (ahead) warp_generate_scene(P, slot ^ 1, b, (uint32_t)(scene + 1), s_keys);
            if (lane == 0) {
                double tv, tc;
                scene_totals(P, b, (uint32_t)scene, tv, tc);
                H.f(F_TOTAL_VAL) = tv; H.f(F_TOTAL_COST) = tc;
                H.n(I_GEN) = ((scene + 1) << 1) | slot;
                if (ahead) H.n(I_NEXT_TAG) = scene + 1;
            }
        } else if (mode == 2) {
            // injected scene (already packed into the current slot); prepare the generated scene that follows it
            if (ahead) {
                warp_generate_scene(P, slot ^ 1, b, (uint32_t)(gen >> 1), s_keys);
                if (lane == 0) H.n(I_NEXT_TAG) = gen >> 1;
            }
            warp_clear_allocation(P, slot, b, lane);
        } else {
            warp_clear_allocation(P, slot, b, lane);
        }
        __syncwarp();
        if (lane == 0) {
            const double total_cost = H.f(F_TOTAL_COST), total_val = H.f(F_TOTAL_VAL);
            double pf, pd;
            float row[kStateDim];
            const UavRec u = P.uav[P.uoff(slot, b)];
            const TgtRec t = P.tgt[P.toff(slot, b)];       // (just cleared)
            eval_pointer_pair(P, u, t, 0.0, 0.0, total_cost, total_val, pf, pd, row);
            store_current_pair(H, u, t, 0, pf, pd);
            H.n(I_K) = 0; H.n(I_M) = 0; H.n(I_NASSIGNED) = 0; H.n(I_NCOVERED) = 0; H.n(I_AGE) = 1;
            H.f(F_REV) = 0.0; H.f(F_COST_SUM) = 0.0; H.f(F_COVERED_VAL) = 0.0;
            H.f(F_SUM_PD) = 0.0; H.f(F_SUM_PF) = 0.0;
            H.n(I_FINISHED) = 0;
            H.n(I_EPISODE) = (mode == 0) ? H.n(I_EPISODE) + 1 : 1;
            const uint32_t head = P.step_ctr[0] % (uint32_t)kSeqLen;
            float2 *dsth = P.ring(b) + head * (kStateDim / 2) * 32;
            for (int f = 0; f < kStateDim / 2; ++f) dsth[f * 32] = make_float2(row[2 * f], row[2 * f + 1]);
            for (int f = 0; f < kStateDim; ++f) s_row[warp][f] = row[f];
        }
        __syncwarp();
        if (obs) {
            float *o = obs + (size_t)b * kObsFloats;
            for (int i = lane; i < kObsFloats; i += 32)
                o[i] = i < (kSeqLen - 1) * kStateDim ? 0.0f : s_row[warp][i - (kSeqLen - 1) * kStateDim];
        }
        __syncwarp();
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
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int e = blockIdx.x * wpb + warp; e < count; e += gridDim.x * wpb) {
        const int b = first_env + e;
        const int slot = P.header(b).n(I_GEN) & 1;
        for (int i = lane; i < P.K1; i += 32) {
            NfzRec z; z.x = s.nfz_x[(size_t)e * P.K1 + i]; z.y = s.nfz_y[(size_t)e * P.K1 + i];
            z.radius = s.nfz_radius ? s.nfz_radius[(size_t)e * P.K1 + i] : 0.0;
            P.nfz[((size_t)slot * P.B + b) * P.K1 + i] = z;
        }
        for (int i = lane; i < P.K2; i += 32) {
            IntRec r; r.x = s.int_x[(size_t)e * P.K2 + i]; r.y = s.int_y[(size_t)e * P.K2 + i];
            r.vx = s.int_vx[(size_t)e * P.K2 + i]; r.vy = s.int_vy[(size_t)e * P.K2 + i];
            P.intc[((size_t)slot * P.B + b) * P.K2 + i] = r;
        }
        __syncwarp();
        double cost_part = 0.0, val_part = 0.0;
        for (int i = lane; i < P.N; i += 32) {
            const size_t g = (size_t)e * P.N + i;
            UavRec u;
            u.x = s.uav_x[g]; u.y = s.uav_y[g];
            u.load = s.uav_load[g]; u.cost = s.uav_cost[g];
            finish_uav(P, slot, b, u, s.uav_vx[g], s.uav_vy[g]);
            P.uav[P.uoff(slot, b) + i] = u;
            P.uav_vel[P.uoff(slot, b) + i] = make_double2(s.uav_vx[g], s.uav_vy[g]);
            P.uav_type[P.uoff(slot, b) + i] = s.uav_type ? s.uav_type[g] : 1;
            cost_part += u.cost;
        }
        for (int j = lane; j < P.M; j += 32) {
            const size_t g = (size_t)e * P.M + j;
            TgtRec t;
            t.x = s.tgt_x[g]; t.y = s.tgt_y[g];
            const double vx = s.tgt_vx[g], vy = s.tgt_vy[g];
            t.speed = sqrt(vx * vx + vy * vy);
            t.value = s.tgt_value[g];
            t.nh = 1.0; t.nh_pure = 1.0; t.lock_cost = 0.0; t.lock_tag = 0; t.id = s.tgt_id[g];
            P.tgt[P.toff(slot, b) + j] = t;
            P.tgt_vel[P.toff(slot, b) + j] = make_double2(vx, vy);
            val_part += t.value;
        }
        const double total_cost = warp_sum(cost_part);
        const double total_val = warp_sum(val_part);
        if (lane == 0) { P.header(b).f(F_TOTAL_COST) = total_cost; P.header(b).f(F_TOTAL_VAL) = total_val; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Score matrix (main.py:38-45 over mechanics.py:167-181): p_final / p_damage [B,N,M].
// A warp owns (env, chunk of 32 targets): lane = target (its x, y, speed stay in registers; stores are coalesced along
// the target index), the loop runs over the env's UAVs, whose 64 B records arrive as warp-uniform loads.  Index
// arithmetic and the header read happen once per warp, not once per pair: the first version (a thread per pair with
// 64-bit div / mod and its own record loads) was ISSUE-bound - 81 % of the issue slots, fp64 pipe 47 % active
// (profiles/r2_score_matrix_ncu.md).

template <typename OutT>
__global__ void __launch_bounds__(256) score_matrix_kernel(const __grid_constant__ Params P, OutT *p_final,
                                                            OutT *p_damage) {
    const int lane = threadIdx.x & 31;
    const int chunks = (P.M + 31) >> 5;
    const long long items = (long long)P.B * chunks;
    const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < items; w += wstride) {
        const int b = (int)(w / chunks), m = (int)(w % chunks) * 32 + lane;
        const bool ok = m < P.M;
        const int slot = P.header(b).n(I_GEN) & 1;
        const TgtRec *t = P.tgt + P.toff(slot, b) + (ok ? m : P.M - 1);
        const double tx = t->x, ty = t->y, ts = t->speed;
        const UavRec *U = P.uav + P.uoff(slot, b);
        const size_t out0 = (size_t)b * P.N * P.M + m;
#pragma unroll 2
        for (int k = 0; k < P.N; ++k) {
            const UavRec u = U[k];                                    // warp-uniform address
            const double pd = damage_prob(P, u, tx, ty, ts);
            if (ok) {
                if (p_damage) p_damage[out0 + (size_t)k * P.M] = (OutT)pd;
                if (p_final) p_final[out0 + (size_t)k * P.M] = (OutT)(pd * u.p_pen);
            }
        }
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
        const int slot = P.header(b).n(I_GEN) & 1;
        const TgtRec *T = P.tgt + P.toff(slot, b);
        const int episode = P.header(b).n(I_EPISODE), kptr = P.header(b).n(I_K);
        for (int j = lane; j < P.M; j += 32) {
            TgtRec t = T[j];
            const int cnt = target_view(t, episode);
            rev += (1.0 - t.nh) * t.value;
            if (cnt > 0) { cval += t.value; n0 += 1; }
        }
        const UavRec *U = P.uav + P.uoff(slot, b);
        const int32_t *asg = P.assigned + (size_t)b * P.N;
        for (int i = lane; i < min(P.N, kptr); i += 32) if (asg[i] >= 0) cost += U[i].cost;   // slots at / past the pointer are stale
        for (int o = 16; o > 0; o >>= 1) {
            rev += __shfl_xor_sync(0xffffffffu, rev, o);
            cval += __shfl_xor_sync(0xffffffffu, cval, o);
            cost += __shfl_xor_sync(0xffffffffu, cost, o);
            n0 += __shfl_xor_sync(0xffffffffu, n0, o);
        }
        if (lane == 0) {
            const double J_fresh = rev - (P.omega * cost);
            const double J_run = P.header(b).f(F_REV) - (P.omega * P.header(b).f(F_COST_SUM));
            worst = fmax(worst, fabs(J_fresh - J_run));
            if (n0 != P.header(b).n(I_NCOVERED)) worst = fmax(worst, 1e30);  // integer state must agree exactly
            if (fix) { P.header(b).f(F_REV) = rev; P.header(b).f(F_COST_SUM) = cost; P.header(b).f(F_COVERED_VAL) = cval; }
        }
    }
    if (lane == 0 && max_abs_diff && worst > 0.0) {
        // doubles >= 0 order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(max_abs_diff), (unsigned long long)__double_as_longlong(worst));
    }
}

// uavenv_set_episode_counters: the new counters, and the records written in an env's current episode follow it (their
// episode tag is what makes them current).  One warp per env.
__global__ void __launch_bounds__(256) set_episode_kernel(const __grid_constant__ Params P, const int32_t *episodes, int first_env,
                                                           int count) {
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int e = wglobal; e < count; e += nwarps) {
        const int b = first_env + e;
        const Hdr h = P.header(b);
        const int old_ep = h.n(I_EPISODE), new_ep = episodes[e];
        TgtRec *T = P.tgt + P.toff(h.n(I_GEN) & 1, b);
        for (int j = lane; j < P.M; j += 32)
            if (tag_current(T[j].lock_tag, old_ep)) T[j].lock_tag = make_tag(tag_count(T[j].lock_tag), new_ep);
        __syncwarp();
        if (lane == 0) h.n(I_EPISODE) = new_ep;
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
