// policy_fused.cu - hand-written tcgen05 kernels of the policy forward (sm_100a).  See tcgen05_util.cuh.
#include "uavpolicy_b200.h"

#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "policy_weights.cuh"
#include "tcgen05_util.cuh"

namespace {
using namespace tc;
using namespace uavp;

// ------------------------------------------------------------------------------------------------------------
// Fused encoder block.  One CTA (8 warps) owns a tile of 25 samples = 125 token rows (3 rows of padding make the
// M = 128 of tcgen05.mma).  Activations never leave the SM: the layer input X lives in shared memory as a UMMA
// A-operand (canonical K-major tile), each GEMM accumulates in TMEM, the epilogue threads (one per row: TMEM lane =
// row) apply bias / ReLU / residual + LayerNorm entirely thread-locally and write the next A-operand back to shared
// memory.  Weights stream L2 -> shared memory by warps 4-7 while warps 0-3 run the previous epilogue.
//
//   shared memory (224 KB):  sX 32 KB | sQ sK sV 3 x 32 KB (attention output overwrites Q; the FFN hidden tile
//                            [128 x 256] later aliases sQ+sK) | sW 96 KB (one weight matrix at a time)
//   TMEM: 512 columns; QKV uses [0,384), the other GEMMs [0,256) / [0,128) / [0,64)
constexpr int kTileSamples = 25, kTileRows = kTileSamples * S;     // 125 valid rows of the 128
constexpr int kFusedThreads = 256;
constexpr uint32_t kTileBytes = 128 * D * 2;                       // a [128 x 128] bf16 canonical tile
constexpr size_t kFusedSmem = 4 * (size_t)kTileBytes + 96 * 1024;  // sX + sQ/sK/sV + sW

struct Phase { uint32_t parity = 0; };

// issue D[tmem cols 0..N) = A[128 x K] * W[N x K]^T as K/16 k-steps (N <= 384 split at 256); thread 0 only
__device__ __forceinline__ void issue_gemm(uint32_t tmem, const unsigned char *sA, const unsigned char *sWt, int N, int K,
                                           uint64_t *mbar) {
    const uint32_t sbo = (uint32_t)K * 16;
    for (int j = 0; j < K / 16; ++j) {
        const uint64_t ad = smem_desc(smem_u32(sA) + j * 256, 128, sbo);
        for (int n0 = 0; n0 < N; n0 += 256) {
            const int nn = min(256, N - n0);
            const uint64_t bd = smem_desc(smem_u32(sWt) + (n0 >> 3) * sbo + j * 256, 128, sbo);
            mma_bf16(tmem + n0, ad, bd, instr_desc_bf16(128, nn), j > 0);
        }
    }
    mma_commit(mbar);
}

__device__ __forceinline__ void unpack8(const uint4 &q, float *o) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
        o[2 * i] = __low2float(v); o[2 * i + 1] = __high2float(v);
    }
}
__device__ __forceinline__ uint4 pack8(const float *v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t *>(&t);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Epilogues run on all 8 warps: warp w reads TMEM lanes 32*(w%4).. (its rows) and the column half w/4.

// plain layer: out = act(acc + bias) -> bf16; column c goes to tile (c / tile_cols) of dst, canonical row length Kout
__device__ __forceinline__ void epilogue_bias_act(uint32_t tmem_lane, int row, const float *__restrict__ bias, int c_begin,
                                                  int c_end, bool relu, unsigned char *dst, int Kout, int tile_cols) {
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_lane + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {                       // (the flat parameter buffer is 8 B aligned at best)
            const float2 b2 = __ldg(reinterpret_cast<const float2 *>(bias + c0 + i));
            v[i] += b2.x; v[i + 1] += b2.y;
            if (relu) { v[i] = fmaxf(v[i], 0.0f); v[i + 1] = fmaxf(v[i + 1], 0.0f); }
        }
        unsigned char *tile = dst + (c0 / tile_cols) * kTileBytes;
#pragma unroll
        for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4 *>(tile + canon_off(row, (c0 % tile_cols) + g * 8, Kout)) = pack8(v + g * 8);
    }
}

// residual layer: X[row] <- LayerNorm(X[row] + acc + bias) * g + beta (post-LN, eps 1e-5), in place.  The two column
// halves of a row exchange their partial sums through shared memory (all 256 threads call: contains a CTA barrier).
__device__ __forceinline__ void epilogue_residual_ln(uint32_t tmem_lane, int row, int half, const float *__restrict__ bias,
                                                     const float *__restrict__ g, const float *__restrict__ beta,
                                                     unsigned char *sX, float (*s_part)[2][128]) {
    float t[64];
    float sum = 0.0f, sq = 0.0f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        float v[32];
        tmem_ld32(tmem_lane + c0, v);
#pragma unroll
        for (int gch = 0; gch < 4; ++gch) {
            float x[8];
            unpack8(*reinterpret_cast<const uint4 *>(sX + canon_off(row, c0 + gch * 8, D)), x);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float u = v[gch * 8 + i] + __ldg(bias + c0 + gch * 8 + i) + x[i];
                t[cc * 32 + gch * 8 + i] = u;
                sum += u; sq = fmaf(u, u, sq);
            }
        }
    }
    s_part[half][0][row] = sum; s_part[half][1][row] = sq;
    __syncthreads();
    sum += s_part[half ^ 1][0][row]; sq += s_part[half ^ 1][1][row];
    const float mean = sum * (1.0f / D);
    const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.0f) + 1e-5f);
#pragma unroll
    for (int gch = 0; gch < 8; ++gch) {
        const int c = half * 64 + gch * 8;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (t[gch * 8 + i] - mean) * rstd * __ldg(g + c + i) + __ldg(beta + c + i);
        *reinterpret_cast<uint4 *>(sX + canon_off(row, c, D)) = pack8(o);
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1)
fused_block_kernel(const float *__restrict__ obs, int B, BlockW w_actor, HeadW head_actor, __nv_bfloat16 *__restrict__ hh_actor,
                   BlockW w_critic, HeadW head_critic, __nv_bfloat16 *__restrict__ hh_critic, int *__restrict__ work_counter) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar, wbar;                       // MMA completion / weight staging (bulk TMA) barriers
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_item;
    __shared__ uint8_t s_pad[128];
    unsigned char *sX = smem, *sQ = smem + kTileBytes, *sK = sQ + kTileBytes, *sV = sK + kTileBytes;
    unsigned char *sH = sQ;                               // [128 x 256] canonical, aliases sQ + sK after attention
    unsigned char *sW = smem + 4 * kTileBytes;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __shared__ float s_part[2][2][128];
    const int row = tid & 127, half = warp >> 2;          // thread = (row, column half); TMEM lane = row
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&mbar, 1); mbar_init(&wbar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t parity = 0, wparity = 0;
    const int num_tiles = (B + kTileSamples - 1) / kTileSamples;

    // work items = (network, tile), handed out dynamically, the two-layer critic tiles first (longest first)
    for (;;) {
        if (tid == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= 2 * num_tiles) break;
        const bool is_critic = item < num_tiles;
        const BlockW &w = is_critic ? w_critic : w_actor;
        const HeadW &head = is_critic ? head_critic : head_actor;
        __nv_bfloat16 *head_hidden = is_critic ? hh_critic : hh_actor;
        const int tile = is_critic ? item : item - num_tiles;
        const int s0 = tile * kTileSamples;
        const int nsamp = min(kTileSamples, B - s0), nrows = nsamp * S;
        // ---- embedding: X = relu(obs W^T + b) + pos (transformer_net.py:24-30,57-59); key-padding mask (:52-54).
        //      On the tensor cores with K = 32: the fp32 observation row is split into bf16 hi + lo parts
        //      (columns 0..13 and 16..29) against the weight duplicated in both halves, so the product keeps ~16
        //      mantissa bits of the input.  A operand staged in sQ, B operand in sK; Win of layer 0 streams into sW. ----
        if (tid == 0) {
            bulk_load(sK, w.emb_w2p, D * 32 * 2, &wbar);                  // embedding weights (B operand, K = 32)
        }
        if (tid < 128) {
            const int r = tid;
            const bool valid = r < nrows;
            float o[16], asum = 0.0f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                o[j] = (valid && j < F) ? obs[((size_t)s0 * S + r) * F + j] : 0.0f;
                asum += fabsf(o[j]);
            }
            s_pad[r] = (valid && asum == 0.0f && r % S != S - 1) ? 1 : 0;
            float hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                hi[j] = __bfloat162float(__float2bfloat16(o[j]));
                lo[j] = o[j] - hi[j];
            }
            *reinterpret_cast<uint4 *>(sQ + canon_off(r, 0, 32)) = pack8(hi);
            *reinterpret_cast<uint4 *>(sQ + canon_off(r, 8, 32)) = pack8(hi + 8);
            *reinterpret_cast<uint4 *>(sQ + canon_off(r, 16, 32)) = pack8(lo);
            *reinterpret_cast<uint4 *>(sQ + canon_off(r, 24, 32)) = pack8(lo + 8);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        mbar_wait(&wbar, wparity); wparity ^= 1;
        if (tid == 0) {
            bulk_load(sW, w.layer[0].in_wp, 3 * D * D * 2, &wbar);       // Win of layer 0 streams in behind the embedding
            issue_gemm(tmem, sQ, sK, D, 32, &mbar);
        }
        mbar_wait(&mbar, parity); parity ^= 1;
        tc_fence_after();
        {
            const bool valid = row < nrows;
            const float *pos = w.pos + (row % S) * D;
            for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_lane + c0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    v[i] = valid ? fmaxf(v[i] + __ldg(w.emb_b + c0 + i), 0.0f) + __ldg(pos + c0 + i) : 0.0f;
#pragma unroll
                for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4 *>(sX + canon_off(row, c0 + g * 8, D)) = pack8(v + g * 8);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        for (int l = 0; l < w.layers; ++l) {
            const LayerW &L = w.layer[l];
            // ---- QKV = X Win^T (N = 384) -> sQ | sK | sV; out-proj weights stream in behind the epilogue ----
            tc_fence_after();
            mbar_wait(&wbar, wparity); wparity ^= 1;                      // Win has landed in sW
            if (tid == 0) issue_gemm(tmem, sX, sW, 3 * D, D, &mbar);
            mbar_wait(&mbar, parity); parity ^= 1;
            tc_fence_after();
            if (tid == 0) bulk_load(sW, L.out_wp, D * D * 2, &wbar);
            epilogue_bias_act(tmem_lane, row, L.in_b, half * 192, half * 192 + 192, false, sQ, D, D);
            tc_fence_before();
            __syncthreads();
            // ---- attention over the 5-token window, per (sample, head, query); output overwrites the Q slice ----
            for (int item = tid; item < nsamp * H * S; item += kFusedThreads) {
                const int smp = item / (H * S), h = (item / S) % H, i = item % S;
                const int r = smp * S + i;
                float q[DH], sc[S], mx = -INFINITY;
                unsigned char *pq = sQ + canon_off(r, h * DH, D);
                unpack8(*reinterpret_cast<const uint4 *>(pq), q);
                unpack8(*reinterpret_cast<const uint4 *>(pq + 128), q + 8);
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    float kk[DH], sdot = 0.0f;
                    const unsigned char *pk = sK + canon_off(smp * S + j, h * DH, D);
                    unpack8(*reinterpret_cast<const uint4 *>(pk), kk);
                    unpack8(*reinterpret_cast<const uint4 *>(pk + 128), kk + 8);
#pragma unroll
                    for (int e = 0; e < DH; ++e) sdot = fmaf(q[e], kk[e], sdot);
                    sc[j] = s_pad[smp * S + j] ? -INFINITY : sdot * 0.25f;
                    mx = fmaxf(mx, sc[j]);
                }
                float den = 0.0f, o[DH];
#pragma unroll
                for (int j = 0; j < S; ++j) { sc[j] = __expf(sc[j] - mx); den += sc[j]; }
                const float inv = 1.0f / den;
#pragma unroll
                for (int e = 0; e < DH; ++e) o[e] = 0.0f;
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    float vv[DH];
                    const unsigned char *pv = sV + canon_off(smp * S + j, h * DH, D);
                    unpack8(*reinterpret_cast<const uint4 *>(pv), vv);
                    unpack8(*reinterpret_cast<const uint4 *>(pv + 128), vv + 8);
                    const float pj = sc[j] * inv;
#pragma unroll
                    for (int e = 0; e < DH; ++e) o[e] = fmaf(pj, vv[e], o[e]);
                }
                *reinterpret_cast<uint4 *>(pq) = pack8(o);
                *reinterpret_cast<uint4 *>(pq + 128) = pack8(o + 8);
            }
            fence_async_smem();
            __syncthreads();
            // ---- out-proj + residual + LayerNorm1 (in place in sX); FFN1 weights stream in ----
            tc_fence_after();
            mbar_wait(&wbar, wparity); wparity ^= 1;
            if (tid == 0) issue_gemm(tmem, sQ, sW, D, D, &mbar);
            mbar_wait(&mbar, parity); parity ^= 1;
            tc_fence_after();
            if (tid == 0) bulk_load(sW, L.l1_wp, FF * D * 2, &wbar);
            epilogue_residual_ln(tmem_lane, row, half, L.out_b, L.n1_w, L.n1_b, sX, s_part);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            // ---- FFN1 + ReLU -> sH [128 x 256]; FFN2 weights ([128 x 256]) stream in ----
            tc_fence_after();
            mbar_wait(&wbar, wparity); wparity ^= 1;
            if (tid == 0) issue_gemm(tmem, sX, sW, FF, D, &mbar);
            mbar_wait(&mbar, parity); parity ^= 1;
            tc_fence_after();
            if (tid == 0) bulk_load(sW, L.l2_wp, D * FF * 2, &wbar);
            epilogue_bias_act(tmem_lane, row, L.l1_b, half * 128, half * 128 + 128, true, sH, FF, FF);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            // ---- FFN2 + residual + LayerNorm2 (in place in sX); the next GEMM's weights stream in ----
            tc_fence_after();
            mbar_wait(&wbar, wparity); wparity ^= 1;
            if (tid == 0) issue_gemm(tmem, sH, sW, D, FF, &mbar);
            mbar_wait(&mbar, parity); parity ^= 1;
            tc_fence_after();
            if (tid == 0) {
                if (l + 1 < w.layers) bulk_load(sW, w.layer[l + 1].in_wp, 3 * D * D * 2, &wbar);
                else bulk_load(sW, head.w1p, HID * D * 2, &wbar);        // last layer: the head's first layer
            }
            epilogue_residual_ln(tmem_lane, row, half, L.l2_b, L.n2_w, L.n2_b, sX, s_part);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
        }
        // ---- head first layer: relu(W1 z + b1) for the newest token of every sample (transformer_net.py:106-108) ----
        tc_fence_after();
        mbar_wait(&wbar, wparity); wparity ^= 1;
        if (tid == 0) issue_gemm(tmem, sX, sW, HID, D, &mbar);
        mbar_wait(&mbar, parity); parity ^= 1;
        tc_fence_after();
        {   // tcgen05.ld is warp-collective: every lane loads, only the newest-token rows store (32 columns per half)
            const bool keep = row < nrows && row % S == S - 1;
            __nv_bfloat16 *dst = head_hidden + (size_t)(s0 + row / S) * HID;
            const int c0 = half * 32;
            float v[32];
            tmem_ld32(tmem_lane + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + __ldg(head.b1 + c0 + i), 0.0f);
            if (keep) {
#pragma unroll
                for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4 *>(dst + c0 + g * 8) = pack8(v + g * 8);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_free(tmem, 512);
}

// self-test: D[128,N] (fp32) = A[128,K] W[N,K]^T with one CTA: canonical smem operands, K/16 tcgen05.mma steps
// (N = 384 as a 256 + 128 pair), accumulator read back from TMEM.  Validates descriptors and the TMEM lane mapping.
__global__ void __launch_bounds__(128) gemm_tile_selftest_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *W, float *D,
                                                                  int N, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    unsigned char *sA = smem, *sW = smem + 128 * K * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
    load_canon(sA, A, 128, K, K, 128, tid, 128);
    load_canon(sW, W, N, K, K, N, tid, 128);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t sbo = (uint32_t)K * 16;
        for (int j = 0; j < K / 16; ++j) {
            const uint64_t ad = smem_desc(smem_u32(sA) + j * 256, 128, sbo);
            for (int n0 = 0; n0 < N; n0 += 256) {
                const int nn = min(256, N - n0);
                const uint64_t bd = smem_desc(smem_u32(sW) + (n0 >> 3) * sbo + j * 256, 128, sbo);
                mma_bf16(tmem + n0, ad, bd, instr_desc_bf16(128, nn), j > 0);
            }
        }
        mma_commit(&mbar);
    }
    mbar_wait(&mbar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}
}  // namespace

namespace uavp {
int fused_block_prepare() {
    return cudaFuncSetAttribute(fused_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem) == cudaSuccess ? 0 : -2;
}
int launch_fused_blocks(const float *d_obs, int B, const BlockW &actor, const HeadW &actor_head, __nv_bfloat16 *hh_actor,
                        const BlockW &critic, const HeadW &critic_head, __nv_bfloat16 *hh_critic, int *d_work_counter,
                        cudaStream_t stream) {
    const int items = 2 * ((B + kTileSamples - 1) / kTileSamples);
    static int num_sms = 0;                       // one persistent CTA per SM of the device
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -2;
    }
    if (cudaMemsetAsync(d_work_counter, 0, sizeof(int), stream) != cudaSuccess) return -2;
    fused_block_kernel<<<items < num_sms ? items : num_sms, kFusedThreads, kFusedSmem, stream>>>(
        d_obs, B, actor, actor_head, hh_actor, critic, critic_head, hh_critic, d_work_counter);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
}  // namespace uavp

extern "C" int uavpolicy_selftest_gemm_tile(const void *d_A, const void *d_W, float *d_D, int32_t N, int32_t K, void *stream) {
    if (!d_A || !d_W || !d_D || N <= 0 || N > 384 || N % 16 || (K != 128 && K != 256)) return -1;
    const size_t smem = (size_t)(128 + N) * K * 2;
    if (cudaFuncSetAttribute(gemm_tile_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    gemm_tile_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)d_A, (const __nv_bfloat16 *)d_W, d_D, N, K);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
