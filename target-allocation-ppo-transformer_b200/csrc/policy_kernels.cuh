// policy_kernels.cuh - the hand-written (non-GEMM) kernels of the policy network shared by the rollout forward
// (policy_forward.cu) and the PPO update (policy_train.cu): weight packing, embedding + positions + padding mask,
// attention over the 5-token window, residual + LayerNorm.  Internal linkage: each translation unit gets its own copy.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "policy_weights.cuh"

namespace {
using namespace uavp;

// element offset of (row n, column k) in the UMMA canonical K-major order (tcgen05_util.cuh: 8x8 core matrices)
__device__ __forceinline__ int canon_elem(int n, int k, int K) { return (n >> 3) * (K * 8) + (k >> 3) * 64 + (n & 7) * 8 + (k & 7); }

// [128 x 14] fp32 embedding weight -> [128 x 32] bf16, canonical order, the 14 columns duplicated at 0.. and 16..
__global__ void emb_w2_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * 32) return;
    const int d = i / 32, c = i % 32, j = c % 16;
    out[canon_elem(d, c, 32)] = __float2bfloat16(j < F ? w[d * F + j] : 0.0f);
}

// row-major fp32 [N x K] -> bf16 in canonical order (one bulk TMA copy then stages a whole B operand)
__global__ void pack_canon_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ out, int N, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * K) return;
    out[canon_elem(i / K, i % K, K)] = __float2bfloat16(w[i]);
}

__global__ void f32_to_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16(src[i]);
}

// embedding of both nets: E = relu(obs W^T + b) + pos (transformer_net.py:24-30,57-59) and the key-padding mask
// (rows that are all zero, newest row never: :52-54).  One CTA = 64 tokens; thread = (net, feature pair), so a warp
// stores 128 contiguous bytes per token.
constexpr int kEmbTok = 64;
__global__ void __launch_bounds__(D) embed_kernel(const float *__restrict__ obs, int R, BlockW a, BlockW c,
                                                  __nv_bfloat16 *__restrict__ Ea, __nv_bfloat16 *__restrict__ Ec,
                                                  uint8_t *__restrict__ pad, __nv_bfloat16 *__restrict__ obs16 = nullptr,
                                                  uint32_t *__restrict__ relu_mask = nullptr) {
    // obs16 / relu_mask (training only): a bf16 copy of the observation rows zero-padded to 64 columns (the B operand of
    // the embedding's weight-gradient product) and the ReLU activity bits of both networks, [R][net][4 words]:
    // word 2h + parity, bit l  <->  feature 64 h + 2 l + parity
    __shared__ float s_obs[kEmbTok][F];
    const int t0 = blockIdx.x * kEmbTok;
    for (int i = threadIdx.x; i < kEmbTok * F; i += D) {
        const int t = t0 + i / F;
        s_obs[i / F][i % F] = t < R ? obs[(size_t)t * F + i % F] : 0.0f;
    }
    __syncthreads();
    if (threadIdx.x < kEmbTok && t0 + threadIdx.x < R) {
        float sum = 0.0f;
        for (int j = 0; j < F; ++j) sum += fabsf(s_obs[threadIdx.x][j]);
        pad[t0 + threadIdx.x] = (sum == 0.0f && (t0 + threadIdx.x) % S != S - 1) ? 1 : 0;
    }
    if (obs16) {
        for (int i = threadIdx.x; i < kEmbTok * 32; i += D) {
            const int tok = i >> 5, c2 = (i & 31) * 2;
            if (t0 + tok < R)
                *reinterpret_cast<__nv_bfloat162 *>(obs16 + (size_t)(t0 + tok) * 64 + c2) =
                    __floats2bfloat162_rn(c2 < F ? s_obs[tok][c2] : 0.0f, c2 + 1 < F ? s_obs[tok][c2 + 1] : 0.0f);
        }
    }
    const bool critic = threadIdx.x >= D / 2;
    const int d = (threadIdx.x % (D / 2)) * 2;                    // features d, d+1
    const BlockW &w = critic ? c : a;
    __nv_bfloat16 *E = critic ? Ec : Ea;
    float w0[F], w1[F];
#pragma unroll
    for (int j = 0; j < F; ++j) { w0[j] = w.emb_w[d * F + j]; w1[j] = w.emb_w[(d + 1) * F + j]; }
    const float b0 = w.emb_b[d], b1 = w.emb_b[d + 1];
    float p0[S], p1[S];
#pragma unroll
    for (int p = 0; p < S; ++p) { p0[p] = w.pos[p * D + d]; p1[p] = w.pos[p * D + d + 1]; }
    for (int i = 0; i < kEmbTok; ++i) {
        const int t = t0 + i;
        if (t >= R) break;
        float x0 = b0, x1 = b1;
#pragma unroll
        for (int j = 0; j < F; ++j) { x0 = fmaf(w0[j], s_obs[i][j], x0); x1 = fmaf(w1[j], s_obs[i][j], x1); }
        if (relu_mask) {
            const unsigned m0 = __ballot_sync(0xffffffffu, x0 > 0.0f), m1 = __ballot_sync(0xffffffffu, x1 > 0.0f);
            if ((threadIdx.x & 31) == 0) {
                uint32_t *dst = relu_mask + ((size_t)t * 2 + (critic ? 1 : 0)) * 4 + ((threadIdx.x >> 5) & 1) * 2;
                dst[0] = m0; dst[1] = m1;
            }
        }
        const int p = t % S;
        float q0 = p0[0], q1 = p1[0];
#pragma unroll
        for (int k = 1; k < S; ++k) if (p == k) { q0 = p0[k]; q1 = p1[k]; }
        *reinterpret_cast<__nv_bfloat162 *>(E + (size_t)t * D + d) = __floats2bfloat162_rn(fmaxf(x0, 0.0f) + q0, fmaxf(x1, 0.0f) + q1);
    }
}

__device__ __forceinline__ void load16(const __nv_bfloat16 *p, float *out) {  // 16 bf16 = 32 B
    const uint4 a = reinterpret_cast<const uint4 *>(p)[0], b = reinterpret_cast<const uint4 *>(p)[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
        out[2 * i] = __low2float(v); out[2 * i + 1] = __high2float(v);
    }
}
__device__ __forceinline__ void store16(__nv_bfloat16 *p, const float *v) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t *>(&t);
    }
    reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// softmax(q k^T / sqrt(16), key padding mask) v for one (sample, head, query): q, and 5 keys / values of 16 dims
__device__ __forceinline__ void attend(const float *q, const __nv_bfloat16 *k0, const __nv_bfloat16 *v0, size_t stride,
                                       const uint8_t *pad, float *out) {
    float sc[S], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        float kk[DH], s = 0.0f;
        load16(k0 + j * stride, kk);
#pragma unroll
        for (int e = 0; e < DH; ++e) s = fmaf(q[e], kk[e], s);
        sc[j] = pad[j] ? -INFINITY : s * 0.25f;
        mx = fmaxf(mx, sc[j]);
    }
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < S; ++j) { sc[j] = __expf(sc[j] - mx); den += sc[j]; }
    const float inv = 1.0f / den;
#pragma unroll
    for (int e = 0; e < DH; ++e) out[e] = 0.0f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        float vv[DH];
        load16(v0 + j * stride, vv);
        const float p = sc[j] * inv;
#pragma unroll
        for (int e = 0; e < DH; ++e) out[e] = fmaf(p, vv[e], out[e]);
    }
}

// all five queries (an inner encoder layer): QKV [R,384] -> ATT [R,128]; thread = (sample, head, query)
__global__ void attn_full_kernel(const __nv_bfloat16 *__restrict__ QKV, const uint8_t *__restrict__ pad, int B,
                                 __nv_bfloat16 *__restrict__ ATT) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H * S) return;
    const int b = idx / (H * S), h = (idx / S) % H, i = idx % S;
    const __nv_bfloat16 *base = QKV + (size_t)b * S * 3 * D + h * DH;
    float q[DH], o[DH];
    load16(base + (size_t)i * 3 * D, q);
    attend(q, base + D, base + 2 * D, 3 * D, pad + b * S, o);
    store16(ATT + ((size_t)b * S + i) * D + h * DH, o);
}

// last query only (the last encoder layer): Q [B,128], KV [R,256] -> ATT [B,128]; thread = (sample, head)
__global__ void attn_last_kernel(const __nv_bfloat16 *__restrict__ Q, const __nv_bfloat16 *__restrict__ KV,
                                 const uint8_t *__restrict__ pad, int B, __nv_bfloat16 *__restrict__ ATT) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, h = idx % H;
    float q[DH], o[DH];
    load16(Q + (size_t)b * D + h * DH, q);
    const __nv_bfloat16 *base = KV + (size_t)b * S * 2 * D + h * DH;
    attend(q, base, base + D, 2 * D, pad + b * S, o);
    store16(ATT + (size_t)b * D + h * DH, o);
}

// out[r] = LayerNorm(x[r] + y[r]) * g + beta over 128 features (post-LN, eps 1e-5); one warp per row
__global__ void add_ln_kernel(const __nv_bfloat16 *__restrict__ x, int64_t x_stride, const __nv_bfloat16 *__restrict__ y,
                              const float *__restrict__ g, const float *__restrict__ beta, int rows,
                              __nv_bfloat16 *__restrict__ out) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    const uint2 xa = *reinterpret_cast<const uint2 *>(x + (size_t)r * x_stride + lane * 4);
    const uint2 ya = *reinterpret_cast<const uint2 *>(y + (size_t)r * D + lane * 4);
    float v[4];
    {
        const __nv_bfloat162 x0 = *reinterpret_cast<const __nv_bfloat162 *>(&xa.x), x1 = *reinterpret_cast<const __nv_bfloat162 *>(&xa.y);
        const __nv_bfloat162 y0 = *reinterpret_cast<const __nv_bfloat162 *>(&ya.x), y1 = *reinterpret_cast<const __nv_bfloat162 *>(&ya.y);
        v[0] = __low2float(x0) + __low2float(y0); v[1] = __high2float(x0) + __high2float(y0);
        v[2] = __low2float(x1) + __low2float(y1); v[3] = __high2float(x1) + __high2float(y1);
    }
    float s = v[0] + v[1] + v[2] + v[3];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + 1e-5f);
    const float *gp = g + lane * 4, *bp = beta + lane * 4;   // (scalar loads: the flat parameter buffer is only 4 B aligned)
    const __nv_bfloat162 o0 = __floats2bfloat162_rn(v[0] * rstd * gp[0] + bp[0], v[1] * rstd * gp[1] + bp[1]);
    const __nv_bfloat162 o1 = __floats2bfloat162_rn(v[2] * rstd * gp[2] + bp[2], v[3] * rstd * gp[3] + bp[3]);
    uint2 ov;
    ov.x = *reinterpret_cast<const uint32_t *>(&o0); ov.y = *reinterpret_cast<const uint32_t *>(&o1);
    *reinterpret_cast<uint2 *>(out + (size_t)r * D + lane * 4) = ov;
}

}  // namespace
