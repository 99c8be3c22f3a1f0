// policy_train.cu - forward + backward of the two transformer trunks for the PPO update (agents/ppo.py:126-160:
// policy.evaluate(...) and loss.backward() of networks/transformer_net.py:47-65) on sm_100a.
//
// Same formulation as the rollout forward (policy_forward.cu): the last encoder layer of a block computes K/V for all
// five tokens but everything else for the newest token only.  Activations are bf16 and are KEPT for the backward;
// LayerNorm statistics, softmax, all reductions and every parameter gradient are fp32.
//   activation-gradient GEMMs  dX = dY W     -> uavp::gemm_bias_act on pre-transposed bf16 weights (tcgen05)
//   weight-gradient GEMMs      dW += dY^T X  -> uavp::wgrad (policy_wgrad.cu: split-K, TMA ring, accumulator resident in TMEM)
//   ReLU backward of the FFN                 -> fused into the epilogue of the activation-gradient GEMM (uavp::gemm_drelu)
//   everything else (LayerNorm / attention / embedding backward, LayerNorm parameter gradients, heads, PPO loss) is
//   hand-written below; bias gradients come from the LayerNorm backward where a bias feeds a LayerNorm, and from the
//   weight-gradient kernel (as dY^T 1 on the tensor cores) everywhere else.
// Three entry levels: uavtrain_forward / _backward stop at the trunks' last-token features ([n,2,128] out, their gradient
// in); uavtrain_forward_heads / _backward_heads also run the two MLP heads (logits [n,2] + value [n] out, their
// gradients in); uavtrain_ppo_loss turns the heads' outputs into the PPO loss statistics and those gradients.
#include "uavpolicy_b200.h"

#include <cstdarg>
#include <cstdio>
#include <new>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "policy_gemm.cuh"
#include "policy_kernels.cuh"
#include "policy_weights.cuh"

namespace uavp {
int wgrad_prepare();
int wgrad(const __nv_bfloat16 *dY, int64_t ld_dy, const __nv_bfloat16 *X, int64_t ld_x, int rows, int Nout, int Kin, float *dW,
          int nout_valid, float *dbias, int num_sms, cudaStream_t stream, int out_ld, int kin_valid);
}  // namespace uavp

namespace {
using bf16 = __nv_bfloat16;

__device__ __forceinline__ void unpack4(uint2 v, float *o) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&v.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&v.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
__device__ __forceinline__ uint2 pack4(const float *v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t *>(&a); o.y = *reinterpret_cast<const uint32_t *>(&b);
    return o;
}
__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------- forward extras

__device__ __forceinline__ void unpack8(uint4 v, float *o) {
    unpack4(make_uint2(v.x, v.y), o);
    unpack4(make_uint2(v.z, v.w), o + 4);
}
__device__ __forceinline__ uint4 pack8(const float *v) {
    const uint2 a = pack4(v), b = pack4(v + 4);
    return make_uint4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float half_warp_sum(float v) {           // over the 16 lanes that share a row
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// out = LayerNorm(x + y) * g + beta, keeping what the backward needs: the normalised row x^ (bf16) and 1/sigma.
// Half a warp per row, 8 features (16 bytes) per lane.  out16 (bf16, dense) and out32 (fp32, row stride out32_stride)
// are both optional.
__global__ void __launch_bounds__(256) add_ln_train_kernel(const bf16 *__restrict__ x, int64_t x_stride, const bf16 *__restrict__ y,
                                                           const float *__restrict__ g, const float *__restrict__ beta, int rows,
                                                           bf16 *__restrict__ out16, float *__restrict__ out32, int64_t out32_stride,
                                                           bf16 *__restrict__ xhat, float *__restrict__ rstd_out) {
    const int hl = threadIdx.x & 15;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool live = r < rows;
    const int rc = live ? r : rows - 1;                               // (idle tail lanes still take part in the shuffles)
    float a[8], b[8], v[8];
    unpack8(*reinterpret_cast<const uint4 *>(x + (size_t)rc * x_stride + hl * 8), a);
    unpack8(*reinterpret_cast<const uint4 *>(y + (size_t)rc * D + hl * 8), b);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = a[i] + b[i]; s += v[i]; }
    const float mean = half_warp_sum(s) * (1.0f / D);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
    const float rstd = rsqrtf(half_warp_sum(q) * (1.0f / D) + 1e-5f);
    if (!live) return;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] *= rstd; o[i] = v[i] * g[hl * 8 + i] + beta[hl * 8 + i]; }
    *reinterpret_cast<uint4 *>(xhat + (size_t)r * D + hl * 8) = pack8(v);
    if (hl == 0) rstd_out[r] = rstd;
    if (out16) *reinterpret_cast<uint4 *>(out16 + (size_t)r * D + hl * 8) = pack8(o);
    if (out32) {
        float4 *dst = reinterpret_cast<float4 *>(out32 + (size_t)r * out32_stride + hl * 8);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]); dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// fp32 row-major [N,K] -> bf16 copy and / or bf16 transposed copy [K,N]; one launch converts every GEMM weight
struct PrepJob { const float *src; bf16 *dst, *dst_t; int N, K; };
constexpr int kMaxPrepJobs = 24;
struct PrepJobs { PrepJob job[kMaxPrepJobs]; };
__global__ void prep_weights_kernel(const __grid_constant__ PrepJobs jobs) {
    const PrepJob &j = jobs.job[blockIdx.y];
    const int n = j.N * j.K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const bf16 v = __float2bfloat16(j.src[i]);
        if (j.dst) j.dst[i] = v;
        if (j.dst_t) j.dst_t[(size_t)(i % j.K) * j.N + i / j.K] = v;
    }
}

// ---------------------------------------------------------------------------------------------- backward kernels

// block-level reduction of per-lane column partial sums: part[NV] of every warp (lane owns columns col(lane, i)) ->
// one atomicAdd per column per CTA.  s_red: [warps][C] floats.
template <int NV, int C, typename ColFn>
__device__ __forceinline__ void block_colsum_flush(const float *part, float *s_red, float *dst, ColFn col) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) s_red[warp * C + col(lane, i)] = part[i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.0f;
        for (int w = 0; w < nwarps; ++w) s += s_red[w * C + c];
        atomicAdd(dst + c, s);
    }
    __syncthreads();
}

// LayerNorm backward over 128 features.  dy = (DY32 ? fp32 rows of stride dy_stride : bf16 dense) [+ add (bf16 dense)];
//   dz = rstd * (g*dy - mean(g*dy) - x^ * mean(g*dy*x^))       -> bf16 dense (gradient of the pre-norm sum: it flows
//   into the residual branch and into the GEMM branch alike)
//   g_gamma += sum_r dy*x^,  g_beta += sum_r dy,  g_bias += sum_r dz (bias of the linear layer that fed the sum)
// A warp walks `rows_per_warp` consecutive rows with its lanes on fixed columns, so the three column sums stay in
// registers until the end.
template <bool DY32>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const void *__restrict__ dy_, int64_t dy_stride, const bf16 *__restrict__ add,
                                                     const bf16 *__restrict__ xhat, const float *__restrict__ rstd,
                                                     const float *__restrict__ gamma, int rows, int rows_per_warp,
                                                     bf16 *__restrict__ dz, float *__restrict__ g_gamma,
                                                     float *__restrict__ g_beta, float *__restrict__ g_bias) {
    __shared__ float s_red[8 * D];
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int r0 = gw * rows_per_warp, r1 = min(rows, r0 + rows_per_warp);
    float gm[4], ag[4] = {0, 0, 0, 0}, ab[4] = {0, 0, 0, 0}, az[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) gm[i] = gamma[lane * 4 + i];
    for (int r = r0; r < r1; ++r) {
        float dy[4], xh[4];
        if (DY32) {
            const float4 t = *reinterpret_cast<const float4 *>(static_cast<const float *>(dy_) + (size_t)r * dy_stride + lane * 4);
            dy[0] = t.x; dy[1] = t.y; dy[2] = t.z; dy[3] = t.w;
        } else {
            unpack4(*reinterpret_cast<const uint2 *>(static_cast<const bf16 *>(dy_) + (size_t)r * D + lane * 4), dy);
        }
        if (add) {
            float t[4];
            unpack4(*reinterpret_cast<const uint2 *>(add + (size_t)r * D + lane * 4), t);
#pragma unroll
            for (int i = 0; i < 4; ++i) dy[i] += t[i];
        }
        unpack4(*reinterpret_cast<const uint2 *>(xhat + (size_t)r * D + lane * 4), xh);
        const float rs = rstd[r];
        float s1 = 0.0f, s2 = 0.0f, gd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { gd[i] = gm[i] * dy[i]; s1 += gd[i]; s2 = fmaf(gd[i], xh[i], s2); }
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        s1 *= (1.0f / D); s2 *= (1.0f / D);
        float z[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            z[i] = rs * (gd[i] - s1 - xh[i] * s2);
            ag[i] = fmaf(dy[i], xh[i], ag[i]); ab[i] += dy[i]; az[i] += z[i];
        }
        *reinterpret_cast<uint2 *>(dz + (size_t)r * D + lane * 4) = pack4(z);
    }
    auto col = [](int l, int i) { return l * 4 + i; };
    block_colsum_flush<4, D>(ag, s_red, g_gamma, col);
    block_colsum_flush<4, D>(ab, s_red, g_beta, col);
    block_colsum_flush<4, D>(az, s_red, g_bias, col);
}

// attention backward.  dS_ij = p_ij (dO_i.v_j - sum_l p_il dO_i.v_l) / 4;  dQ_i = sum_j dS_ij k_j;  dK_j = sum_i dS_ij q_i;
// dV_j = sum_i p_ij dO_i.  The softmax rows are recomputed from Q / K (nothing but Q, K, V was kept by the forward).
struct AttnBwdArgs {
    const bf16 *q, *k, *v, *gout;      // q rows: (b*NQ+i)*q_stride; k/v rows: (b*S+j)*kv_stride; gout rows dense 128
    bf16 *gq, *gk, *gv;                // same strides as q / k / v
    int64_t q_stride, kv_stride;
    const uint8_t *pad;
    int n;
};

// softmax row of one query over the 5 keys of (sample b, head h), and dS for that query
__device__ __forceinline__ void attn_row_grads(const AttnBwdArgs &a, int64_t b, int h, const float *q, const float *go, float *p, float *ds) {
    const uint8_t *pad = a.pad + b * S;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        float kk[DH], s = 0.0f;
        load16(a.k + (b * S + j) * a.kv_stride + h * DH, kk);
#pragma unroll
        for (int e = 0; e < DH; ++e) s = fmaf(q[e], kk[e], s);
        p[j] = pad[j] ? -INFINITY : s * 0.25f;
        mx = fmaxf(mx, p[j]);
    }
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < S; ++j) { p[j] = __expf(p[j] - mx); den += p[j]; }
    const float inv = 1.0f / den;
    float dot = 0.0f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        float vv[DH], s = 0.0f;
        load16(a.v + (b * S + j) * a.kv_stride + h * DH, vv);
#pragma unroll
        for (int e = 0; e < DH; ++e) s = fmaf(go[e], vv[e], s);
        p[j] *= inv;
        ds[j] = s;
        dot = fmaf(p[j], s, dot);
    }
#pragma unroll
    for (int j = 0; j < S; ++j) ds[j] = p[j] * (ds[j] - dot) * 0.25f;
}

// last layer (one query per sample): thread = (sample, head) owns dQ and all five dK / dV rows of its head
__global__ void __launch_bounds__(256) attn_bwd_last_kernel(const AttnBwdArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.n * H) return;
    const int64_t b = idx / H;
    const int h = (int)(idx % H);
    float q[DH], go[DH], p[S], ds[S], dq[DH];
    load16(a.q + b * a.q_stride + h * DH, q);
    load16(a.gout + b * D + h * DH, go);
    attn_row_grads(a, b, h, q, go, p, ds);
#pragma unroll
    for (int e = 0; e < DH; ++e) dq[e] = 0.0f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        float kk[DH], dk[DH], dv[DH];
        load16(a.k + (b * S + j) * a.kv_stride + h * DH, kk);
#pragma unroll
        for (int e = 0; e < DH; ++e) { dq[e] = fmaf(ds[j], kk[e], dq[e]); dk[e] = ds[j] * q[e]; dv[e] = p[j] * go[e]; }
        store16(a.gk + (b * S + j) * a.kv_stride + h * DH, dk);
        store16(a.gv + (b * S + j) * a.kv_stride + h * DH, dv);
    }
    store16(a.gq + b * a.q_stride + h * DH, dq);
}

// inner layer (five queries): warp = sample, lane = 4 feature dimensions of a head (policy_kernels.cuh: attn_full_kernel).
// P is recomputed from Q, K; dP = dO V^T is a second folded 5 x 5 reduction; dS = P o (dP - rowsum(P o dP)) / 4; then
// dQ = dS K, dK = dS^T Q, dV = P^T dO are plain 4-dimension FMAs per lane.  Every input row is loaded once (coalesced
// 256 B per warp access) and every gradient row stored once.
__global__ void __launch_bounds__(256, 2) attn_bwd_full_kernel(const AttnBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int64_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < a.n; b += nwarps) {
        Rows4 q, k, v, go;
#pragma unroll
        for (int t = 0; t < S; ++t) {
            q.t[t] = *reinterpret_cast<const uint2 *>(a.q + (b * S + t) * a.q_stride + lane * 4);
            k.t[t] = *reinterpret_cast<const uint2 *>(a.k + (b * S + t) * a.kv_stride + lane * 4);
            v.t[t] = *reinterpret_cast<const uint2 *>(a.v + (b * S + t) * a.kv_stride + lane * 4);
            go.t[t] = *reinterpret_cast<const uint2 *>(a.gout + (b * S + t) * D + lane * 4);
        }
        float p[S][S], ds[S][S];
        head_dots(q, k, p);
        head_softmax(p, a.pad + b * S);
        head_dots(go, v, ds);                                  // dP_ij = dO_i . v_j
#pragma unroll
        for (int i = 0; i < S; ++i) {
            float dot = 0.0f;
#pragma unroll
            for (int j = 0; j < S; ++j) dot = fmaf(p[i][j], ds[i][j], dot);
#pragma unroll
            for (int j = 0; j < S; ++j) ds[i][j] = p[i][j] * (ds[i][j] - dot) * 0.25f;
        }
        {   // dQ_i = sum_j dS_ij k_j
            float kf[S][4];
#pragma unroll
            for (int j = 0; j < S; ++j) unpack4f(k.t[j], kf[j]);
#pragma unroll
            for (int i = 0; i < S; ++i) {
                float o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) o[e] = fmaf(ds[i][j], kf[j][e], o[e]);
                *reinterpret_cast<uint2 *>(a.gq + (b * S + i) * a.q_stride + lane * 4) = pack4f(o);
            }
        }
        {   // dK_j = sum_i dS_ij q_i ;  dV_j = sum_i p_ij dO_i
            float qf[S][4], gf[S][4];
#pragma unroll
            for (int i = 0; i < S; ++i) { unpack4f(q.t[i], qf[i]); unpack4f(go.t[i], gf[i]); }
#pragma unroll
            for (int j = 0; j < S; ++j) {
                float dk[4] = {0.0f, 0.0f, 0.0f, 0.0f}, dv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int i = 0; i < S; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { dk[e] = fmaf(ds[i][j], qf[i][e], dk[e]); dv[e] = fmaf(p[i][j], gf[i][e], dv[e]); }
                *reinterpret_cast<uint2 *>(a.gk + (b * S + j) * a.kv_stride + lane * 4) = pack4f(dk);
                *reinterpret_cast<uint2 *>(a.gv + (b * S + j) * a.kv_stride + lane * 4) = pack4f(dv);
            }
        }
    }
}

// dX[b*5+4] += t[b] + u[b]  (last layer: the newest token's row also receives the Q-projection and residual gradients)
__global__ void scatter_add_last_kernel(bf16 *__restrict__ dX, const bf16 *__restrict__ t, const bf16 *__restrict__ u, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // one thread = 4 features
    if (i >= n * (D / 4)) return;
    const int b = i / (D / 4), c = (i % (D / 4)) * 4;
    float x[4], y[4], z[4];
    bf16 *dst = dX + ((size_t)b * S + S - 1) * D + c;
    unpack4(*reinterpret_cast<const uint2 *>(dst), x);
    unpack4(*reinterpret_cast<const uint2 *>(t + (size_t)b * D + c), y);
    unpack4(*reinterpret_cast<const uint2 *>(u + (size_t)b * D + c), z);
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] += y[k] + z[k];
    *reinterpret_cast<uint2 *>(dst) = pack4(x);
}

// embedding backward: E = relu(obs W^T + b) + pos.  The forward left the ReLU activity bits (policy_kernels.cuh:
// embed_kernel) and a bf16 copy of the observations padded to 64 columns, so what remains here is elementwise:
// dPre = (g1 [+ g2]) masked -> bf16 rows, dpos = per-position column sums of the unmasked gradient (registers, lane = 4
// features).  dW = dPre^T obs16 and db = dPre^T 1 then come from the tensor-core weight-gradient kernel (Kin = 64).
__global__ void __launch_bounds__(256) embed_mask_kernel(const bf16 *__restrict__ g1, const bf16 *__restrict__ g2,
                                                         const uint32_t *__restrict__ mask /* this net's 4 words of row 0; row stride 8 */,
                                                         int n, int samples_per_warp, bf16 *__restrict__ dpre, float *__restrict__ g_pos) {
    __shared__ float s_red[S][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < S * D; i += blockDim.x) (&s_red[0][0])[i] = 0.0f;
    __syncthreads();
    const int gw = blockIdx.x * 8 + warp;
    const int b0 = gw * samples_per_warp, b1 = min(n, b0 + samples_per_warp);
    float dpos[S][4];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int f = 0; f < 4; ++f) dpos[s][f] = 0.0f;
    for (int b = b0; b < b1; ++b) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const size_t row = (size_t)b * S + s, off = row * D + lane * 4;
            float g[4];
            unpack4(*reinterpret_cast<const uint2 *>(g1 + off), g);
            if (g2) {
                float t[4];
                unpack4(*reinterpret_cast<const uint2 *>(g2 + off), t);
#pragma unroll
                for (int f = 0; f < 4; ++f) g[f] += t[f];
            }
            const uint4 m = *reinterpret_cast<const uint4 *>(mask + row * 8);             // word f, bit lane <-> feature 4 lane + f
            const unsigned mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                dpos[s][f] += g[f];
                g[f] = ((mw[f] >> lane) & 1u) ? g[f] : 0.0f;
            }
            *reinterpret_cast<uint2 *>(dpre + off) = pack4(g);
        }
    }
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int f = 0; f < 4; ++f) atomicAdd(&s_red[s][lane * 4 + f], dpos[s][f]);
    __syncthreads();
    for (int i = threadIdx.x; i < S * D; i += blockDim.x) atomicAdd(g_pos + i, (&s_red[0][0])[i]);
}

// second layers of the two heads (transformer_net.py:78-91): logits = W2a ha + b2a, value = W2c hc + b2c; one thread per sample
__global__ void __launch_bounds__(256) head_out_kernel(const bf16 *__restrict__ Ha, const bf16 *__restrict__ Hc, const float *__restrict__ w2a,
                                                       const float *__restrict__ b2a, const float *__restrict__ w2c,
                                                       const float *__restrict__ b2c, int n, float *__restrict__ logits,
                                                       float *__restrict__ value) {
    __shared__ float s_w[3][HID];
    for (int i = threadIdx.x; i < HID; i += blockDim.x) { s_w[0][i] = w2a[i]; s_w[1][i] = w2a[HID + i]; s_w[2][i] = w2c[i]; }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    float l0 = b2a[0], l1 = b2a[1], vv = b2c[0];
#pragma unroll
    for (int c = 0; c < HID / DH; ++c) {
        float xa[DH], xc[DH];
        load16(Ha + (size_t)b * HID + c * DH, xa);
        load16(Hc + (size_t)b * HID + c * DH, xc);
#pragma unroll
        for (int e = 0; e < DH; ++e) {
            l0 = fmaf(s_w[0][c * DH + e], xa[e], l0);
            l1 = fmaf(s_w[1][c * DH + e], xa[e], l1);
            vv = fmaf(s_w[2][c * DH + e], xc[e], vv);
        }
    }
    logits[2 * b] = l0; logits[2 * b + 1] = l1;
    value[b] = vv;
}

// backward of both second head layers and of the ReLU in front of them.  A warp walks a run of samples, lane = 2 of
// the 64 hidden units: dH = (W2^T dout) * (h > 0) goes out as bf16 rows of 128 (columns 64.. stay zero: the weight-
// gradient kernel consumes 128-column blocks); dW2, db2 and db1 (column sums of dH) are reduced in registers.
__global__ void __launch_bounds__(256) head_bwd_kernel(const bf16 *__restrict__ Ha, const bf16 *__restrict__ Hc, const float *__restrict__ w2a,
                                                       const float *__restrict__ w2c, const float *__restrict__ dlogits,
                                                       const float *__restrict__ dvalue, int n, int samples_per_warp,
                                                       bf16 *__restrict__ dHa, bf16 *__restrict__ dHc, float *__restrict__ g_w2a,
                                                       float *__restrict__ g_b2a, float *__restrict__ g_b1a, float *__restrict__ g_w2c,
                                                       float *__restrict__ g_b2c, float *__restrict__ g_b1c) {
    constexpr int kAcc = 10;                               // per lane: 2 x {dW2a row 0, dW2a row 1, dW2c, db1a, db1c}
    __shared__ float s_red[kAcc / 2][HID];
    __shared__ float s_b2[3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (kAcc / 2) * HID; i += blockDim.x) (&s_red[0][0])[i] = 0.0f;
    if (threadIdx.x < 3) s_b2[threadIdx.x] = 0.0f;
    __syncthreads();
    const int gw = blockIdx.x * 8 + warp;
    const int b0 = gw * samples_per_warp, b1 = min(n, b0 + samples_per_warp);
    const int j = lane * 2;
    const float wa0[2] = {w2a[j], w2a[j + 1]}, wa1[2] = {w2a[HID + j], w2a[HID + j + 1]}, wc[2] = {w2c[j], w2c[j + 1]};
    float acc[5][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}}, sb[3] = {0, 0, 0};
    for (int b = b0; b < b1; ++b) {
        const float dl0 = dlogits[2 * b], dl1 = dlogits[2 * b + 1], dv = dvalue[b];
        const __nv_bfloat162 ha2 = *reinterpret_cast<const __nv_bfloat162 *>(Ha + (size_t)b * HID + j);
        const __nv_bfloat162 hc2 = *reinterpret_cast<const __nv_bfloat162 *>(Hc + (size_t)b * HID + j);
        const float ha[2] = {__low2float(ha2), __high2float(ha2)}, hc[2] = {__low2float(hc2), __high2float(hc2)};
        float da[2], dc[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            da[k] = ha[k] > 0.0f ? dl0 * wa0[k] + dl1 * wa1[k] : 0.0f;
            dc[k] = hc[k] > 0.0f ? dv * wc[k] : 0.0f;
            acc[0][k] = fmaf(dl0, ha[k], acc[0][k]); acc[1][k] = fmaf(dl1, ha[k], acc[1][k]); acc[2][k] = fmaf(dv, hc[k], acc[2][k]);
            acc[3][k] += da[k]; acc[4][k] += dc[k];
        }
        sb[0] += dl0; sb[1] += dl1; sb[2] += dv;
        *reinterpret_cast<__nv_bfloat162 *>(dHa + (size_t)b * D + j) = __floats2bfloat162_rn(da[0], da[1]);
        *reinterpret_cast<__nv_bfloat162 *>(dHc + (size_t)b * D + j) = __floats2bfloat162_rn(dc[0], dc[1]);
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) { atomicAdd(&s_red[a][j], acc[a][0]); atomicAdd(&s_red[a][j + 1], acc[a][1]); }
    if (lane == 0) { atomicAdd(&s_b2[0], sb[0]); atomicAdd(&s_b2[1], sb[1]); atomicAdd(&s_b2[2], sb[2]); }
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * HID; i += blockDim.x) {
        const int a = i / HID, c = i % HID;
        float *dst = a == 0 ? g_w2a + c : a == 1 ? g_w2a + HID + c : a == 2 ? g_w2c + c : a == 3 ? g_b1a + c : g_b1c + c;
        atomicAdd(dst, s_red[a][c]);
    }
    if (threadIdx.x == 0) { atomicAdd(g_b2a, s_b2[0]); atomicAdd(g_b2a + 1, s_b2[1]); atomicAdd(g_b2c, s_b2[2]); }
}

// PPO loss (agents/ppo.py:126-153) on the heads' outputs, forward and backward in two passes over the minibatch:
//   L = -mean(min(r A, clip(r, 1-eps, 1+eps) A)) + c_v max(mean((v - R)^2), mean((v_clip - R)^2)) - c_e mean(entropy)
// pass 1 accumulates the four sums; pass 2 (which needs the means to pick the value-loss branch torch.max picks)
// writes dL/dlogits and dL/dvalue.  acc: [0] sum -surrogate, [1] sum (v-R)^2, [2] sum (v_clip-R)^2, [3] sum entropy.
struct PpoLossArgs {
    const float *logits, *value, *old_logp, *adv, *ret, *old_value;
    const int64_t *action;
    int n;
    float eps, c_value, c_entropy;
};
__device__ __forceinline__ void ppo_sample(const PpoLossArgs &a, int b, float &lp0, float &lp1, float &ratio, float &adv, float &surr1,
                                           float &surr2, float &ent, float &v, float &vclip, float &ret) {
    const float l0 = a.logits[2 * b], l1 = a.logits[2 * b + 1];
    const float m = fmaxf(l0, l1), lse = m + logf(expf(l0 - m) + expf(l1 - m));
    lp0 = l0 - lse; lp1 = l1 - lse;
    const float p0 = expf(lp0), p1 = expf(lp1);
    ent = -(p0 * lp0 + p1 * lp1);
    const float logp = a.action[b] == 1 ? lp1 : lp0;
    ratio = expf(logp - a.old_logp[b]);
    adv = a.adv[b];
    surr1 = ratio * adv;
    surr2 = fminf(fmaxf(ratio, 1.0f - a.eps), 1.0f + a.eps) * adv;
    v = a.value[b]; ret = a.ret[b];
    const float ov = a.old_value[b];
    vclip = ov + fminf(fmaxf(v - ov, -a.eps), a.eps);
}
__global__ void __launch_bounds__(256) ppo_loss_sums_kernel(const PpoLossArgs a, float *__restrict__ acc) {
    __shared__ float s_red[4][8];
    float s[4] = {0, 0, 0, 0};
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.n; b += gridDim.x * blockDim.x) {
        float lp0, lp1, ratio, adv, s1, s2, ent, v, vc, r;
        ppo_sample(a, b, lp0, lp1, ratio, adv, s1, s2, ent, v, vc, r);
        s[0] -= fminf(s1, s2); s[1] += (v - r) * (v - r); s[2] += (vc - r) * (vc - r); s[3] += ent;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s[k] = warp_sum(s[k]);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
        atomicAdd(acc + threadIdx.x, t);
    }
}
__global__ void __launch_bounds__(256) ppo_loss_grad_kernel(const PpoLossArgs a, const float *__restrict__ acc, float *__restrict__ dlogits,
                                                            float *__restrict__ dvalue, float *__restrict__ stats) {
    const float inv_n = 1.0f / (float)a.n;
    const bool clipped_branch = acc[2] > acc[1];          // torch.max(mean1, mean2): the larger mean carries the gradient
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats) {   // the three numbers the reference logs (ppo.py:162-168)
        stats[0] = acc[0] * inv_n; stats[1] = fmaxf(acc[1], acc[2]) * inv_n; stats[2] = acc[3] * inv_n;
    }
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n) return;
    float lp0, lp1, ratio, adv, s1, s2, ent, v, vc, r;
    ppo_sample(a, b, lp0, lp1, ratio, adv, s1, s2, ent, v, vc, r);
    // d(-min(s1, s2))/d ratio: A where the unclipped term is the minimum or the clamp is inactive, else 0
    const bool inside = ratio >= 1.0f - a.eps && ratio <= 1.0f + a.eps;
    const float dlogp = (s1 < s2 || inside) ? -adv * ratio * inv_n : 0.0f;
    const float p0 = expf(lp0), p1 = expf(lp1);
    const bool a1 = a.action[b] == 1;
    const float ce = a.c_entropy * inv_n;                 // -c_e mean(H):  dH/dz_k = -p_k (log p_k + H)
    dlogits[2 * b] = dlogp * ((a1 ? 0.0f : 1.0f) - p0) + ce * p0 * (lp0 + ent);
    dlogits[2 * b + 1] = dlogp * ((a1 ? 1.0f : 0.0f) - p1) + ce * p1 * (lp1 + ent);
    const float ov = a.old_value[b];
    float dv;
    if (!clipped_branch) dv = 2.0f * (v - r);
    else dv = (v - ov >= -a.eps && v - ov <= a.eps) ? 2.0f * (vc - r) : 0.0f;
    dvalue[b] = a.c_value * dv * inv_n;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- host side

namespace {
struct LayerT { const bf16 *in_t, *in_q_t, *in_kv_t, *out_t, *l1_t, *l2_t; };          // transposed bf16 GEMM weights
struct LayerOff { size_t in_w, in_b, out_w, out_b, l1_w, l1_b, l2_w, l2_b, n1_w, n1_b, n2_w, n2_b; };
struct BlockOff { size_t pos, emb_w, emb_b; LayerOff layer[2]; };
struct FullAct { bf16 *QKV, *ATT, *XH1, *Y1, *Hf, *XH2, *Xout; float *rstd1, *rstd2; };   // inner layer, R rows
struct LastAct { bf16 *KV, *Q, *AL, *XH1, *Y1, *Hs, *XH2; float *rstd1, *rstd2; };        // last layer
struct HeadOff { size_t w1, b1, w2, b2; };
struct HeadT { bf16 *w1, *w1_t; bf16 *Z, *Hh, *dHh; };                                     // bf16 weights, feature / hidden / gradient rows
}  // namespace

struct uavtrain {
    int device = 0, max_samples = 0, sms = 0, n = 0;
    bf16 *arena = nullptr;                       // bf16 GEMM weights: row-major copies and transposes
    uavp::BlockW actor, critic;                  // fp32 members point into the caller's flat parameter buffer (set per forward)
    LayerT actor_t[1], critic_t[2];
    BlockOff actor_off, critic_off;
    HeadOff ahead_off, chead_off;
    HeadT ahead, chead;
    bool with_heads = false;                     // the last forward also ran the heads
    PrepJobs jobs;
    int n_jobs = 0;
    size_t job_src_off[kMaxPrepJobs];
    bf16 *Ea, *Ec;
    LastAct la, lc;
    FullAct fc;
    bf16 *T, *T2;                                // forward scratch (GEMM outputs feeding add+LN)
    bf16 *dS1, *dS2, *dH, *dQKV, *dXa, *dXb, *dQ, *tmp;   // backward scratch
    float *zeros = nullptr, *loss_acc = nullptr;
    bf16 *obs16 = nullptr;                       // [R,64] bf16 observations (zero-padded): B operand of the embedding weight gradient
    uint32_t *relu_mask = nullptr;               // [R][2 nets][4 words] ReLU activity bits of the two embeddings
    uint8_t *pad = nullptr;
    const float *params = nullptr;               // likewise (second head layers)
    void *gemm_ws = nullptr;
    std::vector<void *> allocs;
    std::string err;
};

namespace {
thread_local std::string g_terr;
int tfail(uavtrain *p, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (p) p->err = buf; else g_terr = buf;
    return code;
}
#define T_TRY(p, expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return tfail(p, -2, "%s failed: %s", #expr, cudaGetErrorString(e_));      \
    } while (0)

template <typename T>
cudaError_t talloc(uavtrain *p, T **ptr, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) { p->allocs.push_back(q); *ptr = static_cast<T *>(q); }
    return e;
}

// lays out the bf16 weight arena of one block and records the conversion jobs; returns the new flat offset
size_t map_train_block(uavtrain *p, uavp::BlockW &b, LayerT *lt, BlockOff &bo, int layers, size_t off, size_t &arena_off) {
    b.layers = layers;
    bo.pos = off; off += S * D;
    bo.emb_w = off; off += D * F;
    bo.emb_b = off; off += D;
    auto take = [&](size_t elems) { bf16 *q = p->arena + arena_off; arena_off += (elems + 63) / 64 * 64; return q; };
    auto job = [&](size_t src_off, bf16 *dst, bf16 *dst_t, int N, int K) {
        p->job_src_off[p->n_jobs] = src_off;
        p->jobs.job[p->n_jobs++] = PrepJob{nullptr, dst, dst_t, N, K};
    };
    for (int l = 0; l < layers; ++l) {
        uavp::LayerW &L = b.layer[l];
        LayerOff &o = bo.layer[l];
        bf16 *in_w = take(3 * D * D), *in_t = take(3 * D * D), *in_q_t = take(D * D), *in_kv_t = take(2 * D * D);
        bf16 *out_w = take(D * D), *out_t = take(D * D), *l1_w = take(FF * D), *l1_t = take(FF * D), *l2_w = take(D * FF), *l2_t = take(D * FF);
        L.in_w = in_w; L.out_w = out_w; L.l1_w = l1_w; L.l2_w = l2_w;
        L.in_wp = L.out_wp = L.l1_wp = L.l2_wp = nullptr;
        lt[l] = LayerT{in_t, in_q_t, in_kv_t, out_t, l1_t, l2_t};
        o.in_w = off; job(off, in_w, in_t, 3 * D, D); job(off, nullptr, in_q_t, D, D); job(off + D * D, nullptr, in_kv_t, 2 * D, D);
        off += 3 * D * D;
        o.in_b = off; off += 3 * D;
        o.out_w = off; job(off, out_w, out_t, D, D); off += D * D;
        o.out_b = off; off += D;
        o.l1_w = off; job(off, l1_w, l1_t, FF, D); off += FF * D;
        o.l1_b = off; off += FF;
        o.l2_w = off; job(off, l2_w, l2_t, D, FF); off += D * FF;
        o.l2_b = off; off += D;
        o.n1_w = off; off += D; o.n1_b = off; off += D; o.n2_w = off; off += D; o.n2_b = off; off += D;
    }
    return off;
}

void bind_params(uavp::BlockW &b, const BlockOff &bo, const float *w) {
    b.pos = w + bo.pos; b.emb_w = w + bo.emb_w; b.emb_b = w + bo.emb_b; b.emb_w2p = nullptr;
    for (int l = 0; l < b.layers; ++l) {
        uavp::LayerW &L = b.layer[l];
        const LayerOff &o = bo.layer[l];
        L.in_b = w + o.in_b; L.out_b = w + o.out_b; L.l1_b = w + o.l1_b; L.l2_b = w + o.l2_b;
        L.n1_w = w + o.n1_w; L.n1_b = w + o.n1_b; L.n2_w = w + o.n2_w; L.n2_b = w + o.n2_b;
    }
}

struct TCtx { uavtrain *p; cudaStream_t s; int rc; };

void gemm(TCtx &c, const bf16 *A, int64_t lda, const bf16 *W, const float *bias, bf16 *Dst, int M, int N, int K, int relu) {
    if (c.rc) return;
    const int r = uavp::gemm_bias_act(A, lda, W, bias ? bias : c.p->zeros, Dst, M, N, K, relu, c.p->gemm_ws, uavp::gemm_workspace_bytes(), c.s);
    if (r) c.rc = tfail(c.p, -2, "tcgen05 GEMM (M=%d N=%d K=%d) failed with %d", M, N, K, r);
}
// out = LayerNorm(x + A W^T + bias) fused into the dense kernel's epilogue (policy_dense.cu), keeping x^ and 1/sigma
void gemm_ln(TCtx &c, const bf16 *A, int64_t lda, const bf16 *W, const float *bias, const bf16 *x, int64_t xs, const float *g,
             const float *b, int rows, int K, bf16 *out16, bf16 *xhat, float *rstd) {
    if (c.rc) return;
    const int r = uavp::gemm_add_ln(A, lda, W, bias, x, xs, g, b, out16, xhat, rstd, rows, K, c.s);
    if (r) c.rc = tfail(c.p, -2, "tcgen05 GEMM + LayerNorm (M=%d K=%d) failed with %d", rows, K, r);
}
void wgrad(TCtx &c, const bf16 *dY, int64_t ld_dy, const bf16 *X, int64_t ld_x, int rows, int Nout, int Kin, float *dW,
           int nout_valid = -1, float *dbias = nullptr, int out_ld = 0, int kin_valid = 0) {
    if (c.rc) return;
    const int r = uavp::wgrad(dY, ld_dy, X, ld_x, rows, Nout, Kin, dW, nout_valid < 0 ? Nout : nout_valid, dbias, c.p->sms, c.s, out_ld,
                              kin_valid);
    if (r) c.rc = tfail(c.p, -2, "weight-gradient kernel (rows=%d Nout=%d Kin=%d) failed with %d", rows, Nout, Kin, r);
}
void add_ln(TCtx &c, const bf16 *x, int64_t xs, const bf16 *y, const float *g, const float *b, int rows, bf16 *out16, float *out32,
            int64_t out32_stride, bf16 *xhat, float *rstd) {
    if (c.rc) return;
    add_ln_train_kernel<<<(rows * 16 + 255) / 256, 256, 0, c.s>>>(x, xs, y, g, b, rows, out16, out32, out32_stride, xhat, rstd);
}
// launch geometry of the "a warp walks consecutive rows" kernels: about 8 CTAs of 8 warps per SM
inline void row_grid(const uavtrain *p, int rows, int &grid, int &rpw) {
    const int warps = p->sms * 8 * 8;
    rpw = (rows + warps - 1) / warps;
    if (rpw < 4) rpw = 4;
    grid = ((rows + rpw - 1) / rpw + 7) / 8;
}
void ln_bwd(TCtx &c, const float *dy32, int64_t dy_stride, const bf16 *dy16, const bf16 *add, const bf16 *xhat, const float *rstd,
            const float *gamma, int rows, bf16 *dz, float *g_gamma, float *g_beta, float *g_bias) {
    if (c.rc) return;
    int grid, rpw;
    row_grid(c.p, rows, grid, rpw);
    if (dy32) ln_bwd_kernel<true><<<grid, 256, 0, c.s>>>(dy32, dy_stride, add, xhat, rstd, gamma, rows, rpw, dz, g_gamma, g_beta, g_bias);
    else ln_bwd_kernel<false><<<grid, 256, 0, c.s>>>(dy16, D, add, xhat, rstd, gamma, rows, rpw, dz, g_gamma, g_beta, g_bias);
}

// ---- forward -----------------------------------------------------------------------------------------------
void last_layer_fwd(TCtx &c, const uavp::LayerW &L, const bf16 *X, int n, LastAct &A, float *feat, int64_t feat_stride, bf16 *feat16) {
    uavtrain *p = c.p;
    const int R = n * S;
    const bf16 *Xl = X + (S - 1) * D;
    gemm(c, X, D, L.in_w + D * D, L.in_b + D, A.KV, R, 2 * D, D, 0);
    gemm(c, Xl, (int64_t)S * D, L.in_w, L.in_b, A.Q, n, D, D, 0);
    if (!c.rc) attn_last_kernel<<<(n * H + 255) / 256, 256, 0, c.s>>>(A.Q, A.KV, p->pad, n, A.AL);
    gemm_ln(c, A.AL, D, L.out_w, L.out_b, Xl, (int64_t)S * D, L.n1_w, L.n1_b, n, D, A.Y1, A.XH1, A.rstd1);
    gemm(c, A.Y1, D, L.l1_w, L.l1_b, A.Hs, n, FF, D, 1);
    gemm(c, A.Hs, FF, L.l2_w, L.l2_b, p->T2, n, D, FF, 0);
    add_ln(c, A.Y1, D, p->T2, L.n2_w, L.n2_b, n, feat16, feat, feat_stride, A.XH2, A.rstd2);
}
void full_layer_fwd(TCtx &c, const uavp::LayerW &L, const bf16 *X, int n, FullAct &A) {
    uavtrain *p = c.p;
    const int R = n * S;
    gemm(c, X, D, L.in_w, L.in_b, A.QKV, R, 3 * D, D, 0);
    if (!c.rc) attn_full_kernel<<<std::min((n + 7) / 8, p->sms * 16), 256, 0, c.s>>>(A.QKV, p->pad, n, A.ATT);
    gemm_ln(c, A.ATT, D, L.out_w, L.out_b, X, D, L.n1_w, L.n1_b, R, D, A.Y1, A.XH1, A.rstd1);
    gemm(c, A.Y1, D, L.l1_w, L.l1_b, A.Hf, R, FF, D, 1);
    gemm_ln(c, A.Hf, FF, L.l2_w, L.l2_b, A.Y1, D, L.n2_w, L.n2_b, R, FF, A.Xout, A.XH2, A.rstd2);
}

// ---- backward ----------------------------------------------------------------------------------------------
// shared tail of both layer kinds: from the gradient of the layer output down to dS1 (gradient of x + attn-proj)
void ffn_ln_bwd(TCtx &c, const uavp::LayerW &L, const LayerT &T, const LayerOff &o, float *g, int rows, const float *dz32,
                int64_t dz_stride, const bf16 *dz16, const bf16 *XH2, const float *rstd2, const bf16 *Hact, const bf16 *Y1,
                const bf16 *XH1, const float *rstd1) {
    uavtrain *p = c.p;
    ln_bwd(c, dz32, dz_stride, dz16, nullptr, XH2, rstd2, L.n2_w, rows, p->dS2, g + o.n2_w, g + o.n2_b, g + o.l2_b);
    wgrad(c, p->dS2, D, Hact, FF, rows, D, FF, g + o.l2_w);
    if (!c.rc) {                                              // dH = (dS2 W2) masked by the forward's ReLU, in one GEMM
        const int r = uavp::gemm_drelu(p->dS2, D, T.l2_t, Hact, FF, p->dH, rows, FF, D, p->gemm_ws, uavp::gemm_workspace_bytes(), c.s);
        if (r) c.rc = tfail(p, -2, "tcgen05 GEMM with ReLU-backward epilogue (M=%d) failed with %d", rows, r);
    }
    wgrad(c, p->dH, FF, Y1, D, rows, FF, D, g + o.l1_w, -1, g + o.l1_b);                  // + bias gradient dH^T 1
    gemm(c, p->dH, FF, T.l1_t, nullptr, p->tmp, rows, D, FF, 0);
    ln_bwd(c, nullptr, 0, p->tmp, p->dS2, XH1, rstd1, L.n1_w, rows, p->dS1, g + o.n1_w, g + o.n1_b, g + o.out_b);
}
// dX [R,128] <- gradient w.r.t. the layer input (all five tokens)
void last_layer_bwd(TCtx &c, const uavp::LayerW &L, const LayerT &T, const LayerOff &o, float *g, const bf16 *X, int n, LastAct &A,
                    const float *dfeat, int64_t dfeat_stride, const bf16 *dfeat16, bf16 *dX) {
    uavtrain *p = c.p;
    const int R = n * S;
    ffn_ln_bwd(c, L, T, o, g, n, dfeat, dfeat_stride, dfeat16, A.XH2, A.rstd2, A.Hs, A.Y1, A.XH1, A.rstd1);
    wgrad(c, p->dS1, D, A.AL, D, n, D, D, g + o.out_w);
    gemm(c, p->dS1, D, T.out_t, nullptr, p->tmp, n, D, D, 0);                       // dAL
    if (!c.rc) {
        AttnBwdArgs a{A.Q, A.KV, A.KV + D, p->tmp, p->dQ, p->dQKV, p->dQKV + D, D, 2 * D, p->pad, n};
        attn_bwd_last_kernel<<<(n * H + 255) / 256, 256, 0, c.s>>>(a);
    }
    wgrad(c, p->dQ, D, X + (S - 1) * D, (int64_t)S * D, n, D, D, g + o.in_w, -1, g + o.in_b);          // + bias gradients
    wgrad(c, p->dQKV, 2 * D, X, D, R, 2 * D, D, g + o.in_w + D * D, -1, g + o.in_b + D);
    gemm(c, p->dQKV, 2 * D, T.in_kv_t, nullptr, dX, R, D, 2 * D, 0);
    gemm(c, p->dQ, D, T.in_q_t, nullptr, p->tmp, n, D, D, 0);
    if (!c.rc) scatter_add_last_kernel<<<(n * (D / 4) + 255) / 256, 256, 0, c.s>>>(dX, p->tmp, p->dS1, n);
}
// returns the two addends of the gradient w.r.t. the layer input: dXg (GEMM branch) and p->dS1 (residual branch)
void full_layer_bwd(TCtx &c, const uavp::LayerW &L, const LayerT &T, const LayerOff &o, float *g, const bf16 *X, int n, FullAct &A,
                    const bf16 *dY, bf16 *dXg) {
    uavtrain *p = c.p;
    const int R = n * S;
    ffn_ln_bwd(c, L, T, o, g, R, nullptr, 0, dY, A.XH2, A.rstd2, A.Hf, A.Y1, A.XH1, A.rstd1);
    wgrad(c, p->dS1, D, A.ATT, D, R, D, D, g + o.out_w);
    gemm(c, p->dS1, D, T.out_t, nullptr, p->tmp, R, D, D, 0);                       // dATT
    if (!c.rc) {
        AttnBwdArgs a{A.QKV, A.QKV + D, A.QKV + 2 * D, p->tmp, p->dQKV, p->dQKV + D, p->dQKV + 2 * D, 3 * D, 3 * D, p->pad, n};
        attn_bwd_full_kernel<<<std::min((n + 7) / 8, p->sms * 16), 256, 0, c.s>>>(a);
    }
    wgrad(c, p->dQKV, 3 * D, X, D, R, 3 * D, D, g + o.in_w, -1, g + o.in_b);                           // + bias gradient
    gemm(c, p->dQKV, 3 * D, T.in_t, nullptr, dXg, R, D, 3 * D, 0);
}
void embed_bwd(TCtx &c, const BlockOff &bo, int net, float *g, int n, const bf16 *g1, const bf16 *g2) {
    if (c.rc) return;
    uavtrain *p = c.p;
    int grid, rpw;
    row_grid(p, n, grid, rpw);                                  // (here a "row" is a window of 5 token rows)
    embed_mask_kernel<<<grid, 256, 0, c.s>>>(g1, g2, p->relu_mask + net * 4, n, rpw, p->tmp, g + bo.pos);
    wgrad(c, p->tmp, D, p->obs16, 64, n * S, D, 64, g + bo.emb_w, -1, g + bo.emb_b, F, F);
}
}  // namespace

extern "C" const char *uavtrain_last_error(const uavtrain_t *p) { return p ? p->err.c_str() : g_terr.c_str(); }

extern "C" int uavtrain_create(int32_t device, int32_t max_samples, uavtrain_t **out) {
    if (!out) return tfail(nullptr, -1, "uavtrain_create: out is NULL");
    *out = nullptr;
    if (max_samples <= 0) return tfail(nullptr, -1, "uavtrain_create: max_samples must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return tfail(nullptr, -2, "uavtrain_create: no CUDA device; there is no CPU fallback");
    if (device < 0 || device >= ndev) return tfail(nullptr, -1, "uavtrain_create: device %d out of range", device);
    uavtrain *p = new (std::nothrow) uavtrain();
    if (!p) return tfail(nullptr, -3, "out of host memory");
    p->device = device; p->max_samples = max_samples;
    auto bail = [&](int rc) { g_terr = p->err; for (void *q : p->allocs) cudaFree(q); delete p; return rc; };
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device);
    const size_t n = (size_t)max_samples, R = n * S;
    if (e == cudaSuccess) e = talloc(p, &p->arena, (size_t)3 * (2 * 3 * D * D + 3 * D * D + 2 * D * D + 4 * FF * D) + 4 * HID * D + 64 * 64);
    bf16 **rowsR128[] = {&p->Ea, &p->Ec, &p->fc.ATT, &p->fc.XH1, &p->fc.Y1, &p->fc.XH2, &p->fc.Xout, &p->T, &p->dS1, &p->dS2, &p->dXa, &p->dXb, &p->tmp};
    for (auto b : rowsR128) if (e == cudaSuccess) e = talloc(p, b, R * D);
    bf16 **rowsR256[] = {&p->la.KV, &p->lc.KV, &p->fc.Hf, &p->dH};
    for (auto b : rowsR256) if (e == cudaSuccess) e = talloc(p, b, R * FF);
    bf16 **rowsR384[] = {&p->fc.QKV, &p->dQKV};
    for (auto b : rowsR384) if (e == cudaSuccess) e = talloc(p, b, R * 3 * D);
    for (LastAct *A : {&p->la, &p->lc}) {
        bf16 **rowsN128[] = {&A->Q, &A->AL, &A->XH1, &A->Y1, &A->XH2};
        for (auto b : rowsN128) if (e == cudaSuccess) e = talloc(p, b, n * D);
        if (e == cudaSuccess) e = talloc(p, &A->Hs, n * FF);
        if (e == cudaSuccess) e = talloc(p, &A->rstd1, n);
        if (e == cudaSuccess) e = talloc(p, &A->rstd2, n);
    }
    for (HeadT *Hd : {&p->ahead, &p->chead}) {
        if (e == cudaSuccess) e = talloc(p, &Hd->Z, n * D);
        if (e == cudaSuccess) e = talloc(p, &Hd->Hh, n * HID);
        if (e == cudaSuccess) e = talloc(p, &Hd->dHh, n * D);
        if (e == cudaSuccess) e = cudaMemset(Hd->dHh, 0, n * D * sizeof(bf16));      // columns 64.. are never written again
    }
    if (e == cudaSuccess) e = talloc(p, &p->T2, n * D);
    if (e == cudaSuccess) e = talloc(p, &p->dQ, n * D);
    if (e == cudaSuccess) e = talloc(p, &p->fc.rstd1, R);
    if (e == cudaSuccess) e = talloc(p, &p->fc.rstd2, R);
    if (e == cudaSuccess) e = talloc(p, &p->pad, R);
    if (e == cudaSuccess) e = talloc(p, &p->loss_acc, (size_t)4);
    if (e == cudaSuccess) e = talloc(p, &p->obs16, R * 64);
    if (e == cudaSuccess) e = talloc(p, &p->relu_mask, R * 8);
    if (e == cudaSuccess) e = talloc(p, &p->zeros, (size_t)3 * D);
    if (e == cudaSuccess) e = cudaMemset(p->zeros, 0, 3 * D * sizeof(float));
    if (e == cudaSuccess) { void *ws = nullptr; e = cudaMalloc(&ws, uavp::gemm_workspace_bytes()); if (e == cudaSuccess) { p->allocs.push_back(ws); p->gemm_ws = ws; } }
    if (e != cudaSuccess) { tfail(p, -2, "uavtrain_create: %s", cudaGetErrorString(e)); return bail(-2); }
    if (uavp::wgrad_prepare() != 0) { tfail(p, -2, "uavtrain_create: cannot reserve shared memory for the weight-gradient kernel"); return bail(-2); }
    size_t arena_off = 0;
    auto map_head = [&](HeadT &Hd, HeadOff &ho, int outs, size_t off) {
        Hd.w1 = p->arena + arena_off; arena_off += HID * D;
        Hd.w1_t = p->arena + arena_off; arena_off += HID * D;
        ho.w1 = off;
        p->job_src_off[p->n_jobs] = off;
        p->jobs.job[p->n_jobs++] = PrepJob{nullptr, Hd.w1, Hd.w1_t, HID, D};
        off += HID * D;
        ho.b1 = off; off += HID;
        ho.w2 = off; off += (size_t)outs * HID;
        ho.b2 = off; off += outs;
        return off;
    };
    size_t off = map_train_block(p, p->actor, p->actor_t, p->actor_off, 1, 0, arena_off);
    off = map_head(p->ahead, p->ahead_off, NACT, off);
    off = map_train_block(p, p->critic, p->critic_t, p->critic_off, 2, off, arena_off);
    off = map_head(p->chead, p->chead_off, 1, off);
    if (off != (size_t)UAVPOLICY_NUM_PARAMS || p->n_jobs > kMaxPrepJobs) { tfail(p, -1, "internal: parameter layout mismatch"); return bail(-1); }
    *out = p;
    return 0;
}

extern "C" int uavtrain_destroy(uavtrain_t *p) {
    if (!p) return 0;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (void *q : p->allocs) cudaFree(q);
    delete p;
    return 0;
}

namespace {
int forward_impl(uavtrain *p, const float *w, const float *d_obs, int n, float *d_feat, float *d_logits, float *d_value, void *stream) {
    if (n <= 0 || n > p->max_samples) return tfail(p, -1, "uavtrain_forward: n=%d outside (0, %d]", n, p->max_samples);
    T_TRY(p, cudaSetDevice(p->device));
    TCtx c{p, (cudaStream_t)stream, 0};
    const bool heads = d_logits != nullptr;
    p->n = 0;
    bind_params(p->actor, p->actor_off, w);
    bind_params(p->critic, p->critic_off, w);
    for (int j = 0; j < p->n_jobs; ++j) p->jobs.job[j].src = w + p->job_src_off[j];
    prep_weights_kernel<<<dim3(24, p->n_jobs), 256, 0, c.s>>>(p->jobs);
    const int R = n * S;
    embed_kernel<<<(R + kEmbTok - 1) / kEmbTok, kEmbThreads, 0, c.s>>>(d_obs, R, p->actor, p->critic, p->Ea, p->Ec, p->pad, p->obs16, p->relu_mask);
    last_layer_fwd(c, p->actor.layer[0], p->Ea, n, p->la, d_feat, 2 * D, heads ? p->ahead.Z : nullptr);
    full_layer_fwd(c, p->critic.layer[0], p->Ec, n, p->fc);
    last_layer_fwd(c, p->critic.layer[1], p->fc.Xout, n, p->lc, d_feat ? d_feat + D : nullptr, 2 * D, heads ? p->chead.Z : nullptr);
    if (heads) {
        gemm(c, p->ahead.Z, D, p->ahead.w1, w + p->ahead_off.b1, p->ahead.Hh, n, HID, D, 1);
        gemm(c, p->chead.Z, D, p->chead.w1, w + p->chead_off.b1, p->chead.Hh, n, HID, D, 1);
        if (!c.rc)
            head_out_kernel<<<(n + 255) / 256, 256, 0, c.s>>>(p->ahead.Hh, p->chead.Hh, w + p->ahead_off.w2, w + p->ahead_off.b2,
                                                              w + p->chead_off.w2, w + p->chead_off.b2, n, d_logits, d_value);
    }
    if (c.rc) return c.rc;
    T_TRY(p, cudaGetLastError());
    p->n = n;
    p->with_heads = heads;
    p->params = w;
    return 0;
}

int backward_impl(uavtrain *p, const float *d_dfeat, const float *d_dlogits, const float *d_dvalue, float *g, void *stream) {
    T_TRY(p, cudaSetDevice(p->device));
    TCtx c{p, (cudaStream_t)stream, 0};
    const int n = p->n;
    const bool heads = d_dlogits != nullptr;
    T_TRY(p, cudaMemsetAsync(g, 0, (size_t)UAVPOLICY_NUM_PARAMS * sizeof(float), c.s));
    if (heads) {
        const float *w = p->params;
        const int warps = p->sms * 8 * 4, spw = max(1, (n + warps - 1) / warps);
        head_bwd_kernel<<<((n + spw - 1) / spw + 7) / 8, 256, 0, c.s>>>(
            p->ahead.Hh, p->chead.Hh, w + p->ahead_off.w2, w + p->chead_off.w2, d_dlogits, d_dvalue, n, spw, p->ahead.dHh, p->chead.dHh,
            g + p->ahead_off.w2, g + p->ahead_off.b2, g + p->ahead_off.b1, g + p->chead_off.w2, g + p->chead_off.b2, g + p->chead_off.b1);
        wgrad(c, p->ahead.dHh, D, p->ahead.Z, D, n, D, D, g + p->ahead_off.w1, HID);
        wgrad(c, p->chead.dHh, D, p->chead.Z, D, n, D, D, g + p->chead_off.w1, HID);
        gemm(c, p->ahead.dHh, D, p->ahead.w1_t, nullptr, p->ahead.Z, n, D, HID, 0);      // dZ overwrites Z (no longer needed)
        gemm(c, p->chead.dHh, D, p->chead.w1_t, nullptr, p->chead.Z, n, D, HID, 0);
    }
    // actor: one (last) layer on the embedding
    last_layer_bwd(c, p->actor.layer[0], p->actor_t[0], p->actor_off.layer[0], g, p->Ea, n, p->la, heads ? nullptr : d_dfeat, 2 * D,
                   heads ? p->ahead.Z : nullptr, p->dXa);
    embed_bwd(c, p->actor_off, 0, g, n, p->dXa, nullptr);
    // critic: last layer, inner layer, embedding
    last_layer_bwd(c, p->critic.layer[1], p->critic_t[1], p->critic_off.layer[1], g, p->fc.Xout, n, p->lc,
                   heads ? nullptr : d_dfeat + D, 2 * D, heads ? p->chead.Z : nullptr, p->dXa);
    full_layer_bwd(c, p->critic.layer[0], p->critic_t[0], p->critic_off.layer[0], g, p->Ec, n, p->fc, p->dXa, p->dXb);
    embed_bwd(c, p->critic_off, 1, g, n, p->dXb, p->dS1);
    if (c.rc) return c.rc;
    T_TRY(p, cudaGetLastError());
    return 0;
}
}  // namespace

extern "C" int uavtrain_forward(uavtrain_t *p, const float *d_flat_params, const float *d_obs, int32_t n, float *d_feat, void *stream) {
    if (!p) return -1;
    if (!d_flat_params || !d_obs || !d_feat) return tfail(p, -1, "uavtrain_forward: NULL argument");
    return forward_impl(p, d_flat_params, d_obs, n, d_feat, nullptr, nullptr, stream);
}

extern "C" int uavtrain_backward(uavtrain_t *p, const float *d_dfeat, float *d_flat_grad, void *stream) {
    if (!p) return -1;
    if (!d_dfeat || !d_flat_grad) return tfail(p, -1, "uavtrain_backward: NULL argument");
    if (p->n <= 0 || p->with_heads) return tfail(p, -4, "uavtrain_backward without a preceding uavtrain_forward");
    return backward_impl(p, d_dfeat, nullptr, nullptr, d_flat_grad, stream);
}

extern "C" int uavtrain_forward_heads(uavtrain_t *p, const float *d_flat_params, const float *d_obs, int32_t n, float *d_logits,
                                      float *d_value, void *stream) {
    if (!p) return -1;
    if (!d_flat_params || !d_obs || !d_logits || !d_value) return tfail(p, -1, "uavtrain_forward_heads: NULL argument");
    return forward_impl(p, d_flat_params, d_obs, n, nullptr, d_logits, d_value, stream);
}

extern "C" int uavtrain_backward_heads(uavtrain_t *p, const float *d_dlogits, const float *d_dvalue, float *d_flat_grad, void *stream) {
    if (!p) return -1;
    if (!d_dlogits || !d_dvalue || !d_flat_grad) return tfail(p, -1, "uavtrain_backward_heads: NULL argument");
    if (p->n <= 0 || !p->with_heads) return tfail(p, -4, "uavtrain_backward_heads without a preceding uavtrain_forward_heads");
    return backward_impl(p, nullptr, d_dlogits, d_dvalue, d_flat_grad, stream);
}

extern "C" int uavtrain_ppo_loss(uavtrain_t *p, const float *d_logits, const float *d_value, const int64_t *d_action, const float *d_old_logp,
                                 const float *d_adv, const float *d_ret, const float *d_old_value, int32_t n, float eps_clip, float c_value,
                                 float c_entropy, float *d_dlogits, float *d_dvalue, float *d_stats, void *stream) {
    if (!p) return -1;
    if (!d_logits || !d_value || !d_action || !d_old_logp || !d_adv || !d_ret || !d_old_value || !d_dlogits || !d_dvalue || n <= 0)
        return tfail(p, -1, "uavtrain_ppo_loss: NULL argument or n <= 0");
    T_TRY(p, cudaSetDevice(p->device));
    cudaStream_t s = (cudaStream_t)stream;
    PpoLossArgs a{d_logits, d_value, d_old_logp, d_adv, d_ret, d_old_value, d_action, n, eps_clip, c_value, c_entropy};
    T_TRY(p, cudaMemsetAsync(p->loss_acc, 0, 4 * sizeof(float), s));
    ppo_loss_sums_kernel<<<min((n + 255) / 256, p->sms * 4), 256, 0, s>>>(a, p->loss_acc);
    ppo_loss_grad_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, p->loss_acc, d_dlogits, d_dvalue, d_stats);
    T_TRY(p, cudaGetLastError());
    return 0;
}
