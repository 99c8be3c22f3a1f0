// policy_wgrad.cu - weight gradients of the PPO update on tcgen05: dW[Nout, Kin] += dY[rows, Nout]^T X[rows, Kin].
//
// The contraction runs over the batch rows (hundreds of thousands) while the output is a single small matrix, so this
// is a split-K problem: every CTA owns a strided set of 64/128-row slabs, streams the dY and X tiles of each slab into
// shared memory (cp.async, 3-4 stage ring) and accumulates its partial dW in TMEM across ALL its slabs - one
// tcgen05.mma chain, no intermediate traffic - then adds the partial to the fp32 gradient with coalesced RED.ADD.
// Both operands are read MN-major straight from the row-major activations (tcgen05_util.cuh: instr_desc_bf16_mn): no
// transposed copy of dY or X is ever made.  blockIdx.y selects a 128-column block of dY (= 128 rows of dW).
#include "policy_weights.cuh"
#include "tcgen05_util.cuh"

namespace uavp {
namespace {

constexpr int kWgThreads = 512;

// 64-row slabs; KIN = 128: 6 stages of 32 KB, KIN = 256: 4 stages of 48 KB (192 KB either way).  Loads run
// kStages - 2 slabs ahead, so the stage being refilled was read by MMAs issued a whole iteration ago.
template <int KIN>
struct WgCfg {
    static constexpr int kSlab = 64;                         // rows per slab = 4 k-steps of 16
    static constexpr int kStages = KIN == 128 ? 6 : 4;
    static constexpr int kAhead = kStages - 2;
    static constexpr int kA = kSlab * 128 * 2;               // dY sub-tile [kSlab rows x 128 cols] bf16
    static constexpr int kB = kSlab * KIN * 2;               // X tile [kSlab rows x KIN] bf16
    static constexpr int kStage = kA + kB;
    static constexpr int kOnes = kSlab * 16 * 2;             // an all-ones [kSlab x 16] operand: dY^T 1 = the bias gradient
    static constexpr int kTotal = kStages * kStage + kOnes + 64;
    static_assert(128 * (KIN + 1) * 4 <= kStages * kStage, "the epilogue transposes through the operand buffers");
};

template <int KIN>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __nv_bfloat16 *__restrict__ dY, int64_t ld_dy,
                                                              const __nv_bfloat16 *__restrict__ X, int64_t ld_x, int rows,
                                                              float *__restrict__ dW, int nout_valid, float *__restrict__ dbias) {
    using SM = WgCfg<KIN>;
    constexpr int kSlab = SM::kSlab, kStages = SM::kStages;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *ones = smem + kStages * SM::kStage;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(ones + SM::kOnes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ones + SM::kOnes + 48);
    constexpr int kCols = KIN == 128 ? 256 : KIN;           // TMEM columns: KIN accumulators (+ 16 for the bias block)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slabs = (rows + kSlab - 1) / kSlab;
    if ((int)blockIdx.x >= slabs) return;
    const int cnt = (slabs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // slabs of this CTA
    const int nb = blockIdx.y;

    if (warp == 0) tc::tmem_alloc(tmem_slot, kCols);
    if (dbias) {                                             // (any operand layout of all ones is all ones)
        for (int i = tid; i < SM::kOnes / 4; i += kWgThreads) reinterpret_cast<uint32_t *>(ones)[i] = 0x3F803F80u;
        tc::fence_async_smem();
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) tc::mbar_init(&mbar[s], 1);
        tc::fence_mbar_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t idesc = tc::instr_desc_bf16_mn(128, KIN);

    auto issue_loads = [&](int k) {                          // k-th slab of this CTA -> stage k % kStages
        unsigned char *a = smem + (k % kStages) * SM::kStage, *b = a + SM::kA;
        const int r0 = ((int)blockIdx.x + k * (int)gridDim.x) * kSlab, valid = min(kSlab, rows - r0);
        tc::load_canon_async_zfill(a, dY + (int64_t)r0 * ld_dy + nb * 128, kSlab, 128, ld_dy, valid, tid, kWgThreads);
        tc::load_canon_async_zfill(b, X + (int64_t)r0 * ld_x, kSlab, KIN, ld_x, valid, tid, kWgThreads);
    };

    constexpr int kAhead = SM::kAhead;
    for (int k = 0; k < kAhead; ++k) {
        if (k < cnt) issue_loads(k);
        tc::cp_async_commit();
    }
    for (int it = 0; it < cnt; ++it) {
        const int pf = it + kAhead;
        if (pf < cnt) {                                      // refill the stage the MMAs of slab it-2 read
            if (it >= 2) tc::mbar_wait(&mbar[(it - 2) % kStages], (uint32_t)(((it - 2) / kStages) & 1));
            issue_loads(pf);
        }
        tc::cp_async_commit();
        tc::cp_async_wait_group<kAhead>();                   // slab `it` has landed
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            const uint32_t a = tc::smem_u32(smem + (it % kStages) * SM::kStage), b = a + SM::kA;
#pragma unroll
            for (int j = 0; j < kSlab / 16; ++j) {
                const uint64_t ad = tc::smem_desc(a + j * 2 * (128 * 16), 128 * 16, 128);
                const uint64_t bd = tc::smem_desc(b + j * 2 * (KIN * 16), KIN * 16, 128);
                tc::mma_bf16(tmem, ad, bd, idesc, (it > 0 || j > 0) ? 1u : 0u);
                if (KIN == 128 && dbias)
                    tc::mma_bf16(tmem + KIN, ad, tc::smem_desc(tc::smem_u32(ones), 16 * 16, 128), tc::instr_desc_bf16_mn(128, 16),
                                 (it > 0 || j > 0) ? 1u : 0u);
            }
            tc::mma_commit(&mbar[it % kStages]);
        }
    }
    tc::mbar_wait(&mbar[(cnt - 1) % kStages], (uint32_t)(((cnt - 1) / kStages) & 1));   // commits complete in order
    tc::tc_fence_after();
    __syncthreads();

    // epilogue: TMEM lane = row of this dW block, columns = KIN.  Transpose through shared memory so that a warp adds
    // 32 consecutive floats of one dW row per instruction.
    float *stage_f = reinterpret_cast<float *>(smem);
    constexpr int kParts = kWgThreads / 128;                // column ranges: one per group of 4 warps
    const int lrow = (warp & 3) * 32 + lane, part = warp >> 2;
#pragma unroll 1
    for (int c = part * (KIN / kParts); c < (part + 1) * (KIN / kParts); c += 32) {
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) stage_f[lrow * (KIN + 1) + c + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    float *out = dW + (size_t)nb * 128 * KIN;
    const int rows_out = min(128, nout_valid - nb * 128);       // dW rows that exist (a 64-wide dY is padded to 128 columns)
    for (int i = tid; i < rows_out * KIN; i += kWgThreads) {
        const int r = i / KIN, c = i % KIN;
        atomicAdd(out + i, stage_f[r * (KIN + 1) + c]);
    }
    if (KIN == 128 && dbias && warp < 4) {                   // column KIN of the accumulator block: sum over rows of dY
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + KIN, v);
        if (lrow < rows_out) atomicAdd(dbias + nb * 128 + lrow, v[0]);
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_free(tmem, kCols); }
}

}  // namespace

int wgrad_prepare() {
    if (cudaFuncSetAttribute(wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<128>::kTotal) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<256>::kTotal) != cudaSuccess) return -1;
    return 0;
}

// dW[Nout, Kin] (fp32, dense) += dY[rows, Nout]^T X[rows, Kin];  Nout % 128 == 0, Kin in {128, 256}, ld % 8 == 0.
// Only the first nout_valid rows of dW are written (dY columns beyond that are padding).  dbias (optional, Kin = 128
// only): dbias[Nout] += column sums of dY, computed by the tensor cores as dY^T 1 alongside the main product.
int wgrad(const __nv_bfloat16 *dY, int64_t ld_dy, const __nv_bfloat16 *X, int64_t ld_x, int rows, int Nout, int Kin, float *dW,
          int nout_valid, float *dbias, int num_sms, cudaStream_t stream) {
    if (rows <= 0 || Nout % 128 || (Kin != 128 && Kin != 256) || ld_dy % 8 || ld_x % 8 || (dbias && Kin != 128)) return -1;
    const int slab = Kin == 128 ? WgCfg<128>::kSlab : WgCfg<256>::kSlab;
    const int gy = Nout / 128, slabs = (rows + slab - 1) / slab;
    const int gx = max(1, min(slabs, num_sms / gy));
    if (Kin == 128) wgrad_kernel<128><<<dim3(gx, gy), kWgThreads, WgCfg<128>::kTotal, stream>>>(dY, ld_dy, X, ld_x, rows, dW, nout_valid, dbias);
    else wgrad_kernel<256><<<dim3(gx, gy), kWgThreads, WgCfg<256>::kTotal, stream>>>(dY, ld_dy, X, ld_x, rows, dW, nout_valid, dbias);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace uavp

// self-test hook: one weight-gradient product on caller-provided device buffers
extern "C" int uavpolicy_selftest_wgrad(const void *d_dy, int64_t ld_dy, const void *d_x, int64_t ld_x, int32_t rows, int32_t n_out,
                                        int32_t k_in, float *d_dw, float *d_dbias, void *stream) {
    if (uavp::wgrad_prepare()) return -2;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -2;
    return uavp::wgrad(static_cast<const __nv_bfloat16 *>(d_dy), ld_dy, static_cast<const __nv_bfloat16 *>(d_x), ld_x, rows, n_out,
                       k_in, d_dw, n_out, d_dbias, sms, (cudaStream_t)stream);
}
