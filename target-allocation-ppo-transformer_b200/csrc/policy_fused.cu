// policy_fused.cu - hand-written tcgen05 kernels of the policy forward (sm_100a).  See tcgen05_util.cuh.
#include "uavpolicy_b200.h"

#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "tcgen05_util.cuh"

namespace {
using namespace tc;

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

extern "C" int uavpolicy_selftest_gemm_tile(const void *d_A, const void *d_W, float *d_D, int32_t N, int32_t K, void *stream) {
    if (!d_A || !d_W || !d_D || N <= 0 || N > 384 || N % 16 || (K != 128 && K != 256)) return -1;
    const size_t smem = (size_t)(128 + N) * K * 2;
    if (cudaFuncSetAttribute(gemm_tile_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    gemm_tile_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16 *)d_A, (const __nv_bfloat16 *)d_W, d_D, N, K);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
