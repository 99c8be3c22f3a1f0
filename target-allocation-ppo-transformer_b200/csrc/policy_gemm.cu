// policy_gemm.cu - tcgen05 GEMM with fused bias / ReLU epilogue for the policy forward (sm_100a).
// One CTA computes a 128 x 128 output tile: TMA loads (128B swizzle) of A and W K-slices into a multi-stage shared
// memory ring, a single elected thread issues tcgen05.mma into a TMEM accumulator, the epilogue warps read it back
// with tcgen05.ld, add the per-column bias, apply ReLU and store bf16 through TMA.  The pipeline is CUTLASS 4.x's
// sm100 collective (cutlass/gemm/collective/builders/sm100_umma_builder.inl), instantiated here for our shapes.
#include "policy_gemm_impl.cuh"

namespace uavp {
// wide-tile variants live in policy_gemm_wide.cu
int gemm_bias_act_256(const void *A, int64_t lda, const void *W, const float *bias, void *D, int M, int N, int K, int relu, void *workspace,
                      size_t workspace_bytes, cudaStream_t stream);
int gemm_bias_act_192(const void *A, int64_t lda, const void *W, const float *bias, void *D, int M, int N, int K, void *workspace,
                      size_t workspace_bytes, cudaStream_t stream);

size_t gemm_workspace_bytes() { return 1 << 20; }

int gemm_bias_act(const void *A, int64_t lda, const void *W, const float *bias, void *D, int M, int N, int K, int relu,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    // wide outputs: 128 x 256 (or 2 x 192 for the packed Q|K|V projection) tiles read every A tile once per output block
    // instead of once per 128 columns
    if (N % 256 == 0 && M >= 16384) return gemm_bias_act_256(A, lda, W, bias, D, M, N, K, relu, workspace, workspace_bytes, stream);
    if (N % 192 == 0 && M >= 16384 && !relu) return gemm_bias_act_192(A, lda, W, bias, D, M, N, K, workspace, workspace_bytes, stream);
    if (relu) return GemmT<cutlass::epilogue::thread::ReLu>::run(A, lda, W, bias, D, M, N, K, workspace, workspace_bytes, stream);
    return GemmT<cutlass::epilogue::thread::Identity>::run(A, lda, W, bias, D, M, N, K, workspace, workspace_bytes, stream);
}
}  // namespace uavp
