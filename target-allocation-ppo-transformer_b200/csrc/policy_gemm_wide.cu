// policy_gemm_wide.cu - the 128 x 256 and 128 x 192 output-tile instantiations of the tcgen05 GEMM (policy_gemm_impl.cuh)
#include "policy_gemm_impl.cuh"

namespace uavp {
int gemm_bias_act_256(const void *A, int64_t lda, const void *W, const float *bias, void *D, int M, int N, int K, int relu, void *workspace,
                      size_t workspace_bytes, cudaStream_t stream) {
    if (relu) return GemmT<cutlass::epilogue::thread::ReLu, _256>::run(A, lda, W, bias, D, M, N, K, workspace, workspace_bytes, stream);
    return GemmT<cutlass::epilogue::thread::Identity, _256>::run(A, lda, W, bias, D, M, N, K, workspace, workspace_bytes, stream);
}
int gemm_bias_act_192(const void *A, int64_t lda, const void *W, const float *bias, void *D, int M, int N, int K, void *workspace,
                      size_t workspace_bytes, cudaStream_t stream) {
    return GemmT<cutlass::epilogue::thread::Identity, _192>::run(A, lda, W, bias, D, M, N, K, workspace, workspace_bytes, stream);
}
}  // namespace uavp
