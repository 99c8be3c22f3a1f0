// policy_gemm_drelu.cu - activation-gradient GEMM with the ReLU backward fused into its epilogue (sm_100a, CUTLASS 4.x
// sm100 collective with an aux-load epilogue visitor).
#include "policy_gemm.cuh"

#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"


namespace uavp {
using namespace cute;

// gradient of ReLU given its OUTPUT z: d where z > 0, else 0 (element and fragment forms, as the epilogue visitor calls them)
template <class T>
struct ReluGradByOutput {
    CUTLASS_HOST_DEVICE T operator()(T d, T z) const { return z > T(0) ? d : T(0); }
};
template <class T, int N>
struct ReluGradByOutput<cutlass::Array<T, N>> {
    template <class U>
    CUTLASS_HOST_DEVICE cutlass::Array<T, N> operator()(cutlass::Array<T, N> const &d, cutlass::Array<U, N> const &z) const {
        cutlass::Array<T, N> y;
        CUTLASS_PRAGMA_UNROLL
        for (int i = 0; i < N; ++i) y[i] = float(static_cast<U>(z[i])) > 0.0f ? static_cast<T>(d[i]) : T(0);
        return y;
    }
};

// D[M,N] = (A[M,K] W[N,K]^T) masked by aux[M,N] != 0: the activation-gradient GEMM in front of a ReLU with the ReLU's
// backward fused into the epilogue (aux = the forward's post-ReLU activations, loaded by TMA next to the accumulator)
struct GemmDRelu {
    using Elt = cutlass::bfloat16_t;
    using TileShape = Shape<_128, _128, _64>;
    using ClusterShape = Shape<_1, _1, _1>;
    using Fusion = cutlass::epilogue::fusion::LinCombDeEltAct<cutlass::layout::RowMajor, ReluGradByOutput, Elt, float, Elt>;
    using Epilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
        cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, TileShape, ClusterShape,
        cutlass::epilogue::collective::EpilogueTileAuto, float, float, Elt, cutlass::layout::RowMajor, 8, Elt,
        cutlass::layout::RowMajor, 8, cutlass::epilogue::collective::EpilogueScheduleAuto, Fusion>::CollectiveOp;
    using Mainloop = typename cutlass::gemm::collective::CollectiveBuilder<
        cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, Elt, cutlass::layout::RowMajor, 8, Elt,
        cutlass::layout::ColumnMajor, 8, float, TileShape, ClusterShape,
        cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename Epilogue::SharedStorage))>,
        cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
    using Kernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, Mainloop, Epilogue, void>;
    using Gemm = cutlass::gemm::device::GemmUniversalAdapter<Kernel>;

    static int run(const void *A, int64_t lda, const void *W, const void *aux, int64_t ld_aux, void *D, int M, int N, int K, void *ws,
                   size_t ws_bytes, cudaStream_t stream) {
        typename Kernel::StrideA sa;
        typename Kernel::StrideB sb;
        typename Kernel::StrideC sc;
        typename Kernel::StrideD sd;
        get<0>(sa) = lda; get<2>(sa) = 0;
        get<0>(sb) = (int64_t)K; get<2>(sb) = 0;
        get<0>(sc) = (int64_t)N; get<2>(sc) = 0;
        get<0>(sd) = (int64_t)N; get<2>(sd) = 0;
        typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                      {M, N, K, 1},
                                      {static_cast<const Elt *>(A), sa, static_cast<const Elt *>(W), sb},
                                      {{}, nullptr, sc, static_cast<Elt *>(D), sd}};
        args.epilogue.thread.alpha = 1.0f;
        args.epilogue.thread.beta = 0.0f;
        args.epilogue.thread.aux_ptr = static_cast<const Elt *>(aux);
        get<0>(args.epilogue.thread.dAux) = ld_aux; get<2>(args.epilogue.thread.dAux) = 0;
        Gemm gemm;
        if (gemm.can_implement(args) != cutlass::Status::kSuccess) return -1;
        if (Gemm::get_workspace_size(args) > ws_bytes) return -3;
        if (gemm.initialize(args, ws, stream) != cutlass::Status::kSuccess) return -2;
        return gemm.run(stream) == cutlass::Status::kSuccess ? 0 : -2;
    }
};

int gemm_drelu(const void *A, int64_t lda, const void *W, const void *aux, int64_t ld_aux, void *D, int M, int N, int K, void *workspace,
               size_t workspace_bytes, cudaStream_t stream) {
    return GemmDRelu::run(A, lda, W, aux, ld_aux, D, M, N, K, workspace, workspace_bytes, stream);
}
}  // namespace uavp
