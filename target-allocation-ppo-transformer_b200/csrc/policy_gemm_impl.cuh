// policy_gemm_impl.cuh - the CUTLASS 4.x sm100 collective (TMA loads with 128B swizzle into a multi-stage shared-memory
// ring, one elected thread issuing tcgen05.mma into TMEM, epilogue warps reading it back with tcgen05.ld, bias / ReLU,
// TMA store) instantiated for one output tile width.  Included by the policy_gemm*.cu translation units, which each
// instantiate a few variants so that they compile in parallel.
#pragma once
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

template <template <class> class Act, class TileN = _128>
struct GemmT {
    using Elt = cutlass::bfloat16_t;
    using TileShape = Shape<_128, TileN, _64>;
    using ClusterShape = Shape<_1, _1, _1>;
    using Fusion = cutlass::epilogue::fusion::LinCombPerColBiasEltAct<Act, Elt, float, float>;
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

    static int run(const void *A, int64_t lda, const void *W, const float *bias, void *D, int M, int N, int K, void *ws,
                   size_t ws_bytes, cudaStream_t stream) {
        typename Kernel::StrideA sa;   // (lda, 1, batch)
        typename Kernel::StrideB sb;   // W [N,K] row-major == B [K,N] column-major: (K, 1, batch)
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
        args.epilogue.thread.bias_ptr = bias;
        Gemm gemm;
        if (gemm.can_implement(args) != cutlass::Status::kSuccess) return -1;
        if (Gemm::get_workspace_size(args) > ws_bytes) return -3;
        if (gemm.initialize(args, ws, stream) != cutlass::Status::kSuccess) return -2;
        return gemm.run(stream) == cutlass::Status::kSuccess ? 0 : -2;
    }
};

}  // namespace uavp
