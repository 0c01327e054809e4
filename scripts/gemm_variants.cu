// Experiment harness (not part of the library): times FastF32 GEMM variants on the MLP shape and
// reports the error against an fp64 reference on sampled outputs.  nvcc -DVARIANT=n ...
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
using namespace cute;

#ifndef VARIANT
#define VARIANT 0
#endif
#if VARIANT == 0
constexpr int TM = 128, TN = 128, TK = 16, CM = 1, BANDS = 5; constexpr bool TWO = false;
#elif VARIANT == 1
constexpr int TM = 128, TN = 128, TK = 16, CM = 1, BANDS = 3; constexpr bool TWO = false;
#elif VARIANT == 2
constexpr int TM = 256, TN = 128, TK = 16, CM = 2, BANDS = 5; constexpr bool TWO = true;
#elif VARIANT == 3
constexpr int TM = 256, TN = 128, TK = 16, CM = 2, BANDS = 3; constexpr bool TWO = true;
#elif VARIANT == 4
constexpr int TM = 128, TN = 128, TK = 32, CM = 1, BANDS = 5; constexpr bool TWO = false;
#elif VARIANT == 5
constexpr int TM = 256, TN = 128, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#elif VARIANT == 6
constexpr int TM = 128, TN = 128, TK = 32, CM = 1, BANDS = 3; constexpr bool TWO = false;
#elif VARIANT == 7
constexpr int TM = 256, TN = 128, TK = 16, CM = 2, BANDS = 4; constexpr bool TWO = true;
#elif VARIANT == 20
constexpr int TM = 128, TN = 128, TK = 32, CM = 1, BANDS = 5; constexpr bool TWO = false;
#define ACCP 2
#elif VARIANT == 21
constexpr int TM = 128, TN = 128, TK = 64, CM = 1, BANDS = 5; constexpr bool TWO = false;
#define ACCP 4
#elif VARIANT == 22
constexpr int TM = 256, TN = 128, TK = 32, CM = 2, BANDS = 5; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 23
constexpr int TM = 256, TN = 128, TK = 64, CM = 2, BANDS = 5; constexpr bool TWO = true;
#define ACCP 4
#elif VARIANT == 24
constexpr int TM = 256, TN = 128, TK = 64, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 4
#elif VARIANT == 25
constexpr int TM = 256, TN = 128, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 26
constexpr int TM = 256, TN = 128, TK = 64, CM = 2, BANDS = 4; constexpr bool TWO = true;
#define ACCP 4
#elif VARIANT == 27
constexpr int TM = 256, TN = 128, TK = 64, CM = 2, BANDS = 5; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 28
constexpr int TM = 256, TN = 128, TK = 48, CM = 2, BANDS = 5; constexpr bool TWO = true;
#define ACCP 3
#elif VARIANT == 29
constexpr int TM = 128, TN = 128, TK = 32, CM = 1, BANDS = 3; constexpr bool TWO = false;
#define ACCP 2
#elif VARIANT == 30
constexpr int TM = 256, TN = 208, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 31
constexpr int TM = 256, TN = 192, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 32
constexpr int TM = 256, TN = 224, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 33
constexpr int TM = 256, TN = 256, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 40
constexpr int TM = 128, TN = 80, TK = 32, CM = 1, BANDS = 3; constexpr bool TWO = false;
#define ACCP 2
#elif VARIANT == 41
constexpr int TM = 256, TN = 96, TK = 32, CM = 2, BANDS = 3; constexpr bool TWO = true;
#define ACCP 2
#elif VARIANT == 42
constexpr int TM = 128, TN = 208, TK = 32, CM = 1, BANDS = 3; constexpr bool TWO = false;
#define ACCP 2
#elif VARIANT == 8
constexpr int TM = 128, TN = 80, TK = 16, CM = 1, BANDS = 5; constexpr bool TWO = false;
#elif VARIANT == 9
constexpr int TM = 128, TN = 112, TK = 16, CM = 1, BANDS = 5; constexpr bool TWO = false;
#elif VARIANT == 10
constexpr int TM = 256, TN = 80, TK = 16, CM = 2, BANDS = 5; constexpr bool TWO = true;
#elif VARIANT == 11
constexpr int TM = 128, TN = 64, TK = 16, CM = 1, BANDS = 5; constexpr bool TWO = false;
#endif

#ifdef ACCP
constexpr int ACCPV = ACCP;
#else
constexpr int ACCPV = 1;
#endif
using LayoutA = cutlass::layout::RowMajor;
using LayoutB = cutlass::layout::ColumnMajor;
using LayoutC = cutlass::layout::RowMajor;
using Arch = cutlass::arch::Sm100;
using Op = cutlass::arch::OpClassTensorOp;
using TileS = Shape<Int<TM>, Int<TN>, Int<TK>>;
using ClusterS = Shape<Int<CM>, _1, _1>;
using Fusion = cutlass::epilogue::fusion::LinCombPerColBias<float, float, float>;
using EpiSched = cute::conditional_t<TWO, cutlass::epilogue::TmaWarpSpecialized2Sm, cutlass::epilogue::TmaWarpSpecialized1Sm>;
using MainSched = cute::conditional_t<TWO, cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32Sm100,
                                      cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>;
using Epi = typename cutlass::epilogue::collective::CollectiveBuilder<Arch, Op, TileS, ClusterS,
    cutlass::epilogue::collective::EpilogueTileAuto, float, float, float, LayoutC, 4, float, LayoutC, 4, EpiSched, Fusion>::CollectiveOp;
using Builder = cutlass::gemm::collective::CollectiveBuilder<Arch, Op, float, LayoutA, 4, float, LayoutB, 4, float, TileS, ClusterS,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename Epi::SharedStorage))>, MainSched>;
using Policy = cutlass::gemm::MainloopSm100TmaUmmaWarpSpecializedFastF32<
    Builder::Load2TransformPipelineStageCount, Builder::Transform2MmaPipelineStageCount, Builder::SchedulerPipelineStageCount,
    Builder::AccumulatorPipelineStageCount, BANDS, Builder::ScalingFactor, ACCPV, ClusterS,
    typename Builder::AccumulatorCopyAtom, Arch>;
using Main = cutlass::gemm::collective::CollectiveMma<Policy, TileS, float, cutlass::gemm::TagToStrideA_t<LayoutA>, float,
    cutlass::gemm::TagToStrideB_t<LayoutB>, typename Builder::TiledMma, typename Builder::GmemTiledCopyA,
    typename Builder::SmemLayoutAtomPairA, typename Builder::CopyAtomPairA, cute::identity, typename Builder::GmemTiledCopyB,
    typename Builder::SmemLayoutAtomPairB, typename Builder::CopyAtomPairB, cute::identity>;
using Kernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, Main, Epi>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<Kernel>;

int main(int argc, char** argv) {
  int M = argc > 1 ? atoi(argv[1]) : 65536, N = argc > 2 ? atoi(argv[2]) : 400, K = argc > 3 ? atoi(argv[3]) : 624;
  std::vector<float> hA((size_t)M * K), hB((size_t)N * K);
  srand(1);
  for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& v : hB) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  float *A, *B, *D;
  cudaMalloc(&A, hA.size() * 4); cudaMalloc(&B, hB.size() * 4); cudaMalloc(&D, (size_t)M * N * 4);
  cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  typename Kernel::StrideA sa; get<0>(sa) = K; get<2>(sa) = 0;
  typename Kernel::StrideB sb; get<0>(sb) = K; get<2>(sb) = 0;
  typename Kernel::StrideC sc; get<0>(sc) = N; get<2>(sc) = 0;
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {A, sa, B, sb}, {{}, D, sc, D, sc}};
  args.epilogue.thread.alpha = 1.f; args.epilogue.thread.beta = 0.f;
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) { printf("VARIANT %d cannot implement\n", VARIANT); return 1; }
  size_t wsb = Gemm::get_workspace_size(args);
  void* ws = nullptr; cudaMalloc(&ws, wsb + 256);
  if (gemm.initialize(args, ws) != cutlass::Status::kSuccess) { printf("init failed\n"); return 2; }
  for (int i = 0; i < 3; ++i) gemm.run();
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  for (int i = 0; i < 20; ++i) gemm.run();
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e); ms /= 20;
  cudaError_t err = cudaDeviceSynchronize();
  std::vector<float> hD((size_t)M * N);
  cudaMemcpy(hD.data(), D, hD.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int t = 0; t < 400; ++t) {
    int i = rand() % M, j = rand() % N;
    double r = 0; for (int k = 0; k < K; ++k) r += (double)hA[(size_t)i * K + k] * hB[(size_t)j * K + k];
    maxerr = fmax(maxerr, fabs(r - hD[(size_t)i * N + j])); maxref = fmax(maxref, fabs(r));
  }
  printf("VARIANT %d accp %d tile %dx%dx%d cluster %d bands %d stages(l2t %d, t2m %d, acc %d): %.3f ms %.1f TFLOP/s relerr %.2e (%s)\n",
         VARIANT, ACCPV, TM, TN, TK, CM, BANDS, Builder::Load2TransformPipelineStageCount, Builder::Transform2MmaPipelineStageCount,
         Builder::AccumulatorPipelineStageCount, ms, 2.0 * M * N * K / ms / 1e9, maxerr / maxref, cudaGetErrorString(err));
  return 0;
}
