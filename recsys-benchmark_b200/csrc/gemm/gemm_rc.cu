// FastF32 GEMM instantiation: A RowMajor, B ColumnMajor (see fast_f32_gemm.cuh).
#include "fast_f32_gemm.cuh"

namespace rsb_gemm {
int64_t gemm_rc(const Problem& p, bool query_ws) {
  if (p.N <= 64) return run<FastF32<cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, 64>, true, true>(p, query_ws);
  // large-M problems: CTA pairs on 256-row tiles; small M (weight gradients, split-K) keep 128-row tiles
  if (p.M >= 2048) return run<FastF32<cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, 128, true>, true, true>(p, query_ws);
  return run<FastF32<cutlass::layout::RowMajor, cutlass::layout::ColumnMajor, 128>, true, true>(p, query_ws);
}
}  // namespace rsb_gemm
