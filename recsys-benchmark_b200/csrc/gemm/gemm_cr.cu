// FastF32 GEMM instantiation: A ColumnMajor, B RowMajor (see fast_f32_gemm.cuh).
#include "fast_f32_gemm.cuh"

namespace rsb_gemm {
int64_t gemm_cr(const Problem& p, bool query_ws) {
  if (p.N <= 64) return run<FastF32<cutlass::layout::ColumnMajor, cutlass::layout::RowMajor, 64>, false, false>(p, query_ws);
  // large-M problems: CTA pairs on 256-row tiles; small M (weight gradients, split-K) keep 128-row tiles
  if (p.M >= 2048) return run<FastF32<cutlass::layout::ColumnMajor, cutlass::layout::RowMajor, 128, true>, false, false>(p, query_ws);
  return run<FastF32<cutlass::layout::ColumnMajor, cutlass::layout::RowMajor, 128>, false, false>(p, query_ws);
}
}  // namespace rsb_gemm
