// C ABI of the fp32-accurate tensor-core GEMM (see fast_f32_gemm.cuh).
#include "fast_f32_gemm.cuh"
#include "rsb.h"

namespace rsb {
void note_launch(int n);
}

static bool ok16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int make_problem(rsb_gemm::Problem& p, int32_t trans_a, int32_t trans_b, int64_t M, int64_t N, int64_t K,
                        int64_t batch, const float* A, int64_t lda, int64_t stride_a, const float* B, int64_t ldb,
                        int64_t stride_b, const float* C, float* D, int64_t ldd, int64_t stride_d,
                        const float* bias, float alpha, float beta) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0 || !A || !B || !D) return RSB_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff || batch > 0x7fffffff) return RSB_ERR_BAD_ARG;
  if (beta != 0.f && !C) return RSB_ERR_BAD_ARG;
  // TMA: 16-byte aligned bases, leading dimensions / batch strides / contiguous extents in multiples of 4 floats
  if (!ok16(A) || !ok16(B) || !ok16(D) || (C && !ok16(C)) || (bias && !ok16(bias))) return RSB_ERR_UNSUPPORTED;
  if (lda % 4 || ldb % 4 || ldd % 4 || stride_a % 4 || stride_b % 4 || stride_d % 4) return RSB_ERR_UNSUPPORTED;
  if (N % 4) return RSB_ERR_UNSUPPORTED;
  if ((trans_a ? M : K) % 4 || (trans_b ? K : N) % 4) return RSB_ERR_UNSUPPORTED;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.L = (int)batch;
  p.A = A; p.lda = lda; p.stride_a = stride_a;
  p.B = B; p.ldb = ldb; p.stride_b = stride_b;
  p.C = C; p.D = D; p.ldd = ldd; p.stride_d = stride_d;
  p.bias = bias; p.alpha = alpha; p.beta = beta;
  return RSB_OK;
}

static int64_t dispatch(const rsb_gemm::Problem& p, int32_t trans_a, int32_t trans_b, bool query) {
  if (!trans_a && trans_b) return rsb_gemm::gemm_rc(p, query);
  if (!trans_a && !trans_b) return rsb_gemm::gemm_rr(p, query);
  if (trans_a && !trans_b) return rsb_gemm::gemm_cr(p, query);
  return -2;
}

extern "C" RSB_API int64_t rsb_gemm_f32_workspace_bytes(int32_t trans_a, int32_t trans_b, int64_t M, int64_t N,
                                                        int64_t K, int64_t batch) {
  rsb_gemm::Problem p = {};
  static float dummy[4] __attribute__((aligned(16)));
  if (make_problem(p, trans_a, trans_b, M, N, K, batch, dummy, 4 * ((trans_a ? M : K) / 4 + 1), 0, dummy,
                   4 * ((trans_b ? K : N) / 4 + 1), 0, nullptr, dummy, 4 * (N / 4 + 1), 0, nullptr, 1.f, 0.f))
    return -1;
  int64_t r = dispatch(p, trans_a, trans_b, true);
  return r < 0 ? -1 : r + 256;
}

extern "C" RSB_API int rsb_gemm_f32(int32_t trans_a, int32_t trans_b, int64_t M, int64_t N, int64_t K, int64_t batch,
                                    const float* A, int64_t lda, int64_t stride_a, const float* B, int64_t ldb,
                                    int64_t stride_b, const float* C, float* D, int64_t ldd, int64_t stride_d,
                                    const float* bias, float alpha, float beta, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  rsb_gemm::Problem p = {};
  int rc = make_problem(p, trans_a, trans_b, M, N, K, batch, A, lda, stride_a, B, ldb, stride_b, C, D, ldd, stride_d,
                        bias, alpha, beta);
  if (rc) return rc;
  if (trans_a && trans_b) return RSB_ERR_UNSUPPORTED;
  p.ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  p.ws_bytes = workspace ? (size_t)(workspace_bytes > 256 ? workspace_bytes - 256 : 0) : 0;
  p.stream = reinterpret_cast<cudaStream_t>(stream);
  {
    // The TMA descriptors are encoded with a DRIVER API call, which needs a context current on the
    // calling thread; autograd worker threads may not have touched the runtime yet (error 201).
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaSetDevice(dev);
    if (e != cudaSuccess) return (int)e;
  }
  int64_t r = dispatch(p, trans_a, trans_b, false);
  if (r == 0) {
    rsb::note_launch(1);
    return RSB_OK;
  }
  if (r == 2) return RSB_ERR_WORKSPACE;
  if (r == 1) return RSB_ERR_UNSUPPORTED;
  cudaError_t e = cudaGetLastError();
  return e != cudaSuccess ? (int)e : RSB_ERR_BAD_ARG;
}
