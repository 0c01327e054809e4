// Library bookkeeping and the full-table helpers of the lightweight-embedding variants.
#include <stdio.h>

#include <atomic>

#include "common.cuh"

namespace rsb {

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

__device__ __forceinline__ float pep_s_at(const float* s, int type, long long row, int d, int D) {
  switch (type) {
    case RSB_PEP_GLOBAL: return __ldg(s);
    case RSB_PEP_DIMENSION: return __ldg(s + d);
    case RSB_PEP_FEATURE: return __ldg(s + row);
    default: return __ldg(s + row * D + d);
  }
}

__device__ __forceinline__ void block_count_add(unsigned long long local, unsigned long long* count) {
  // warp reduce then one atomic per warp
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(kFull, local, off);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// soft_threshold over the whole table (pep_embedding.py:91-92) + count_nonzero (:127-130)
__global__ void pep_threshold_table_kernel(const float* __restrict__ w, const float* __restrict__ s, int type,
                                           long long n_rows, int D, float* __restrict__ out,
                                           unsigned long long* count) {
  const long long total = n_rows * D;
  unsigned long long nz = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long row = i / D;
    int d = (int)(i - row * D);
    float v = __ldg(w + i);
    float sg = sigmoidf_exact(pep_s_at(s, type, row, d, D));
    float m = fmaxf(fabsf(v) - sg, 0.0f);
    float r = (v > 0.f) ? m : ((v < 0.f) ? -m : 0.f);
    if (out) out[i] = r;
    nz += (r != 0.f);
  }
  if (count) block_count_add(nz, count);
}

__global__ void pep_dense_bwd_kernel(const float* __restrict__ w, const float* __restrict__ s, int type,
                                     long long n_rows, int D, const float* __restrict__ g_table,
                                     float* __restrict__ g_w, float* __restrict__ g_s) {
  const long long total = n_rows * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long row = i / D;
    int d = (int)(i - row * D);
    float v = __ldg(w + i);
    float g = __ldg(g_table + i);
    float sg = sigmoidf_exact(pep_s_at(s, type, row, d, D));
    bool keep = (fabsf(v) - sg) > 0.f;
    float sgn = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
    g_w[i] = keep ? g * sgn * sgn : 0.f;
    if (g_s) g_s[i] = keep ? -(g * sgn) * (sg * (1.0f - sg)) : 0.f;
  }
}

// one lane group per row would be overkill here: one thread per row, D is small.
__global__ void optembed_eval_weight_kernel(const float* __restrict__ w, const float* __restrict__ t_row,
                                            const long long* __restrict__ mask_d_row, int norm, long long n_rows,
                                            int D, float* __restrict__ out, unsigned long long* count) {
  unsigned long long nz = 0;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += (long long)gridDim.x * blockDim.x) {
    const float* wr = w + row * D;
    float keep = 1.f;
    if (t_row) {
      float nrm = 0.f;
      for (int d = 0; d < D; ++d) {
        float v = __ldg(wr + d);
        nrm += (norm == 2) ? v * v : fabsf(v);
      }
      if (norm == 2) nrm = sqrtf(nrm);
      keep = (nrm - __ldg(t_row + row) > 0.f) ? 1.f : 0.f;
    }
    long long k = mask_d_row ? __ldg(mask_d_row + row) : (long long)D;
    for (int d = 0; d < D; ++d) {
      float r = ((long long)d <= k) ? __ldg(wr + d) * keep : 0.f;
      if (out) out[row * D + d] = r;
      nz += (r != 0.f);
    }
  }
  if (count) block_count_add(nz, count);
}

__global__ void mask_table_kernel(const float* __restrict__ w, const unsigned char* __restrict__ m, long long numel,
                                  float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __ldg(w + i) * (float)(__ldg(m + i) != 0);
}

static unsigned grid_for(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int64_t rsb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" RSB_API const char* rsb_version(void) { return "rsb 0.1.0 (sm_100a)"; }

extern "C" RSB_API const char* rsb_error_string(int code) {
  switch (code) {
    case RSB_OK: return "ok";
    case RSB_ERR_BAD_ARG: return "rsb: bad argument (null pointer, negative size or unknown enum)";
    case RSB_ERR_UNSUPPORTED: return "rsb: unsupported row width (need width %4==0 && <=128 with 16B-aligned rows, or width <=32)";
    case RSB_ERR_WORKSPACE: return "rsb: workspace too small";
    default: return cudaGetErrorString((cudaError_t)code);
  }
}

extern "C" RSB_API int rsb_row_width_supported(int32_t width) { return row_shape(width, true).ok ? 1 : 0; }

extern "C" RSB_API int rsb_pep_threshold_table(const float* weight, const float* s, int32_t threshold_type, int64_t n_rows,
                                       int32_t D, float* out, int64_t* count, void* stream) {
  if (!weight || !s || n_rows < 0 || D <= 0) return RSB_ERR_BAD_ARG;
  if (threshold_type < RSB_PEP_GLOBAL || threshold_type > RSB_PEP_FEATURE_DIM) return RSB_ERR_BAD_ARG;
  if (n_rows == 0) return RSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  pep_threshold_table_kernel<<<grid_for(n_rows * D, 256), 256, 0, st>>>(
      weight, s, threshold_type, n_rows, D, out, reinterpret_cast<unsigned long long*>(count));
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_pep_dense_bwd(const float* weight, const float* s, int32_t threshold_type, int64_t n_rows,
                                 int32_t D, const float* g_table, float* g_weight, float* g_s_full, void* stream) {
  if (!weight || !s || !g_table || !g_weight || n_rows < 0 || D <= 0) return RSB_ERR_BAD_ARG;
  if (threshold_type < RSB_PEP_GLOBAL || threshold_type > RSB_PEP_FEATURE_DIM) return RSB_ERR_BAD_ARG;
  if (n_rows == 0) return RSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  pep_dense_bwd_kernel<<<grid_for(n_rows * D, 256), 256, 0, st>>>(weight, s, threshold_type, n_rows, D, g_table,
                                                                   g_weight, g_s_full);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_optembed_eval_weight(const float* weight, const float* t_row, const int64_t* mask_d_row,
                                        int32_t norm, int64_t n_rows, int32_t D, float* out, int64_t* count,
                                        void* stream) {
  if (!weight || n_rows < 0 || D <= 0 || (norm != 1 && norm != 2)) return RSB_ERR_BAD_ARG;
  if (n_rows == 0) return RSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  optembed_eval_weight_kernel<<<grid_for(n_rows, 128), 128, 0, st>>>(
      weight, t_row, reinterpret_cast<const long long*>(mask_d_row), norm, n_rows, D, out,
      reinterpret_cast<unsigned long long*>(count));
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

__global__ void sigmoid_kernel(const float* __restrict__ s, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = sigmoidf_exact(__ldg(s + i));
}

extern "C" RSB_API int rsb_sigmoid(const float* s, int64_t numel, float* out, void* stream) {
  if (!s || !out || numel < 0) return RSB_ERR_BAD_ARG;
  if (numel == 0) return RSB_OK;
  sigmoid_kernel<<<grid_for(numel, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(s, numel, out);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_mask_table(const float* weight, const uint8_t* mask, int64_t numel, float* out, void* stream) {
  if (!weight || !mask || !out || numel < 0) return RSB_ERR_BAD_ARG;
  if (numel == 0) return RSB_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  mask_table_kernel<<<grid_for(numel, 256), 256, 0, st>>>(weight, mask, numel, out);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}
