// Inference gather from a pruned table stored as CSR.
//
// Replaces the reference's numba kernel (src/models/embeddings/pruned_embedding.py:140-173:
// one THREAD per looked-up id, a zero-fill loop and a scattered store loop per row) with the
// same launch shape as the training gather: one warp per sample, a lane group per looked-up
// row, the row expanded to its dense form in registers, and DeepFM's first-order term + FM
// second order (src/models/deepfm.py:88-98) reduced in the same pass.
//
// Expansion without a scatter: the lane group loads the row's (col, value) pairs, ORs the column
// bits into a presence mask M; dense dim d is present iff bit d of M is set and then its value
// sits in pair number popc(M & ((1<<d)-1)) (CSR columns are sorted and unique inside a row - the
// host wrapper guarantees it), fetched with a shuffle from the lane that loaded that pair.
//
// Index arrays: crow int64 or int32, col int64 / int32 / uint8 (the compact layout the wrapper
// offers: 5 bytes per kept weight instead of the reference's 12).
#include "common.cuh"

namespace rsb {

struct CsrArgs {
  const void* idx;
  int idx_i32;
  const long long* offsets;
  long long B;
  int F, D;
  const float* values;
  const void* crow;
  int crow_bytes;
  const void* col;
  int col_bytes;
  long long n_rows;
  const float* fc;
  const float* bias;
  float* out_emb;
  float* out_y;
  int* err;
};

__device__ __forceinline__ long long csr_crow(const CsrArgs& a, long long i) {
  return a.crow_bytes == 8 ? __ldg(reinterpret_cast<const long long*>(a.crow) + i)
                           : (long long)__ldg(reinterpret_cast<const int*>(a.crow) + i);
}

__device__ __forceinline__ unsigned csr_col(const CsrArgs& a, long long i) {
  if (a.col_bytes == 8) return (unsigned)__ldg(reinterpret_cast<const long long*>(a.col) + i);
  if (a.col_bytes == 4) return (unsigned)__ldg(reinterpret_cast<const int*>(a.col) + i);
  return (unsigned)__ldg(reinterpret_cast<const unsigned char*>(a.col) + i);
}

// G lanes per looked-up row, each lane owning PPL consecutive dense dims (G * PPL >= D) and loading PPL of
// the row's (col, value) pairs (pair j sits in lane j % G, register j / G).  Fewer lanes per row means more
// rows in flight per warp: the kernel is bound by the id -> row extent -> pair chain of dependent loads, so
// the field loop is batched kIter deep like the training gather (lookup.cu).
template <int G, int PPL, int kIter>
__global__ void __launch_bounds__(256) csr_lookup_fwd_kernel(CsrArgs a) {
  constexpr int GPW = kWarp / G;
  const int lane = threadIdx.x & 31;
  const int g = lane / G, c = lane % G;
  const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << (g * G));
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool fm = a.out_y != nullptr;
  const bool vec = (PPL == 2) && (a.D % 2 == 0);   // float2 stores need D even (row base stays 8-byte aligned)

  for (long long b = warp; b < a.B; b += nwarps) {
    float S[PPL], Q[PPL], first = 0.f;
#pragma unroll
    for (int k = 0; k < PPL; ++k) S[k] = Q[k] = 0.f;
    for (int f0 = 0; f0 < a.F; f0 += GPW * kIter) {
      long long row[kIter], left[kIter];
      int nnz[kIter];
      bool act[kIter];
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int f = f0 + it * GPW + g;
        act[it] = f < a.F;
        long long id = 0;
        if (act[it]) {
          id = a.idx_i32 ? (long long)__ldg(reinterpret_cast<const int*>(a.idx) + b * a.F + f)
                         : __ldg(reinterpret_cast<const long long*>(a.idx) + b * a.F + f);
          if (a.offsets) id += __ldg(a.offsets + f);
          if (id < 0 || id >= a.n_rows) {
            if (a.err) *a.err = 1;
            id = 0;
          }
        }
        row[it] = id;
      }
      float fcv[kIter];
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        left[it] = 0;
        nnz[it] = 0;
        fcv[it] = 0.f;
        if (act[it]) {
          left[it] = csr_crow(a, row[it]);
          nnz[it] = (int)(csr_crow(a, row[it] + 1) - left[it]);
          if (c == 0 && a.fc) fcv[it] = __ldg(a.fc + row[it]);
        }
      }
      unsigned colv[kIter];
      float valv[kIter][PPL];
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        colv[it] = 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const int j = c + k * G;
          valv[it][k] = 0.f;
          if (j < nnz[it]) {
            colv[it] |= 1u << (csr_col(a, left[it] + j) & 31u);
            valv[it][k] = __ldg(a.values + left[it] + j);
          }
        }
      }
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int f = f0 + it * GPW + g;
        const unsigned M = __reduce_or_sync(gmask, colv[it]);
        float e[PPL];
#pragma unroll
        for (int k = 0; k < PPL; ++k) {
          const int d = c * PPL + k;
          const int q = __popc(M & ((1u << d) - 1u));   // pair number holding dim d (if present)
          float v = __shfl_sync(gmask, valv[it][0], q % G, G);
          if (PPL == 2) {
            const float v1 = __shfl_sync(gmask, valv[it][1], q % G, G);
            v = (q >= G) ? v1 : v;
          }
          e[k] = ((M >> d) & 1u) ? v : 0.f;
        }
        first += fcv[it];
        if (act[it]) {
          float* o = a.out_emb + (b * a.F + f) * (long long)a.D + c * PPL;
          if (vec) {
            if (c * PPL < a.D) *reinterpret_cast<float2*>(o) = make_float2(e[0], e[PPL - 1]);
          } else {
#pragma unroll
            for (int k = 0; k < PPL; ++k)
              if (c * PPL + k < a.D) o[k] = e[k];
          }
#pragma unroll
          for (int k = 0; k < PPL; ++k) {   // dims >= D are never present in M: e == 0
            S[k] += e[k];
            Q[k] = fmaf(e[k], e[k], Q[k]);
          }
        }
      }
    }
    if (fm) {
      float y2 = 0.f;
#pragma unroll
      for (int k = 0; k < PPL; ++k) {
#pragma unroll
        for (int off = G; off < kWarp; off <<= 1) {
          S[k] += __shfl_xor_sync(kFull, S[k], off);
          Q[k] += __shfl_xor_sync(kFull, Q[k], off);
        }
        y2 += S[k] * S[k] - Q[k];
      }
#pragma unroll
      for (int off = 1; off < G; off <<= 1) y2 += __shfl_xor_sync(kFull, y2, off);
#pragma unroll
      for (int off = 1; off < kWarp; off <<= 1) first += __shfl_xor_sync(kFull, first, off);
      if (lane == 0) a.out_y[b] = first + (a.bias ? __ldg(a.bias) : 0.f) + 0.5f * y2;
    }
  }
}

template <int G, int PPL, int kIter>
static int launch_csr(const CsrArgs& a, cudaStream_t s) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (a.B + 7) / 8;  // 8 warps (samples) per CTA
  const long long cap = (long long)sms * 8 * 4;
  if (blocks > cap) blocks = cap;
  csr_lookup_fwd_kernel<G, PPL, kIter><<<(unsigned)blocks, 256, 0, s>>>(a);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int rsb_csr_lookup_fwd(const void* idx, int32_t idx_is_i32, const int64_t* offsets, int64_t B,
                                          int32_t F, int32_t D, const float* values, const void* crow,
                                          int32_t crow_bytes, const void* col, int32_t col_bytes, int64_t n_rows,
                                          const float* fc, const float* bias, float* out_emb, float* out_yfm,
                                          int32_t* err_flag, void* stream) {
  if (B < 0 || F <= 0 || D <= 0 || n_rows <= 0) return RSB_ERR_BAD_ARG;
  if (D > 32) return RSB_ERR_UNSUPPORTED;
  if (crow_bytes != 8 && crow_bytes != 4) return RSB_ERR_BAD_ARG;
  if (col_bytes != 8 && col_bytes != 4 && col_bytes != 1) return RSB_ERR_BAD_ARG;
  if (B == 0) return RSB_OK;
  if (!idx || !crow || !out_emb) return RSB_ERR_BAD_ARG;   // values/col may be NULL for an all-zero table
  CsrArgs a;
  a.idx = idx;
  a.idx_i32 = idx_is_i32;
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.B = B;
  a.F = F;
  a.D = D;
  a.values = values;
  a.crow = crow;
  a.crow_bytes = crow_bytes;
  a.col = col;
  a.col_bytes = col_bytes;
  a.n_rows = n_rows;
  a.fc = fc;
  a.bias = bias;
  a.out_emb = out_emb;
  a.out_y = out_yfm;
  a.err = err_flag;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (D <= 4) return launch_csr<4, 1, 4>(a, s);
  if (D <= 8) return launch_csr<4, 2, 4>(a, s);
  if (D <= 16) return launch_csr<8, 2, 5>(a, s);
  return launch_csr<16, 2, 5>(a, s);
}
