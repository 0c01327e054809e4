// Bandwidth-bound glue of the dense tails (src/models/deepfm.py:55-66, src/models/dcn.py:56-66:
// Linear -> [BatchNorm1d] -> ReLU -> Dropout), fused so that every activation tensor is read and
// written once per direction:
//   relu_dropout_fwd : y = dropout(relu(x)) + a 1-byte keep mask        (torch: clamp + fused_dropout)
//   relu_dropout_bwd : gx = g * mask * 1/(1-p), and the column sums of gx (= the bias gradient of the
//                      preceding Linear) accumulated in the same pass     (torch: masked_scale + threshold_backward + sum)
//   colsum           : deterministic two-stage column sum (bias gradients, split-K partial sums)
// Dropout uses a counter-based Philox4x32-10 stream keyed by (seed, offset); the mask is random by
// definition, so only its statistics (keep probability 1-p, scale 1/(1-p)) are part of the contract.
#include "common.cuh"

namespace rsb {

// numel must be a multiple of 4 and x/y 16-byte aligned (checked by the host function)
__global__ void __launch_bounds__(256) relu_dropout_fwd_kernel(const float4* __restrict__ x, long long n4,
                                                               float p, float scale, unsigned long long seed,
                                                               unsigned long long offset,
                                                               const unsigned long long* __restrict__ offset_dev,
                                                               float4* __restrict__ y, uchar4* __restrict__ mask) {
  Philox rng{(unsigned)seed, (unsigned)(seed >> 32)};
  if (offset_dev) offset += *offset_dev;   // CUDA-graph replays: the stream position lives in device memory
  const unsigned thr = (unsigned)(p * 4294967296.0);  // keep iff r >= thr
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    const uint4 r = rng(offset + (unsigned long long)i);
    uchar4 m;
    m.x = (v.x > 0.f) && (r.x >= thr);
    m.y = (v.y > 0.f) && (r.y >= thr);
    m.z = (v.z > 0.f) && (r.z >= thr);
    m.w = (v.w > 0.f) && (r.w >= thr);
    float4 o;
    o.x = m.x ? v.x * scale : 0.f;
    o.y = m.y ? v.y * scale : 0.f;
    o.z = m.z ? v.z * scale : 0.f;
    o.w = m.w ? v.w * scale : 0.f;
    y[i] = o;
    mask[i] = m;
  }
}

// y = dropout(relu(x)) as above AND out[r] = y[r,:] . w + bias: the last hidden block of the MLP feeds a Linear
// with ONE output (src/models/deepfm.py:64: Linear(hidden, 1)), a matrix-vector product that would re-read y.
// One warp per row; same Philox counters as the flat kernel (float4 index r * n4 + c4), so the masks agree.
__global__ void __launch_bounds__(256) relu_dropout_dot_fwd_kernel(const float* __restrict__ x, long long M, int N,
                                                                   float p, float scale, unsigned long long seed,
                                                                   unsigned long long offset,
                                                                   const unsigned long long* __restrict__ offset_dev,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ affine,
                                                                   float* __restrict__ y,
                                                                   unsigned char* __restrict__ mask,
                                                                   float* __restrict__ out) {
  Philox rng{(unsigned)seed, (unsigned)(seed >> 32)};
  if (offset_dev) offset += *offset_dev;
  const unsigned thr = (unsigned)(p * 4294967296.0);
  const int lane = threadIdx.x & 31;
  const int n4 = N / 4;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float b0 = bias ? __ldg(bias) : 0.f;
  for (long long r = warp; r < M; r += nwarps) {
    float dot = 0.f;
    for (int c0 = 0; c0 < n4; c0 += 128) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c4 = c0 + u * 32 + lane;
        v[u] = (c4 < n4) ? __ldg(reinterpret_cast<const float4*>(x + r * (long long)N) + c4)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c4 = c0 + u * 32 + lane;
        if (c4 < n4) {
          const long long i = r * (long long)n4 + c4;
          const uint4 q = rng(offset + (unsigned long long)i);
          const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + c4);
          if (affine) {     // BatchNorm folded in: x * scale + shift per column
            const float4 sc = __ldg(reinterpret_cast<const float4*>(affine) + c4);
            const float4 sh = __ldg(reinterpret_cast<const float4*>(affine + N) + c4);
            v[u] = make_float4(fmaf(v[u].x, sc.x, sh.x), fmaf(v[u].y, sc.y, sh.y), fmaf(v[u].z, sc.z, sh.z),
                               fmaf(v[u].w, sc.w, sh.w));
          }
          uchar4 m;
          m.x = (v[u].x > 0.f) && (q.x >= thr);
          m.y = (v[u].y > 0.f) && (q.y >= thr);
          m.z = (v[u].z > 0.f) && (q.z >= thr);
          m.w = (v[u].w > 0.f) && (q.w >= thr);
          float4 o;
          o.x = m.x ? v[u].x * scale : 0.f;
          o.y = m.y ? v[u].y * scale : 0.f;
          o.z = m.z ? v[u].z * scale : 0.f;
          o.w = m.w ? v[u].w * scale : 0.f;
          reinterpret_cast<float4*>(y)[i] = o;
          reinterpret_cast<uchar4*>(mask)[i] = m;
          dot = fmaf(o.x, wv.x, dot);
          dot = fmaf(o.y, wv.y, dot);
          dot = fmaf(o.z, wv.z, dot);
          dot = fmaf(o.w, wv.w, dot);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(kFull, dot, off);
    if (lane == 0) out[r] = dot + b0;
  }
}

// One CTA owns a slab of rows.  The 256 threads are arranged as (row lane ty) x (column group tx):
// tx owns the float4 column groups tx, tx+ctx, ... (a row is read with coalesced 128-bit loads), the
// cty row lanes interleave over the slab's rows; column sums stay in registers, are folded over ty in
// shared memory (fixed order) and leave the CTA as one partial row.
constexpr int kColsPerThread = 2;  // float4 groups per thread per pass

// MODE 0: column sums of g.  1: gx = g * mask * scale (relu_dropout_bwd) + column sums of gx.
// 2: as 1 with the rank-1 upstream gradient g[r,c] = rowv[r] * colv[c] formed on the fly (the layer in front of
//    a Linear with ONE output: its dX = g_logit[r] * w[c] is never materialised).
// 3: column sums of rowv[r] * g[r,c] (weight gradient of a one-output Linear: sum_r g_logit[r] * y[r,:]).
template <int MODE>
__global__ void __launch_bounds__(256, MODE == 1 ? 6 : 4) colsum_slab_kernel(const float* __restrict__ g,
                                                          const unsigned char* __restrict__ mask, float scale,
                                                          long long M, int N, long long ld, float* __restrict__ gx,
                                                          float* __restrict__ partials, int col0,
                                                          const float* __restrict__ rowv,
                                                          const float* __restrict__ colv) {
  constexpr bool MASKED = (MODE == 1 || MODE == 2);
  __shared__ float4 red[256][kColsPerThread];
  const int ctx = blockDim.x, cty = blockDim.y;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n4 = N / 4;
  const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = rows_per * blockIdx.x;
  long long r1 = r0 + rows_per;
  if (r1 > M) r1 = M;
  float4 acc[kColsPerThread], cv[kColsPerThread];
#pragma unroll
  for (int j = 0; j < kColsPerThread; ++j) {
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    cv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c4 = col0 / 4 + tx + j * ctx;
    if (MODE == 2 && c4 < n4) cv[j] = __ldg(reinterpret_cast<const float4*>(colv) + c4);
  }
  // rows in flight per thread / resident CTAs per SM, tuned per mode on the MLP shapes (B = 65536, N = 400)
  constexpr int RU = (MODE == 1) ? 1 : 2;
  for (long long rb = r0 + ty; rb < r1; rb += (long long)cty * RU) {
    float4 v[RU][kColsPerThread];
    uchar4 m[RU][kColsPerThread];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
#pragma unroll
      for (int j = 0; j < kColsPerThread; ++j) {
        const int c4 = col0 / 4 + tx + j * ctx;
        const long long r = rb + (long long)u * cty;
        v[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        m[u][j] = make_uchar4(0, 0, 0, 0);
        if (c4 < n4 && r < r1) {
          if (MODE == 2) {
            const float rv = __ldg(rowv + r);
            v[u][j] = make_float4(rv * cv[j].x, rv * cv[j].y, rv * cv[j].z, rv * cv[j].w);
          } else {
            v[u][j] = __ldg(reinterpret_cast<const float4*>(g + r * ld) + c4);
            if (MODE == 3) {
              const float rv = __ldg(rowv + r);
              v[u][j].x *= rv; v[u][j].y *= rv; v[u][j].z *= rv; v[u][j].w *= rv;
            }
          }
          if (MASKED) m[u][j] = __ldg(reinterpret_cast<const uchar4*>(mask + r * (long long)N) + c4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
#pragma unroll
      for (int j = 0; j < kColsPerThread; ++j) {
        const int c4 = col0 / 4 + tx + j * ctx;
        const long long r = rb + (long long)u * cty;
        if (c4 < n4 && r < r1) {
          float4 x = v[u][j];
          if (MASKED) {
            x.x = m[u][j].x ? x.x * scale : 0.f;
            x.y = m[u][j].y ? x.y * scale : 0.f;
            x.z = m[u][j].z ? x.z * scale : 0.f;
            x.w = m[u][j].w ? x.w * scale : 0.f;
            reinterpret_cast<float4*>(gx + r * (long long)N)[c4] = x;
          }
          acc[j].x += x.x;
          acc[j].y += x.y;
          acc[j].z += x.z;
          acc[j].w += x.w;
        }
      }
    }
  }
  if (partials) {
    const int t = ty * ctx + tx;
#pragma unroll
    for (int j = 0; j < kColsPerThread; ++j) red[t][j] = acc[j];
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int j = 0; j < kColsPerThread; ++j) {
        const int c4 = col0 / 4 + tx + j * ctx;
        float4 s4 = red[tx][j];
        for (int y = 1; y < cty; ++y) {
          const float4 o = red[y * ctx + tx][j];
          s4.x += o.x; s4.y += o.y; s4.z += o.z; s4.w += o.w;
        }
        if (c4 < n4) reinterpret_cast<float4*>(partials + (long long)blockIdx.x * N)[c4] = s4;
      }
    }
  }
}

// final reduction of the CTA partials: 32 columns x 32 row lanes per CTA, coalesced along the columns,
// fixed-order fold over the row lanes in shared memory
__global__ void __launch_bounds__(1024) colsum_final_kernel(const float* __restrict__ partials, int nblk, int N,
                                                            float* __restrict__ out) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < N) {
#pragma unroll 4
    for (int b = ty; b < nblk; b += 32) s += __ldg(partials + (long long)b * N + c);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) t += red[y][tx];
    out[c] = t;
  }
}

static int slab_blocks(long long M, int per_sm = 6) {
  long long b = (long long)per_sm * sm_count();      // = the resident CTAs (launch bound 3 per SM): one full wave; >= 16 rows per CTA
  if (b > M / 16) b = M / 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int rsb_relu_dropout_fwd(const float* x, int64_t numel, float p, uint64_t seed, uint64_t offset,
                                            const uint64_t* offset_dev, float* y, uint8_t* mask, void* stream) {
  if (numel < 0 || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  if (numel == 0) return RSB_OK;
  if (!x || !y || !mask) return RSB_ERR_BAD_ARG;
  if (numel % 4 || !aligned16(x) || !aligned16(y) || (reinterpret_cast<uintptr_t>(mask) & 3u)) return RSB_ERR_UNSUPPORTED;
  const long long n4 = numel / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  relu_dropout_fwd_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), n4, p, 1.0f / (1.0f - p), seed, offset,
      reinterpret_cast<const unsigned long long*>(offset_dev), reinterpret_cast<float4*>(y),
      reinterpret_cast<uchar4*>(mask));
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int64_t rsb_colsum_workspace_bytes(int64_t M, int32_t N) {
  if (M < 0 || N <= 0) return 0;
  return (int64_t)slab_blocks(M) * N * 4 + 256;
}

static int colsum_impl(const float* g, const uint8_t* mask, float scale, int64_t M, int32_t N, int64_t ld, float* gx,
                       float* colsum, void* workspace, int64_t workspace_bytes, cudaStream_t s,
                       const float* rowv = nullptr, const float* colv = nullptr) {
  if (M < 0 || N <= 0) return RSB_ERR_BAD_ARG;
  if (M == 0) {
    if (colsum) cudaMemsetAsync(colsum, 0, (size_t)N * 4, s);
    return RSB_OK;
  }
  const int mode = mask ? (colv ? 2 : 1) : (rowv ? 3 : 0);
  if (mode != 2 && !g) return RSB_ERR_BAD_ARG;
  if ((mode == 2 || mode == 3) && !rowv) return RSB_ERR_BAD_ARG;
  if (colv && !aligned16(colv)) return RSB_ERR_UNSUPPORTED;
  if (N % 4 || ld % 4 || (g && !aligned16(g)) || (gx && !aligned16(gx)) || (mask && (reinterpret_cast<uintptr_t>(mask) & 3u)))
    return RSB_ERR_UNSUPPORTED;
  float* partials = nullptr;
  const int mode_ = mask ? (colv ? 2 : 1) : (rowv ? 3 : 0);
  const int nblk = slab_blocks(M, mode_ == 1 ? 6 : 4);
  if (colsum) {
    if (!workspace || workspace_bytes < rsb_colsum_workspace_bytes(M, N)) return RSB_ERR_WORKSPACE;
    partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  }
  // thread arrangement: exactly enough column threads for one pass over the row (N = 400: 50 threads x 2 float4
  // groups), the rest of the 256 as row lanes (5) - a power-of-two split left 22 % of the lanes idle there
  int ctx = (N / 4 + kColsPerThread - 1) / kColsPerThread;
  if (ctx > 256) ctx = 256;
  const dim3 block(ctx, 256 / ctx);
  const int per_pass = ctx * 4 * kColsPerThread;
  for (int col0 = 0; col0 < N; col0 += per_pass) {
    if (mode == 1)
      colsum_slab_kernel<1><<<nblk, block, 0, s>>>(g, mask, scale, M, N, ld, gx, partials, col0, nullptr, nullptr);
    else if (mode == 2)
      colsum_slab_kernel<2><<<nblk, block, 0, s>>>(nullptr, mask, scale, M, N, ld, gx, partials, col0, rowv, colv);
    else if (mode == 3)
      colsum_slab_kernel<3><<<nblk, block, 0, s>>>(g, nullptr, 1.f, M, N, ld, nullptr, partials, col0, rowv, nullptr);
    else
      colsum_slab_kernel<0><<<nblk, block, 0, s>>>(g, nullptr, 1.f, M, N, ld, nullptr, partials, col0, nullptr,
                                                   nullptr);
    RSB_CHECK_LAUNCH();
    note_launch(1);
  }
  if (colsum) {
    colsum_final_kernel<<<(N + 31) / 32, dim3(32, 32), 0, s>>>(partials, nblk, N, colsum);
    RSB_CHECK_LAUNCH();
    note_launch(1);
  }
  return RSB_OK;
}

extern "C" RSB_API int rsb_colsum(const float* x, int64_t M, int32_t N, int64_t ld, float* out, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  if (!out) return RSB_ERR_BAD_ARG;
  return colsum_impl(x, nullptr, 1.f, M, N, ld, nullptr, out, workspace, workspace_bytes,
                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" RSB_API int rsb_relu_dropout_bwd(const float* g, const uint8_t* mask, int64_t M, int32_t N, float p,
                                            float* gx, float* colsum, void* workspace, int64_t workspace_bytes,
                                            void* stream) {
  if (!mask || !gx || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  return colsum_impl(g, mask, 1.0f / (1.0f - p), M, N, N, gx, colsum, workspace, workspace_bytes,
                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" RSB_API int rsb_colsum_weighted(const float* x, const float* row_weight, int64_t M, int32_t N, int64_t ld,
                                           float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!out || !row_weight) return RSB_ERR_BAD_ARG;
  return colsum_impl(x, nullptr, 1.f, M, N, ld, nullptr, out, workspace, workspace_bytes,
                     reinterpret_cast<cudaStream_t>(stream), row_weight, nullptr);
}

extern "C" RSB_API int rsb_relu_dropout_bwd_rank1(const float* g_row, const float* w_col, const uint8_t* mask, int64_t M,
                                                  int32_t N, float p, float* gx, float* colsum, void* workspace,
                                                  int64_t workspace_bytes, void* stream) {
  if (!g_row || !w_col || !mask || !gx || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  return colsum_impl(nullptr, mask, 1.0f / (1.0f - p), M, N, N, gx, colsum, workspace, workspace_bytes,
                     reinterpret_cast<cudaStream_t>(stream), g_row, w_col);
}

extern "C" RSB_API int rsb_relu_dropout_dot_fwd(const float* x, int64_t M, int32_t N, float p, uint64_t seed,
                                                uint64_t offset, const uint64_t* offset_dev, const float* w,
                                                const float* bias, const float* affine, float* y, uint8_t* mask,
                                                float* out, void* stream) {
  if (M < 0 || N <= 0 || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  if (M == 0) return RSB_OK;
  if (!x || !y || !mask || !w || !out) return RSB_ERR_BAD_ARG;
  if (N % 4 || !aligned16(x) || !aligned16(y) || !aligned16(w) || (affine && !aligned16(affine)) ||
      (reinterpret_cast<uintptr_t>(mask) & 3u))
    return RSB_ERR_UNSUPPORTED;
  long long blocks = (M + 7) / 8;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  relu_dropout_dot_fwd_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, M, N, p, 1.0f / (1.0f - p), seed, offset, reinterpret_cast<const unsigned long long*>(offset_dev), w, bias,
      affine, y, mask, out);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}
