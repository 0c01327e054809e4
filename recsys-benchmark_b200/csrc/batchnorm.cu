// BatchNorm1d of the dense tails in training mode (src/models/deepfm.py:57-58, src/models/dcn.py:56-66:
// Linear -> BatchNorm1d -> ReLU -> Dropout), fused with its neighbours so that every activation is read / written
// once per pass and leaves in the operand format of the next tensor-core GEMM (bf16 planes, see gemm/planes_gemm.cu):
//
//   forward   bn_stats        column sums of (z - k), (z - k)^2 over row slabs (k = row 0 of z: a shift that keeps the
//                             variance free of the E[z^2] - mean^2 cancellation), deterministic two-stage reduction
//             bn_finalize     mean, biased variance -> rstd, the affine (scale, shift), running statistics update
//             bn_apply        y = dropout_p(relu(z * scale + shift)) -> planes (+ ones column) + keep-and-positive mask
//   backward  bn_bwd_stats    column sums of g and g * xhat  (= d beta, d gamma), g = gradient w.r.t. the BN output
//             bn_bwd_apply    gz = gamma * rstd * (g - d beta / M - xhat * d gamma / M) -> planes
//
// torch semantics reproduced: normalisation by the biased batch variance, running_var updated with the unbiased one,
// eps inside the square root, momentum as an exponential moving average factor.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rsb {

constexpr int kBnThreads = 128;   // one float4 column group per thread per pass: 512 columns per pass

// partials[blk][0][c] = sum_r a[r,c] ; partials[blk][1][c] = sum_r b[r,c] over the block's row slab, where
//   MODE 0 (forward):  a = z - k_c,  b = (z - k_c)^2          (k = z[0, :])
//   MODE 1 (backward): a = g,        b = g * (z - mean_c) * rstd_c
template <int MODE>
__global__ void __launch_bounds__(kBnThreads) bn_slab_sums_kernel(const float* __restrict__ z, const float* __restrict__ g,
                                                                  long long M, int N, long long ldz, long long ldg,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  float* __restrict__ partials) {
  const int n4 = N / 4;
  const long long rows_per = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = rows_per * blockIdx.x;
  long long r1 = r0 + rows_per;
  if (r1 > M) r1 = M;
  for (int c4 = threadIdx.x; c4 < n4; c4 += kBnThreads) {
    float4 k, rs = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == 0) {
      k = __ldg(reinterpret_cast<const float4*>(z) + c4);
    } else {
      k = __ldg(reinterpret_cast<const float4*>(mean) + c4);
      rs = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
    }
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    long long r = r0;
    constexpr int U = MODE == 0 ? 8 : 4;   // rows in flight (the backward reads two matrices per row)
    for (; r + U <= r1; r += U) {
      float4 zv[U], gv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        zv[u] = __ldg(reinterpret_cast<const float4*>(z + (r + u) * ldz) + c4);
        if (MODE == 1) gv[u] = __ldg(reinterpret_cast<const float4*>(g + (r + u) * ldg) + c4);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 d = make_float4(zv[u].x - k.x, zv[u].y - k.y, zv[u].z - k.z, zv[u].w - k.w);
        if (MODE == 0) {
          sa.x += d.x; sa.y += d.y; sa.z += d.z; sa.w += d.w;
          sb.x = fmaf(d.x, d.x, sb.x); sb.y = fmaf(d.y, d.y, sb.y); sb.z = fmaf(d.z, d.z, sb.z); sb.w = fmaf(d.w, d.w, sb.w);
        } else {
          sa.x += gv[u].x; sa.y += gv[u].y; sa.z += gv[u].z; sa.w += gv[u].w;
          sb.x = fmaf(gv[u].x, d.x * rs.x, sb.x); sb.y = fmaf(gv[u].y, d.y * rs.y, sb.y);
          sb.z = fmaf(gv[u].z, d.z * rs.z, sb.z); sb.w = fmaf(gv[u].w, d.w * rs.w, sb.w);
        }
      }
    }
    for (; r < r1; ++r) {
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * ldz) + c4);
      const float4 d = make_float4(zv.x - k.x, zv.y - k.y, zv.z - k.z, zv.w - k.w);
      if (MODE == 0) {
        sa.x += d.x; sa.y += d.y; sa.z += d.z; sa.w += d.w;
        sb.x = fmaf(d.x, d.x, sb.x); sb.y = fmaf(d.y, d.y, sb.y); sb.z = fmaf(d.z, d.z, sb.z); sb.w = fmaf(d.w, d.w, sb.w);
      } else {
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + r * ldg) + c4);
        sa.x += gv.x; sa.y += gv.y; sa.z += gv.z; sa.w += gv.w;
        sb.x = fmaf(gv.x, d.x * rs.x, sb.x); sb.y = fmaf(gv.y, d.y * rs.y, sb.y);
        sb.z = fmaf(gv.z, d.z * rs.z, sb.z); sb.w = fmaf(gv.w, d.w * rs.w, sb.w);
      }
    }
    float* pa = partials + ((long long)blockIdx.x * 2) * N;
    reinterpret_cast<float4*>(pa)[c4] = sa;
    reinterpret_cast<float4*>(pa + N)[c4] = sb;
  }
}

// Fold the slab partials in a fixed order (fp64 accumulation) and finish.
//   MODE 0: stats[0..N) = mean, stats[N..2N) = rstd, affine[0..N) = gamma * rstd, affine[N..2N) = beta - mean * scale;
//           running_mean / running_var updated in place (momentum, unbiased variance) when given.
//   MODE 1: out[0..N) = sum g (d beta), out[N..2N) = sum g * xhat (d gamma).
//   amax != NULL: *amax is raised to a bound on the plane output that follows (the scale of an FP16X2 writer, common.cuh).
//   A normalised value obeys |xhat| <= sqrt(M) (sum xhat^2 <= M), and mean |xhat| <= 1, hence
//     MODE 0: |dropout(relu(gamma xhat + beta))| <= bound_mul * (|gamma_c| sqrt(M) + |beta_c|)
//     MODE 1: |gz| = |gamma rstd (g - mean g - xhat mean(g xhat))| <= |gamma_c rstd_c| * gmax * (2 + sqrt(M)),
//             gmax = *g_amax >= max |g| (from the producer of g: GEMM epilogue / rank-1 head bound)
//   The bounds overshoot the true maximum by ~sqrt(M) / 5; the two fp16 planes keep 22 bits for every element down to
//   2^-27 of the bound, so that costs nothing.
template <int MODE>
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partials, int nblk, int N, long long M,
                                                          const float* __restrict__ z_row0, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps, float momentum,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                                          float* __restrict__ stats, float* __restrict__ affine,
                                                          const float* __restrict__ fwd_stats, float bound_mul,
                                                          const float* __restrict__ g_amax, float* __restrict__ amax) {
  // 32 columns x 32 partial lanes per CTA: lane (ty) folds partials ty, ty + 32, ... of column tx in fp64 - the loads of
  // a warp are 32 consecutive columns of one partial row (one line; a warp per column made every load 32 sectors and
  // the kernel latency-bound at ~20 us) - then a fixed-order fold over ty in shared memory.
  // (the per-lane fold is a compensated fp32 sum - the fp64 pipe of this GPU made a plain double fold of 2 x 18 values
  // per thread cost 10 us - and only the final 32-term fold and the mean / variance arithmetic run in fp64)
  __shared__ double red_a[32][33], red_b[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float fa = 0.f, fb = 0.f, ca = 0.f, cb = 0.f;       // Kahan sums and their compensations
  if (c < N) {
#pragma unroll 4
    for (int b = ty; b < nblk; b += 32) {
      const float va = __ldg(partials + ((long long)b * 2) * N + c) - ca;
      const float vb = __ldg(partials + ((long long)b * 2 + 1) * N + c) - cb;
      const float ta = __fadd_rn(fa, va), tb = __fadd_rn(fb, vb);
      ca = __fsub_rn(__fsub_rn(ta, fa), va);
      cb = __fsub_rn(__fsub_rn(tb, fb), vb);
      fa = ta;
      fb = tb;
    }
  }
  red_a[ty][tx] = (double)fa - (double)ca;
  red_b[ty][tx] = (double)fb - (double)cb;
  __syncthreads();
  if (ty != 0 || c >= N) return;
  double sa = 0.0, sb = 0.0;
#pragma unroll
  for (int y = 0; y < 32; ++y) {
    sa += red_a[y][tx];
    sb += red_b[y][tx];
  }
  if (MODE == 0) {
    const double m = (double)M;
    const double dm = sa / m;                          // mean - k
    double var = sb / m - dm * dm;                     // biased variance of z (shift-invariant)
    if (var < 0.0) var = 0.0;
    const float mean = (float)((double)__ldg(z_row0 + c) + dm);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    stats[c] = mean;
    stats[N + c] = rstd;
    const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
    const float sc = ga * rstd;
    affine[c] = sc;
    affine[N + c] = be - mean * sc;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) {
      const float unbiased = (float)(var * (m / (m > 1.0 ? m - 1.0 : 1.0)));
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
    if (amax) {
      const float bound = bound_mul * (fabsf(ga) * sqrtf((float)M) * 1.001f + fabsf(be));
      if (bound > 0.f && bound < 3.0e38f) atomic_max_nonneg(amax, bound);
    }
  } else {
    stats[c] = (float)sa;
    stats[N + c] = (float)sb;
    if (amax) {
      const float rs = __ldg(fwd_stats + N + c), ga = gamma ? __ldg(gamma + c) : 1.f;
      const float bound = fabsf(ga * rs) * __ldg(g_amax) * (2.f + sqrtf((float)M)) * 1.001f;
      if (bound > 0.f && bound < 3.0e38f) atomic_max_nonneg(amax, bound);
    }
  }
}

// Per-32-row-group statistics a GEMM epilogue left behind (rsb_gemm_epilogue.bn_partials: parts[g][0][c] = S1 = sum (z - k_g),
// parts[g][1][c] = S2 = sum (z - k_g)^2 over the group's n_g valid rows, parts[g][2][c] = k_g, the group's first row)
// -> shifted sums with ONE common shift k_0 (group 0's), folded over a slice of the groups:
//     sum (z - k_0) = sum_g [n_g d_g + S1],   sum (z - k_0)^2 = sum_g [S2 + 2 d_g S1 + n_g d_g^2],   d_g = k_g - k_0
// (exact identities), written as out[slice][0..1][c] - the layout bn_finalize_kernel<0> folds, with k_0 as its row 0.
// Grid (N / 32, slices); 32 columns x 32 group lanes per CTA, coalesced loads, fixed-order fold over the lanes.
__global__ void __launch_bounds__(1024) bn_parts_fold_kernel(const float* __restrict__ parts, long long groups, int N, long long M,
                                                             float* __restrict__ out) {
  __shared__ float red_a[32][33], red_b[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long per = (groups + gridDim.y - 1) / gridDim.y;
  const long long g0 = per * blockIdx.y;
  long long g1 = g0 + per;
  if (g1 > groups) g1 = groups;
  float t1 = 0.f, t2 = 0.f;
  if (c < N) {
    const float k0 = __ldg(parts + 2 * (long long)N + c);
#pragma unroll 4
    for (long long g = g0 + ty; g < g1; g += 32) {
      long long n = M - g * 32;
      n = n > 32 ? 32 : n;
      if (n <= 0) break;
      const float* pg = parts + g * 3 * (long long)N;
      const float s1 = __ldg(pg + c), s2 = __ldg(pg + (long long)N + c), d = __ldg(pg + 2 * (long long)N + c) - k0;
      const float nd = (float)n * d;
      t1 += nd + s1;
      t2 += fmaf(d, 2.f * s1 + nd, s2);
    }
  }
  red_a[ty][tx] = t1;
  red_b[ty][tx] = t2;
  __syncthreads();
  if (ty != 0 || c >= N) return;
  t1 = 0.f;
  t2 = 0.f;
#pragma unroll
  for (int y = 0; y < 32; ++y) {
    t1 += red_a[y][tx];
    t2 += red_b[y][tx];
  }
  out[((long long)blockIdx.y * 2) * N + c] = t1;
  out[((long long)blockIdx.y * 2 + 1) * N + c] = t2;
}

// y = dropout_p(relu(z * scale + shift)) -> planes (+ ones column) + keep-and-positive mask; p = 0: no dropout
__global__ void __launch_bounds__(256) bn_apply_planes_kernel(const float* __restrict__ z, long long M, int N, long long ldz,
                                                              const float* __restrict__ affine, float drop_scale,
                                                              unsigned thr, unsigned long long seed, unsigned long long offset,
                                                              const unsigned long long* __restrict__ offset_dev,
                                                              uint16_t* __restrict__ out, long long out_ld,
                                                              long long plane_stride, unsigned char* __restrict__ mask,
                                                              int ones_col, PlaneFmt fmt, FastDiv row_div) {
  Philox rng{(unsigned)seed, (unsigned)(seed >> 32)};
  if (offset_dev) offset += *offset_dev;
  const float ps = fmt.scale();
  const int n4 = N / 4;
  const int n4o = (N + 7) / 8 * 2 + (ones_col ? 2 : 0);      // 4-column groups written per row (padding + ones column)
  const long long total = M * n4o;
  const bool small = total < (1ll << 32);                    // 32-bit multiply-high division instead of the 64-bit one
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = small ? (long long)fastdiv((unsigned)i, row_div) : i / n4o;
    const int c4 = (int)(i - r * n4o);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < n4) {
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * ldz) + c4);
      const float4 sc = __ldg(reinterpret_cast<const float4*>(affine) + c4);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(affine + N) + c4);
      const float4 y = make_float4(fmaf(zv.x, sc.x, sh.x), fmaf(zv.y, sc.y, sh.y), fmaf(zv.z, sc.z, sh.z), fmaf(zv.w, sc.w, sh.w));
      uint4 q = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (thr) q = rng(offset + (unsigned long long)(r * n4 + c4));
      uchar4 m;
      m.x = (y.x > 0.f) && (q.x >= thr); m.y = (y.y > 0.f) && (q.y >= thr);
      m.z = (y.z > 0.f) && (q.z >= thr); m.w = (y.w > 0.f) && (q.w >= thr);
      *reinterpret_cast<uchar4*>(mask + r * (long long)N + c4 * 4) = m;
      v = make_float4(m.x ? y.x * drop_scale : 0.f, m.y ? y.y * drop_scale : 0.f, m.z ? y.z * drop_scale : 0.f,
                      m.w ? y.w * drop_scale : 0.f);
    } else if (ones_col && c4 * 4 == ((N + 7) & ~7)) {
      v.x = 1.f;
    }
    const float vv[4] = {v.x, v.y, v.z, v.w};
    store_planes<4>(out + r * out_ld + c4 * 4, plane_stride, vv, fmt.format, ps);
  }
}

// gz = gamma * rstd * (g - d beta / M - xhat * d gamma / M) -> planes   (g already carries the ReLU / dropout mask)
__global__ void __launch_bounds__(256) bn_bwd_planes_kernel(const float* __restrict__ g, const float* __restrict__ z,
                                                            long long M, int N, long long ldg, long long ldz,
                                                            const float* __restrict__ stats, const float* __restrict__ sums,
                                                            const float* __restrict__ gamma, uint16_t* __restrict__ out,
                                                            long long out_ld, long long plane_stride, PlaneFmt fmt,
                                                            FastDiv row_div) {
  const float ps = fmt.scale();
  const int n4 = N / 4;
  const int n4o = (N + 7) / 8 * 2;
  const float inv_m = 1.0f / (float)M;
  const long long total = M * n4o;
  const bool small = total < (1ll << 32);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = small ? (long long)fastdiv((unsigned)i, row_div) : i / n4o;
    const int c4 = (int)(i - r * n4o);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < n4) {
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + r * ldg) + c4);
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * ldz) + c4);
      const float4 mu = __ldg(reinterpret_cast<const float4*>(stats) + c4);
      const float4 rs = __ldg(reinterpret_cast<const float4*>(stats + N) + c4);
      const float4 db = __ldg(reinterpret_cast<const float4*>(sums) + c4);
      const float4 dg = __ldg(reinterpret_cast<const float4*>(sums + N) + c4);
      const float4 ga = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
#define RSB_BN_BWD(f) (ga.f * rs.f * (gv.f - db.f * inv_m - (zv.f - mu.f) * rs.f * dg.f * inv_m))
      v = make_float4(RSB_BN_BWD(x), RSB_BN_BWD(y), RSB_BN_BWD(z), RSB_BN_BWD(w));
#undef RSB_BN_BWD
    }
    const float vv[4] = {v.x, v.y, v.z, v.w};
    store_planes<4>(out + r * out_ld + c4 * 4, plane_stride, vv, fmt.format, ps);
  }
}

// slabs per SM: the one-operand statistics pass likes 8 (0.051 -> 0.045 ms at 65536x400), the two-operand backward pass 4
static int bn_blocks(long long M, int per_sm = 8) {
  long long b = (long long)sm_count() * per_sm;
  const long long need = (M + 31) / 32;      // at least 32 rows per slab
  if (b > need) b = need;
  return b < 1 ? 1 : (int)b;
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int64_t rsb_bn_workspace_bytes(int64_t M, int32_t N) {
  if (M <= 0 || N <= 0) return 256;
  return (int64_t)bn_blocks(M) * 2 * N * 4 + 256;
}

extern "C" RSB_API int rsb_bn_train_fwd_stats(const float* z, int64_t M, int32_t N, int64_t ldz, const float* gamma,
                                              const float* beta, float eps, float momentum, float* running_mean,
                                              float* running_var, float* stats, float* affine, float bound_mul,
                                              float* act_amax, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!z || !stats || !affine || M <= 0 || N <= 0) return RSB_ERR_BAD_ARG;
  if (N % 4 || ldz % 4 || !aligned16(z) || !aligned16(stats) || !aligned16(affine) || (gamma && !aligned16(gamma)) ||
      (beta && !aligned16(beta)))
    return RSB_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < rsb_bn_workspace_bytes(M, N)) return RSB_ERR_WORKSPACE;
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = bn_blocks(M);
  bn_slab_sums_kernel<0><<<nblk, kBnThreads, 0, st>>>(z, nullptr, M, N, ldz, 0, nullptr, nullptr, partials);
  RSB_CHECK_LAUNCH();
  bn_finalize_kernel<0><<<(N + 31) / 32, 1024, 0, st>>>(partials, nblk, N, M, z, gamma, beta, eps, momentum, running_mean,
                                                     running_var, stats, affine, nullptr, bound_mul, nullptr, act_amax);
  RSB_CHECK_LAUNCH();
  note_launch(2);
  return RSB_OK;
}

constexpr int kBnPartSlices = 16;

extern "C" RSB_API int rsb_bn_finalize_partials(const float* parts, int64_t M, int32_t N, const float* gamma, const float* beta,
                                                float eps, float momentum, float* running_mean, float* running_var, float* stats,
                                                float* affine, float bound_mul, float* act_amax, void* workspace,
                                                int64_t workspace_bytes, void* stream) {
  if (!parts || !stats || !affine || M <= 0 || N <= 0) return RSB_ERR_BAD_ARG;
  if (!workspace || workspace_bytes < (int64_t)kBnPartSlices * 2 * N * 4 + 256) return RSB_ERR_WORKSPACE;
  float* sl = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long groups = (M + 127) / 128 * 4;
  int slices = kBnPartSlices;
  if (slices > groups) slices = (int)groups;
  bn_parts_fold_kernel<<<dim3((N + 31) / 32, slices), 1024, 0, st>>>(parts, groups, N, M, sl);
  RSB_CHECK_LAUNCH();
  // the slices are shifted sums w.r.t. group 0's first row (= row 0 of z): the ordinary finish applies
  bn_finalize_kernel<0><<<(N + 31) / 32, 1024, 0, st>>>(sl, slices, N, M, parts + 2 * (long long)N, gamma, beta, eps, momentum,
                                                     running_mean, running_var, stats, affine, nullptr, bound_mul, nullptr, act_amax);
  RSB_CHECK_LAUNCH();
  note_launch(2);
  return RSB_OK;
}

extern "C" RSB_API int rsb_bn_relu_dropout_planes(const float* z, int64_t M, int32_t N, int64_t ldz, const float* affine, float p,
                                                  uint64_t seed, uint64_t offset, const uint64_t* offset_dev, int32_t ones_col,
                                                  void* out_planes, int64_t out_ld, int64_t plane_stride, uint8_t* mask,
                                                  const rsb_planes_format* fmt, void* stream) {
  if (!z || !affine || !out_planes || !mask || M < 0 || N <= 0 || p < 0.f || p >= 1.f || !plane_fmt_ok(fmt)) return RSB_ERR_BAD_ARG;
  if (M == 0) return RSB_OK;
  if (N % 4 || ldz % 4 || out_ld % 8 || plane_stride % 8 || out_ld < ((N + 7) / 8) * 8 + (ones_col ? 8 : 0) || !aligned16(z) ||
      !aligned16(affine) || !aligned16(out_planes) || (reinterpret_cast<uintptr_t>(mask) & 3u))
    return RSB_ERR_UNSUPPORTED;
  const long long total = M * ((N + 7) / 8 * 2 + (ones_col ? 2 : 0));
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  bn_apply_planes_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      z, M, N, ldz, affine, 1.0f / (1.0f - p), (unsigned)(p * 4294967296.0), seed, offset,
      reinterpret_cast<const unsigned long long*>(offset_dev), reinterpret_cast<uint16_t*>(out_planes), out_ld, plane_stride,
      mask, ones_col, plane_fmt(fmt), make_fastdiv((unsigned long long)((N + 7) / 8 * 2 + (ones_col ? 2 : 0))));
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_bn_train_bwd_planes(const float* g, const float* z, int64_t M, int32_t N, int64_t ldg, int64_t ldz,
                                               const float* stats, const float* gamma, float* sums, void* out_planes,
                                               int64_t out_ld, int64_t plane_stride, const rsb_planes_format* fmt,
                                               const float* g_amax, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!g || !z || !stats || !sums || !out_planes || M <= 0 || N <= 0 || !plane_fmt_ok(fmt)) return RSB_ERR_BAD_ARG;
  if (fmt && fmt->format == kPlanesFp16x2 && !g_amax) return RSB_ERR_BAD_ARG;
  if (N % 4 || ldg % 4 || ldz % 4 || out_ld % 8 || plane_stride % 8 || out_ld < ((N + 7) / 8) * 8 || !aligned16(g) ||
      !aligned16(z) || !aligned16(stats) || !aligned16(sums) || !aligned16(out_planes) || (gamma && !aligned16(gamma)))
    return RSB_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < rsb_bn_workspace_bytes(M, N)) return RSB_ERR_WORKSPACE;
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nblk = bn_blocks(M, 4);
  const bool h = fmt && fmt->format == kPlanesFp16x2;
  bn_slab_sums_kernel<1><<<nblk, kBnThreads, 0, st>>>(z, g, M, N, ldz, ldg, stats, stats + N, partials);
  RSB_CHECK_LAUNCH();
  bn_finalize_kernel<1><<<(N + 31) / 32, 1024, 0, st>>>(partials, nblk, N, M, nullptr, gamma, nullptr, 0.f, 0.f, nullptr, nullptr,
                                                     sums, nullptr, stats, 1.f, h ? g_amax : nullptr, h ? fmt->amax : nullptr);
  RSB_CHECK_LAUNCH();
  const long long total = M * ((N + 7) / 8 * 2);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  bn_bwd_planes_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, z, M, N, ldg, ldz, stats, sums, gamma,
                                                         reinterpret_cast<uint16_t*>(out_planes), out_ld, plane_stride,
                                                         plane_fmt(fmt), make_fastdiv((unsigned long long)((N + 7) / 8 * 2)));
  RSB_CHECK_LAUNCH();
  note_launch(3);
  return RSB_OK;
}
