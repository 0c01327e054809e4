// Shared device helpers for librsb (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "rsb.h"

#define RSB_CHECK_LAUNCH()                         \
  do {                                             \
    cudaError_t _e = cudaGetLastError();           \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

namespace rsb {

// Exact unsigned 32-bit division by a divisor fixed for the whole launch (Granlund & Montgomery, "Division by
// invariant integers using multiplication", fig. 4.1, N = 32): q = (t + ((n - t) >> s1)) >> s2 with t = umulhi(m, n).
// Five instructions instead of the ~20 of the emulated 32-bit divide; bit-exact for every n < 2^32, d >= 1.
struct FastDiv {
  unsigned d, m, s1, s2;
};
__host__ __device__ inline FastDiv make_fastdiv(unsigned long long d64) {
  FastDiv f;
  unsigned d = (unsigned)d64;
  if (d == 0) d = 1;
  int L = 0;
  while ((1ull << L) < d) ++L;                       // ceil(log2 d)
  f.d = d;
  f.m = (unsigned)(((1ull << 32) * ((1ull << L) - d)) / d + 1);
  f.s1 = L < 1 ? L : 1;
  f.s2 = L > 1 ? L - 1 : 0;
  return f;
}
__device__ __forceinline__ unsigned fastdiv(unsigned n, const FastDiv& f) {
  const unsigned t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

// Counter-based Philox4x32-10 stream keyed by (seed, offset): the dropout mask of element group `ctr` (4 consecutive
// elements of a row-major activation) is a pure function of the counter, so every kernel that needs the mask - the
// stand-alone ReLU+dropout pass, the row-dot variant, the GEMM epilogue - draws the same one.
struct Philox {
  unsigned k0, k1;
  __device__ __forceinline__ uint4 operator()(unsigned long long ctr) const {
    unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
    unsigned a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a;
      c1 = lo1;
      c2 = hi0 ^ c3 ^ b;
      c3 = lo0;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// ---- operand formats of the tensor-core GEMM (gemm/planes_gemm.cu) ----
// BF16X3: x = x0 + x1 + x2, three bf16 planes (3 x 8 mantissa bits: every fp32 value exactly), 6 plane products.
// FP16X2: x * s = h0 + h1, two fp16 planes (2 x 11 bits) of the matrix scaled by a power of two s chosen from a bound
//         `amax` >= max |x| so that amax * s lies in [2^13, 2^14): the big elements keep 22 bits, every element is
//         exact to 2^-25 / s (the fp16 subnormal step), i.e. to 2^-38 of the bound; 3 plane products (h1 * h1' <=
//         2^-22 |a||b| is dropped).  Writer and reader derive s from the same device scalar with plane_scale().
constexpr int kPlanesBf16x3 = 0, kPlanesFp16x2 = 1;
constexpr int kPlaneTopExp = 14;
__host__ __device__ __forceinline__ float plane_scale(float amax, int max_exp) {
  if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;      // empty / all-zero / non-finite: unscaled
  int e;
  frexpf(amax, &e);                                     // amax in [2^(e-1), 2^e)
  int se = kPlaneTopExp - e;
  if (se > max_exp) se = max_exp;
  if (se < -100) se = -100;
  return ldexpf(1.f, se);
}
// non-negative floats order like their bit patterns: a max over a grid is one integer atomic per warp
__device__ __forceinline__ void atomic_max_nonneg(float* slot, float v) {
  atomicMax(reinterpret_cast<unsigned*>(slot), __float_as_uint(v));
}

__device__ __forceinline__ void split3_bf16(float x, __nv_bfloat16& h0, __nv_bfloat16& h1, __nv_bfloat16& h2) {
  h0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h0);            // exact in fp32
  h1 = __float2bfloat16_rn(r1);
  h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));  // exact
}
__device__ __forceinline__ void split2_fp16(float xs, __half& h0, __half& h1) {
  h0 = __float2half_rn(xs);
  h1 = __float2half_rn(xs - __half2float(h0));          // the remainder is exact in fp32
}
// N (4 or 8) consecutive values of one row -> one 8- or 16-byte store per plane.  `dst` points at plane 0 (16-bit
// elements of either format), `s` is the FP16X2 scale (ignored for BF16X3).  The 16-bit results are packed into 32-bit
// words by hand (no local arrays the compiler might leave in local memory).
template <int N>
__device__ __forceinline__ void store_planes(void* dst, long long plane_stride, const float* v, int format, float s) {
  static_assert(N == 4 || N == 8, "4 or 8 values");
  uint32_t w0[N / 2], w1[N / 2], w2[N / 2];
  uint16_t* o = reinterpret_cast<uint16_t*>(dst);
  if (format == kPlanesFp16x2) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      __half a0, a1, b0, b1;
      split2_fp16(v[j] * s, a0, a1);
      split2_fp16(v[j + 1] * s, b0, b1);
      w0[j / 2] = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(b0) << 16);
      w1[j / 2] = (uint32_t)__half_as_ushort(a1) | ((uint32_t)__half_as_ushort(b1) << 16);
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      __nv_bfloat16 a0, a1, a2, b0, b1, b2;
      split3_bf16(v[j], a0, a1, a2);
      split3_bf16(v[j + 1], b0, b1, b2);
      w0[j / 2] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(b0) << 16);
      w1[j / 2] = (uint32_t)__bfloat16_as_ushort(a1) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
      w2[j / 2] = (uint32_t)__bfloat16_as_ushort(a2) | ((uint32_t)__bfloat16_as_ushort(b2) << 16);
    }
  }
  if constexpr (N == 8) {
    *reinterpret_cast<uint4*>(o) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    *reinterpret_cast<uint4*>(o + plane_stride) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    if (format != kPlanesFp16x2) *reinterpret_cast<uint4*>(o + 2 * plane_stride) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
  } else {
    *reinterpret_cast<uint2*>(o) = make_uint2(w0[0], w0[1]);
    *reinterpret_cast<uint2*>(o + plane_stride) = make_uint2(w1[0], w1[1]);
    if (format != kPlanesFp16x2) *reinterpret_cast<uint2*>(o + 2 * plane_stride) = make_uint2(w2[0], w2[1]);
  }
}
struct PlaneFmt {
  int format;            // kPlanesBf16x3 / kPlanesFp16x2
  int max_exp;           // FP16X2: the scale is at most 2^max_exp
  const float* amax;     // FP16X2: device scalar, bound on |x|
  __device__ __forceinline__ float scale() const {
    return format == kPlanesFp16x2 ? plane_scale(amax ? *amax : 0.f, max_exp) : 1.f;
  }
};
inline PlaneFmt plane_fmt(const rsb_planes_format* f) {
  PlaneFmt r;
  r.format = f ? f->format : kPlanesBf16x3;
  r.max_exp = f ? f->max_scale_exp : 0;
  r.amax = f ? f->amax : nullptr;
  return r;
}
inline bool plane_fmt_ok(const rsb_planes_format* f) {
  return !f || f->format == kPlanesBf16x3 || (f->format == kPlanesFp16x2 && f->amax != nullptr);
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// A row chunk of V floats held in registers. V = 4 -> one 128-bit access.
template <int V>
struct FV {
  float v[V];
  __device__ __forceinline__ static FV zero() {
    FV r;
#pragma unroll
    for (int i = 0; i < V; ++i) r.v[i] = 0.f;
    return r;
  }
};

template <int V>
__device__ __forceinline__ FV<V> ldg(const float* p);
template <>
__device__ __forceinline__ FV<4> ldg<4>(const float* p) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  FV<4> r;
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}
template <>
__device__ __forceinline__ FV<1> ldg<1>(const float* p) {
  FV<1> r;
  r.v[0] = __ldg(p);
  return r;
}
// plain (coherent) load: for tables that this same kernel also writes
template <int V>
__device__ __forceinline__ FV<V> ld(const float* p);
template <>
__device__ __forceinline__ FV<4> ld<4>(const float* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  FV<4> r;
  r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}
template <>
__device__ __forceinline__ FV<1> ld<1>(const float* p) {
  FV<1> r;
  r.v[0] = *p;
  return r;
}
template <int V>
__device__ __forceinline__ void st(float* p, const FV<V>& x);
template <>
__device__ __forceinline__ void st<4>(float* p, const FV<4>& x) {
  *reinterpret_cast<float4*>(p) = make_float4(x.v[0], x.v[1], x.v[2], x.v[3]);
}
template <>
__device__ __forceinline__ void st<1>(float* p, const FV<1>& x) {
  *p = x.v[0];
}
// streaming store: written once, not re-read by this kernel
template <int V>
__device__ __forceinline__ void st_cs(float* p, const FV<V>& x);
template <>
__device__ __forceinline__ void st_cs<4>(float* p, const FV<4>& x) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
}
template <>
__device__ __forceinline__ void st_cs<1>(float* p, const FV<1>& x) {
  __stcs(p, x.v[0]);
}

template <int V>
__device__ __forceinline__ FV<V> shfl_xor(const FV<V>& x, int off) {
  FV<V> r;
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = __shfl_xor_sync(kFull, x.v[i], off);
  return r;
}

__device__ __forceinline__ float sigmoidf_exact(float s) { return 1.0f / (1.0f + expf(-s)); }

// Lanes-per-row for a row of `chunks` vector chunks: next power of two, <= 32.
inline int lanes_per_row(int chunks) {
  int l = 1;
  while (l < chunks) l <<= 1;
  return l;
}

struct RowShape {
  int V;    // floats per lane access (4 or 1)
  int LPR;  // lanes per row (power of two)
  bool ok;
};
inline RowShape row_shape(int E, bool aligned16) {
  RowShape s;
  s.ok = true;
  if (E > 0 && E % 4 == 0 && aligned16 && E <= 128) {
    s.V = 4;
    s.LPR = lanes_per_row(E / 4);
  } else if (E > 0 && E <= 32) {
    s.V = 1;
    s.LPR = lanes_per_row(E);
  } else {
    s.V = 0; s.LPR = 0; s.ok = false;
  }
  return s;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();
void note_launch(int n);

// Dispatch helper: calls F.template run<V, LPR>() for the runtime shape.
#define RSB_DISPATCH_SHAPE(shape, CALL)                       \
  do {                                                        \
    if ((shape).V == 4) {                                     \
      switch ((shape).LPR) {                                  \
        case 1: { CALL(4, 1); } break;                        \
        case 2: { CALL(4, 2); } break;                        \
        case 4: { CALL(4, 4); } break;                        \
        case 8: { CALL(4, 8); } break;                        \
        case 16: { CALL(4, 16); } break;                      \
        default: { CALL(4, 32); } break;                      \
      }                                                       \
    } else {                                                  \
      switch ((shape).LPR) {                                  \
        case 1: { CALL(1, 1); } break;                        \
        case 2: { CALL(1, 2); } break;                        \
        case 4: { CALL(1, 4); } break;                        \
        case 8: { CALL(1, 8); } break;                        \
        case 16: { CALL(1, 16); } break;                      \
        default: { CALL(1, 32); } break;                      \
      }                                                       \
    }                                                         \
  } while (0)

}  // namespace rsb
