// Bandwidth-bound glue of a DCN-Mix cross layer (src/models/layer_dcn.py:8-24,90-115), fused so that
// between the three tensor-core GEMMs every [B, E*r] / [B, Dm] activation is read and written once:
//   gate_mix_fwd   g[b,e] = x_l[b,:].gates[e,:] ; H2 = tanh(P2) ; G2 = g[b,e] * H2[b,e,:]
//   cross_out_fwd  x_next = x_0 * (T0 + bias * sum_e g[b,e]) + x_l          (T0 = G2 @ U_cat)
//   cross_out_bwd  gT = g_next * x_0 ; gx0 = g_next * (T0 + bias*sg) ; dsg[b] = sum_d gT[b,d] bias[d]
//   gate_mix_bwd   gP2 = gG2 * g * (1 - H2^2) ; dg[b,e] = sum_k gG2 H2 + dsg[b] ;
//                  gxl = g_next + sum_e dg[b,e] * gates[e,:]
// One warp per sample row, 128-bit accesses, warp-shuffle reductions.  The gate is the identity
// (the reference default, layer_dcn.py:72-76); the softmax gate keeps the unfused path.
#include "common.cuh"

namespace rsb {

constexpr int kMaxExperts = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  return v;
}

__device__ __forceinline__ float pick(const float (&g)[kMaxExperts], int e) {
  float r = g[0];
#pragma unroll
  for (int i = 1; i < kMaxExperts; ++i) r = (e == i) ? g[i] : r;
  return r;
}

__global__ void __launch_bounds__(256) dcn_gate_mix_fwd_kernel(const float* __restrict__ p2,
                                                               const float* __restrict__ xl,
                                                               const float* __restrict__ gates, long long B,
                                                               int Dm, int E, int r, float* __restrict__ h2,
                                                               float* __restrict__ g_out, float* __restrict__ g2) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int dm4 = Dm / 4, er4 = E * r / 4, r4 = r / 4;
  for (long long b = warp; b < B; b += nw) {
    float g[kMaxExperts];
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e) g[e] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(xl + b * Dm);
    for (int i = lane; i < dm4; i += 32) {
      const float4 x = __ldg(xr + i);
#pragma unroll
      for (int e = 0; e < kMaxExperts; ++e) {
        if (e < E) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(gates + (long long)e * Dm) + i);
          g[e] += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e) g[e] = warp_sum(g[e]);
    if (lane < E) g_out[b * E + lane] = pick(g, lane);
    const float4* pr = reinterpret_cast<const float4*>(p2 + b * (long long)E * r);
    float4* hr = reinterpret_cast<float4*>(h2 + b * (long long)E * r);
    float4* gr = reinterpret_cast<float4*>(g2 + b * (long long)E * r);
    for (int i = lane; i < er4; i += 32) {
      float4 v = __ldg(pr + i);
      v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w);
      const float ge = pick(g, i / r4);
      hr[i] = v;
      gr[i] = make_float4(v.x * ge, v.y * ge, v.z * ge, v.w * ge);
    }
  }
}

__global__ void __launch_bounds__(256) dcn_cross_out_fwd_kernel(const float* __restrict__ t0,
                                                                const float* __restrict__ x0,
                                                                const float* __restrict__ xl,
                                                                const float* __restrict__ bias,
                                                                const float* __restrict__ g, long long B, int Dm,
                                                                int E, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int dm4 = Dm / 4;
  for (long long b = warp; b < B; b += nw) {
    float sg = 0.f;
    for (int e = 0; e < E; ++e) sg += __ldg(g + b * E + e);
    for (int i = lane; i < dm4; i += 32) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(t0 + b * Dm) + i);
      const float4 a = __ldg(reinterpret_cast<const float4*>(x0 + b * Dm) + i);
      const float4 l = __ldg(reinterpret_cast<const float4*>(xl + b * Dm) + i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + i);
      float4 o;
      o.x = a.x * (t.x + bb.x * sg) + l.x;
      o.y = a.y * (t.y + bb.y * sg) + l.y;
      o.z = a.z * (t.z + bb.z * sg) + l.z;
      o.w = a.w * (t.w + bb.w * sg) + l.w;
      reinterpret_cast<float4*>(out + b * Dm)[i] = o;
    }
  }
}

__global__ void __launch_bounds__(256) dcn_cross_out_bwd_kernel(const float* __restrict__ gn,
                                                                const float* __restrict__ x0,
                                                                const float* __restrict__ t0,
                                                                const float* __restrict__ bias,
                                                                const float* __restrict__ g, long long B, int Dm,
                                                                int E, float* __restrict__ gT,
                                                                float* __restrict__ gx0, float* __restrict__ dsg) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int dm4 = Dm / 4;
  for (long long b = warp; b < B; b += nw) {
    float sg = 0.f;
    for (int e = 0; e < E; ++e) sg += __ldg(g + b * E + e);
    float acc = 0.f;
    for (int i = lane; i < dm4; i += 32) {
      const float4 n = __ldg(reinterpret_cast<const float4*>(gn + b * Dm) + i);
      const float4 a = __ldg(reinterpret_cast<const float4*>(x0 + b * Dm) + i);
      const float4 t = __ldg(reinterpret_cast<const float4*>(t0 + b * Dm) + i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + i);
      const float4 gt = make_float4(n.x * a.x, n.y * a.y, n.z * a.z, n.w * a.w);
      reinterpret_cast<float4*>(gT + b * Dm)[i] = gt;
      reinterpret_cast<float4*>(gx0 + b * Dm)[i] =
          make_float4(n.x * (t.x + bb.x * sg), n.y * (t.y + bb.y * sg), n.z * (t.z + bb.z * sg),
                      n.w * (t.w + bb.w * sg));
      acc += gt.x * bb.x + gt.y * bb.y + gt.z * bb.z + gt.w * bb.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) dsg[b] = acc;
  }
}

__global__ void __launch_bounds__(256) dcn_gate_mix_bwd_kernel(const float* __restrict__ gg2,
                                                               const float* __restrict__ h2,
                                                               const float* __restrict__ g,
                                                               const float* __restrict__ dsg,
                                                               const float* __restrict__ gates,
                                                               const float* __restrict__ gn, long long B, int Dm,
                                                               int E, int r, float* __restrict__ gp2,
                                                               float* __restrict__ dg_out, float* __restrict__ gxl) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int dm4 = Dm / 4, er4 = E * r / 4, r4 = r / 4;
  for (long long b = warp; b < B; b += nw) {
    float gv[kMaxExperts], dg[kMaxExperts];
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e) {
      gv[e] = (e < E) ? __ldg(g + b * E + e) : 0.f;
      dg[e] = 0.f;
    }
    const float4* ar = reinterpret_cast<const float4*>(gg2 + b * (long long)E * r);
    const float4* hr = reinterpret_cast<const float4*>(h2 + b * (long long)E * r);
    float4* pr = reinterpret_cast<float4*>(gp2 + b * (long long)E * r);
    for (int i = lane; i < er4; i += 32) {
      const float4 a = __ldg(ar + i);
      const float4 h = __ldg(hr + i);
      const int e = i / r4;
      const float ge = pick(gv, e);
      const float d = a.x * h.x + a.y * h.y + a.z * h.z + a.w * h.w;
#pragma unroll
      for (int k = 0; k < kMaxExperts; ++k) dg[k] += (k == e) ? d : 0.f;
      pr[i] = make_float4(a.x * ge * (1.f - h.x * h.x), a.y * ge * (1.f - h.y * h.y), a.z * ge * (1.f - h.z * h.z),
                          a.w * ge * (1.f - h.w * h.w));
    }
    const float ds = __ldg(dsg + b);
#pragma unroll
    for (int e = 0; e < kMaxExperts; ++e) dg[e] = warp_sum(dg[e]) + ds;
    if (lane < E) dg_out[b * E + lane] = pick(dg, lane);
    for (int i = lane; i < dm4; i += 32) {
      float4 o = __ldg(reinterpret_cast<const float4*>(gn + b * Dm) + i);
#pragma unroll
      for (int e = 0; e < kMaxExperts; ++e) {
        if (e < E) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(gates + (long long)e * Dm) + i);
          o.x += dg[e] * w.x; o.y += dg[e] * w.y; o.z += dg[e] * w.z; o.w += dg[e] * w.w;
        }
      }
      reinterpret_cast<float4*>(gxl + b * Dm)[i] = o;
    }
  }
}

static unsigned row_grid(long long B) {
  long long blocks = (B * 32 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static int check_dims(int64_t B, int32_t Dm, int32_t E, int32_t r) {
  if (B < 0 || Dm <= 0 || E <= 0 || r <= 0) return RSB_ERR_BAD_ARG;
  if (Dm % 4 || r % 4 || E > kMaxExperts) return RSB_ERR_UNSUPPORTED;
  return RSB_OK;
}

}  // namespace rsb

using namespace rsb;

#define RSB_AL(p) if (!aligned16(p)) return RSB_ERR_UNSUPPORTED

extern "C" RSB_API int rsb_dcn_gate_mix_fwd(const float* p2, const float* xl, const float* gates, int64_t B,
                                            int32_t Dm, int32_t E, int32_t r, float* h2, float* g, float* g2,
                                            void* stream) {
  int rc = check_dims(B, Dm, E, r);
  if (rc) return rc;
  if (B == 0) return RSB_OK;
  if (!p2 || !xl || !gates || !h2 || !g || !g2) return RSB_ERR_BAD_ARG;
  RSB_AL(p2); RSB_AL(xl); RSB_AL(gates); RSB_AL(h2); RSB_AL(g2);
  dcn_gate_mix_fwd_kernel<<<row_grid(B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p2, xl, gates, B, Dm, E, r,
                                                                                            h2, g, g2);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_dcn_cross_out_fwd(const float* t0, const float* x0, const float* xl, const float* bias,
                                             const float* g, int64_t B, int32_t Dm, int32_t E, float* x_next,
                                             void* stream) {
  int rc = check_dims(B, Dm, E, 4);
  if (rc) return rc;
  if (B == 0) return RSB_OK;
  if (!t0 || !x0 || !xl || !bias || !g || !x_next) return RSB_ERR_BAD_ARG;
  RSB_AL(t0); RSB_AL(x0); RSB_AL(xl); RSB_AL(bias); RSB_AL(x_next);
  dcn_cross_out_fwd_kernel<<<row_grid(B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t0, x0, xl, bias, g, B, Dm,
                                                                                             E, x_next);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_dcn_cross_out_bwd(const float* g_next, const float* x0, const float* t0, const float* bias,
                                             const float* g, int64_t B, int32_t Dm, int32_t E, float* gT, float* gx0,
                                             float* dsg, void* stream) {
  int rc = check_dims(B, Dm, E, 4);
  if (rc) return rc;
  if (B == 0) return RSB_OK;
  if (!g_next || !x0 || !t0 || !bias || !g || !gT || !gx0 || !dsg) return RSB_ERR_BAD_ARG;
  RSB_AL(g_next); RSB_AL(x0); RSB_AL(t0); RSB_AL(bias); RSB_AL(gT); RSB_AL(gx0);
  dcn_cross_out_bwd_kernel<<<row_grid(B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g_next, x0, t0, bias, g, B,
                                                                                             Dm, E, gT, gx0, dsg);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_dcn_gate_mix_bwd(const float* g_g2, const float* h2, const float* g, const float* dsg,
                                            const float* gates, const float* g_next, int64_t B, int32_t Dm,
                                            int32_t E, int32_t r, float* g_p2, float* dg, float* g_xl, void* stream) {
  int rc = check_dims(B, Dm, E, r);
  if (rc) return rc;
  if (B == 0) return RSB_OK;
  if (!g_g2 || !h2 || !g || !dsg || !gates || !g_next || !g_p2 || !dg || !g_xl) return RSB_ERR_BAD_ARG;
  RSB_AL(g_g2); RSB_AL(h2); RSB_AL(gates); RSB_AL(g_next); RSB_AL(g_p2); RSB_AL(g_xl);
  dcn_gate_mix_bwd_kernel<<<row_grid(B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g_g2, h2, g, dsg, gates, g_next, B, Dm, E, r, g_p2, dg, g_xl);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}
