// Fused gather (+ lightweight-embedding transform) + FM forward, and the per-lookup
// backward (stage 1).  One warp per sample; a row of E floats is spread over LPR lanes
// with 128-bit accesses (V = 4) so a warp touches 32/LPR random rows per instruction and
// every 32-byte sector it requests is fully used.  Field-axis reductions are xor-shuffles
// over the lane groups.
//
// Reference semantics (paths relative to /root/reference):
//   offsets add          src/models/deepfm.py:88, src/models/dcn.py:84
//   vanilla gather       src/models/embeddings/base.py:75
//   QR                   src/models/embeddings/qr_embedding.py:95-109
//   PEP soft threshold   src/models/embeddings/pep_embedding.py:82-92
//   retrain masks        pep_embedding.py:215-221, deepfm_opt_embed.py:693-706
//   OptEmbed supernet    deepfm_opt_embed.py:219-226, optembed_utils.py:25-44,101-104
//   first order + FM     src/models/deepfm.py:49,91-98
#include <stdlib.h>

#include "common.cuh"

namespace rsb {

struct LookupArgs {
  const void* idx;
  int idx_i32;
  const long long* offsets;
  const long long* rows_in;  // backward: saved global rows
  long long B;
  int F, VF, E;
  const float* table;
  long long n_rows, n_global;
  const float* table1;
  long long divider;
  long long modulus;  // i1 = row % modulus (== divider for QR; the CERP bucket size otherwise)
  int small32;
  FastDiv fd_div, fd_mod, fd_g;   // invariant-divisor division for the small32 paths (divider, modulus, G)
  const void* aux;
  int aux_mode;
  const long long* mask_d;
  const float* fc;
  const float* bias;
  // forward outputs
  float* out_emb;
  float* out_y;
  float* out_sum;
  long long* out_rows;
  int* err;
  float* amax_slots;   // optional [RSB_LOOKUP_AMAX_SLOTS]: raised to max |out_emb| (one integer atomic per warp)
  // backward inputs / outputs
  const float* emb;
  const float* S;
  const float* g_y;
  const float* g_deep;
  float* rg_main;
  float* rg_aux;
  float* fc_grad;
  // row-sharded tables (VANILLA only): shard g holds rows r with r % G == g at r / G
  const float* const* table_shards;
  const void* const* aux_shards;   // per-row aux arrays (PEP s of the feature / feature_dim kinds, retrain masks) sharded like the table
  int G;
  // small fields replicated on every rank instead of sharded: hot_map[f] = (lo, hi, delta) int64 - a row of field f
  // inside [lo, hi) lives at hot_table[row + delta]; hi == lo for a sharded field
  const float* hot_table;
  const long long* hot_map;
  // QR emb1 gradient accumulated in registers (divider <= kTinyRows): per-CTA partials [grid][kTinyRows][E]
  float* tiny_partials;
};

constexpr int kTinyRows = 8;

__device__ __forceinline__ void shard_split(const LookupArgs& a, long long row, int& owner, long long& lrow) {
  if (a.small32) {
    unsigned r = (unsigned)row, g = (unsigned)a.G;
    unsigned q = fastdiv(r, a.fd_g);
    lrow = q;
    owner = (int)(r - q * g);
  } else {
    lrow = row / a.G;
    owner = (int)(row - lrow * a.G);
  }
}

__device__ __forceinline__ void qr_split(const LookupArgs& a, long long row, long long& i1, long long& i2) {
  if (a.small32) {
    unsigned r = (unsigned)row, d = (unsigned)a.divider;
    unsigned q = fastdiv(r, a.fd_div);
    i2 = q;
    i1 = (a.modulus == a.divider) ? r - q * d : r - fastdiv(r, a.fd_mod) * (unsigned)a.modulus;
  } else {
    i2 = row / a.divider;
    i1 = (a.modulus == a.divider) ? row - i2 * a.divider : row % a.modulus;
  }
}

// Where a row of the MAIN table (and of a per-row aux array) lives: the one local table, or (owner shard, local row).
struct RowLoc {
  int owner;
  long long lrow;
};
__device__ __forceinline__ RowLoc locate(const LookupArgs& a, long long row) {
  RowLoc l;
  l.owner = 0;
  l.lrow = row;
  if (a.table_shards != nullptr) shard_split(a, row, l.owner, l.lrow);
  return l;
}
__device__ __forceinline__ const float* main_row(const LookupArgs& a, RowLoc l) {
  return (a.table_shards != nullptr ? a.table_shards[l.owner] : a.table) + l.lrow * a.E;   // a peer shard: the load crosses NVLink
}
__device__ __forceinline__ const void* aux_base(const LookupArgs& a, RowLoc l) {
  return a.aux_shards != nullptr ? a.aux_shards[l.owner] : a.aux;
}

template <int LPR>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int off = 1; off < LPR; off <<= 1) x += __shfl_xor_sync(kFull, x, off);
  return x;
}

__device__ __forceinline__ float pep_s(const LookupArgs& a, RowLoc l, int d) {
  switch (a.aux_mode) {
    case RSB_PEP_GLOBAL: return __ldg(reinterpret_cast<const float*>(a.aux));
    case RSB_PEP_DIMENSION: return __ldg(reinterpret_cast<const float*>(a.aux) + d);
    case RSB_PEP_FEATURE: return __ldg(reinterpret_cast<const float*>(aux_base(a, l)) + l.lrow);
    default: return __ldg(reinterpret_cast<const float*>(aux_base(a, l)) + l.lrow * a.E + d);
  }
}

// BinaryStep surrogate gradient (optembed_utils.py:35-44)
__device__ __forceinline__ float binary_step_grad(float z) {
  float az = fabsf(z);
  if (az > 1.0f) return 0.0f;
  if (az > 0.4f) return 0.4f;
  return 2.0f - 4.0f * az;
}

// Transformed row chunk for lookup (b, vf).  Must be called by ALL lanes of the warp
// (OPTEMBED reduces over the lane group); `act` = this lane owns a valid (row, chunk).
template <int K, int V, int LPR>
__device__ __forceinline__ FV<V> load_transformed(const LookupArgs& a, long long row, long long b, int f, int vf,
                                                  int c, bool act) {
  FV<V> e = FV<V>::zero();
  const int d0 = c * V;
  if (K == RSB_KIND_VANILLA) {
    if (act) {
      if (a.table_shards != nullptr) {
        bool hot = false;
        if (a.hot_map != nullptr) {
          const long long lo = __ldg(a.hot_map + 3 * f), hi = __ldg(a.hot_map + 3 * f + 1);
          hot = row >= lo && row < hi;
          if (hot) e = ldg<V>(a.hot_table + (row + __ldg(a.hot_map + 3 * f + 2)) * a.E + d0);
          else if (hi > lo && a.err) *a.err = 1;               // an id outside its own (replicated) field
        }
        if (!hot) {
          int owner;
          long long lrow;
          shard_split(a, row, owner, lrow);
          e = ldg<V>(a.table_shards[owner] + lrow * a.E + d0);   // peer shard: the load crosses NVLink
        }
      } else {
        e = ldg<V>(a.table + row * a.E + d0);
      }
    }
  } else if (K == RSB_KIND_QR_MULT || K == RSB_KIND_QR_ADD) {
    if (act) {
      long long i1, i2;
      qr_split(a, row, i1, i2);
      FV<V> e1 = ldg<V>(a.table1 + i1 * a.E + d0);
      FV<V> e2 = ldg<V>(main_row(a, locate(a, i2)) + d0);
#pragma unroll
      for (int i = 0; i < V; ++i) e.v[i] = (K == RSB_KIND_QR_MULT) ? e1.v[i] * e2.v[i] : e1.v[i] + e2.v[i];
    }
  } else if (K == RSB_KIND_QR_CAT) {
    if (act) {
      long long i1, i2;
      qr_split(a, row, i1, i2);
      e = (vf < a.F) ? ldg<V>(a.table1 + i1 * a.E + d0) : ldg<V>(main_row(a, locate(a, i2)) + d0);
    }
  } else if (K == RSB_KIND_PEP) {
    if (act) {
      const RowLoc l = locate(a, row);
      FV<V> w = ldg<V>(main_row(a, l) + d0);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float sg = sigmoidf_exact(pep_s(a, l, d0 + i));
        float m = fmaxf(fabsf(w.v[i]) - sg, 0.0f);
        e.v[i] = (w.v[i] > 0.f) ? m : ((w.v[i] < 0.f) ? -m : 0.0f * m);
      }
    }
  } else if (K == RSB_KIND_MASK) {
    if (act) {
      const RowLoc l = locate(a, row);
      FV<V> w = ldg<V>(main_row(a, l) + d0);
      const unsigned char* m = reinterpret_cast<const unsigned char*>(aux_base(a, l)) + l.lrow * a.E + d0;
#pragma unroll
      for (int i = 0; i < V; ++i) e.v[i] = w.v[i] * (float)(m[i] != 0);
    }
  } else if (K == RSB_KIND_OPTEMBED) {
    FV<V> w = FV<V>::zero();
    if (act) w = ldg<V>(main_row(a, locate(a, row)) + d0);
    float keep = 1.0f;
    if (a.aux != nullptr) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) part += (a.aux_mode == 2) ? w.v[i] * w.v[i] : fabsf(w.v[i]);
      float nrm = group_sum<LPR>(part);
      if (a.aux_mode == 2) nrm = sqrtf(nrm);
      float t = act ? __ldg(reinterpret_cast<const float*>(a.aux) + f) : 0.f;
      keep = (nrm - t > 0.0f) ? 1.0f : 0.0f;
    }
    long long k = (a.mask_d != nullptr && act) ? __ldg(a.mask_d + b * a.F + f) : (long long)a.E;
#pragma unroll
    for (int i = 0; i < V; ++i) e.v[i] = ((long long)(d0 + i) <= k) ? w.v[i] * keep : 0.0f * w.v[i];
  }
  return e;
}

__device__ __forceinline__ long long load_id(const LookupArgs& a, long long b, int f) {
  long long id = a.idx_i32 ? (long long)__ldg(reinterpret_cast<const int*>(a.idx) + b * a.F + f)
                           : __ldg(reinterpret_cast<const long long*>(a.idx) + b * a.F + f);
  if (a.offsets) id += __ldg(a.offsets + f);
  return id;
}

// Field iterations are processed in batches of kIter: all ids of a batch are loaded first, then all
// first-order weights and table rows, then the arithmetic and the stores.  (A plain per-field loop
// serialises id -> row -> store round trips, because the compiler may not hoist the next id load
// above the previous emb store; ncu showed the warps stalled on exactly those three loads.)
constexpr int kIter = 5;   // 8 lane groups x 5 = 40 lookups per batch: the 39 Criteo fields in ONE round of dependent loads

// W = lanes that share one sample (32, 16 or 8: a warp works on 32 / W samples at once), KI = field iterations whose
// loads are issued together.  One sample per warp with KI = 5 gives 8 lane groups x 5 = 40 lookup slots - the 39 Criteo
// fields in one round - but leaves 29 of 40 slots idle on a KDD-shaped batch (11 fields): there two samples share a
// warp (W = 16) with KI = 3 (12 slots), and the Avazu shape (22 fields) takes W = 32, KI = 3 (24 slots).
template <int K, int V, int LPR, int W = kWarp, int KI = kIter>
__global__ void __launch_bounds__(256) lookup_fwd_kernel(LookupArgs a) {
  constexpr int GPW = W / LPR;          // lane groups per sample
  constexpr int SPW = kWarp / W;        // samples per warp
  static_assert(W >= LPR && GPW >= 1, "a row must fit the lanes of one sample");
  const int lane = threadIdx.x & 31;
  const int sl = lane % W, sw = lane / W;
  const int g = sl / LPR, c = sl % LPR;
  const bool cact = c * V < a.E;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool fm = a.out_y != nullptr || a.out_sum != nullptr;
  float emax = 0.f;

  for (long long b0 = warp * SPW; b0 < a.B; b0 += nwarps * SPW) {
    const long long b = b0 + sw;
    const bool bvalid = b < a.B;      // the lanes of a sample past the end still take part in the shuffles
    FV<V> S = FV<V>::zero(), Q = FV<V>::zero();
    float first = 0.f;
    for (int vf0 = 0; vf0 < a.VF; vf0 += GPW * KI) {
      long long rowv[KI];
      int fv[KI];
      bool vactv[KI];
#pragma unroll
      for (int it = 0; it < KI; ++it) {
        const int vf = vf0 + it * GPW + g;
        vactv[it] = bvalid && vf < a.VF;
        fv[it] = (K == RSB_KIND_QR_CAT && vf >= a.F) ? vf - a.F : vf;
        rowv[it] = vactv[it] ? load_id(a, b, fv[it]) : 0;
      }
      float fcv[KI];
#pragma unroll
      for (int it = 0; it < KI; ++it) {
        const int vf = vf0 + it * GPW + g;
        fcv[it] = 0.f;
        if (vactv[it]) {
          if (rowv[it] < 0 || rowv[it] >= a.n_global) {
            if (a.err) *a.err = 1;
            rowv[it] = 0;
          }
          if (c == 0 && vf < a.F) {
            if (a.out_rows) a.out_rows[b * a.F + fv[it]] = rowv[it];
            if (a.fc) fcv[it] = __ldg(a.fc + rowv[it]);
          }
        }
      }
      FV<V> ev[KI];
#pragma unroll
      for (int it = 0; it < KI; ++it) {
        const int vf = vf0 + it * GPW + g;
        ev[it] = load_transformed<K, V, LPR>(a, rowv[it], b, fv[it], vf, c, vactv[it] && cact);
      }
#pragma unroll
      for (int it = 0; it < KI; ++it) {
        const int vf = vf0 + it * GPW + g;
        first += fcv[it];
        if (vactv[it] && cact) {
          st<V>(a.out_emb + ((b * a.VF + vf) * (long long)a.E + c * V), ev[it]);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            S.v[i] += ev[it].v[i];
            Q.v[i] = fmaf(ev[it].v[i], ev[it].v[i], Q.v[i]);
            emax = fmaxf(emax, fabsf(ev[it].v[i]));
          }
        }
      }
    }
    if (fm) {
#pragma unroll
      for (int off = LPR; off < W; off <<= 1) {
        FV<V> s2 = shfl_xor<V>(S, off), q2 = shfl_xor<V>(Q, off);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          S.v[i] += s2.v[i];
          Q.v[i] += q2.v[i];
        }
      }
      float y2 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) y2 += S.v[i] * S.v[i] - Q.v[i];
      y2 = group_sum<LPR>(y2);
#pragma unroll
      for (int off = 1; off < W; off <<= 1) first += __shfl_xor_sync(kFull, first, off);
      if (a.out_sum && bvalid && g == 0 && cact) st<V>(a.out_sum + b * a.E + c * V, S);
      if (a.out_y && bvalid && sl == 0) {
        float x1 = first + (a.bias ? __ldg(a.bias) : 0.f);
        a.out_y[b] = x1 + 0.5f * y2;
      }
    }
  }
  if (a.amax_slots != nullptr) {
    // max |emb| of this warp's samples -> one of the slots (the dense tail's FP16X2 split takes the max over them
    // instead of re-reading the whole activation: rsb_absmax over 1024 floats)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) emax = fmaxf(emax, __shfl_xor_sync(kFull, emax, off));
    if (lane == 0 && emax > 0.f && emax < 3.0e38f)
      atomic_max_nonneg(a.amax_slots + (unsigned)(warp % RSB_LOOKUP_AMAX_SLOTS), emax);
  }
}

// ---------------------------------------------------------------------------------
// Backward stage 1
// ---------------------------------------------------------------------------------
// First-order weight gradient fc_grad[row] += g_y[b] without hot-address atomics: a warp takes ONE
// field of 32 consecutive samples, lanes holding the same row (small fields: a 4-value field is hit
// by every sample) are combined with match_any + a masked warp sum, one atomic per distinct row.
// (Per-lookup atomics serialise ~B updates on each hot row inside L2 and stall the whole gather.)
__global__ void __launch_bounds__(256) fc_grad_kernel(const long long* __restrict__ rows,
                                                      const float* __restrict__ g_y, long long B, int F,
                                                      float* __restrict__ fc_grad) {
  extern __shared__ long long tile[];  // [32][F]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long b0 = (long long)blockIdx.x * 32;
  const long long nvalid = (B - b0 < 32) ? (B - b0) : 32;
  for (long long i = threadIdx.x; i < nvalid * F; i += blockDim.x) tile[i] = __ldg(rows + b0 * F + i);
  __syncthreads();
  const bool valid = lane < nvalid;
  const float gy = valid ? __ldg(g_y + b0 + lane) : 0.f;
  for (int f = warp; f < F; f += 8) {
    const long long row = valid ? tile[(long long)lane * F + f] : (long long)(-1 - lane);
    const unsigned peers = __match_any_sync(kFull, row);
    const int leader = __ffs(peers) - 1;
    float sum = gy;
    // every lane sums its peer set in lane order; the trip count is the size of the LARGEST class in the warp
    // (a 4-value field: ~12 of 32 samples, a 100-value field: ~3, a big field: 1 = no loop at all)
    const int maxc = __reduce_max_sync(kFull, __popc(peers));
    if (maxc > 1) {
      sum = 0.f;
      unsigned m = peers;
      for (int k = 0; k < maxc; ++k) {
        const int src = m ? (__ffs(m) - 1) : lane;
        const float v = __shfl_sync(kFull, gy, src);
        sum += m ? v : 0.f;
        m &= m - 1;
      }
    }
    if (valid && lane == leader) atomicAdd(fc_grad + row, sum);
  }
}

// Sum of the per-CTA partial tables: 32 columns x 32 row lanes per CTA, coalesced along the columns,
// fixed-order fold over the row lanes.
__global__ void __launch_bounds__(1024) partials_reduce_kernel(const float* __restrict__ partials, int nblk, int N,
                                                               float* __restrict__ dst) {
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
    dst[c] = t;
  }
}

// TR = number of register accumulators of the TINY variant (emb1 rows actually used, rounded up to 5 or 8)
// W / KN: lanes per sample and field iterations in flight, as in lookup_fwd_kernel (KN applies to the non-TINY variant)
template <int K, int V, int LPR, bool TINY = false, int KT = 2, int TR = kTinyRows, int W = kWarp, int KN = kIter>
__global__ void __launch_bounds__(256, TINY ? (KT >= 2 ? 2 : 4) : 1) lookup_bwd_rows_kernel(LookupArgs a) {
  constexpr int GPW = W / LPR;
  constexpr int SPW = kWarp / W;
  // the register-accumulating QR variant keeps registers for the emb1 accumulators: shallower batching there
  constexpr int KI = TINY ? KT : KN;
  FV<V> tacc[TINY ? TR : 1];
#pragma unroll
  for (int r = 0; r < (TINY ? TR : 1); ++r) tacc[r] = FV<V>::zero();
  const int lane = threadIdx.x & 31;
  const int sl = lane % W, sw = lane / W;
  const int g = sl / LPR, c = sl % LPR;
  const bool cact = c * V < a.E;
  const int d0 = c * V;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;

  for (long long b0 = warp * SPW; b0 < a.B; b0 += nwarps * SPW) {
    const bool bvalid = b0 + sw < a.B;
    const long long b = bvalid ? b0 + sw : a.B - 1;     // past-the-end samples: clamp the reads, suppress the writes
    const float gy = a.g_y ? __ldg(a.g_y + b) : 0.f;
    FV<V> S = FV<V>::zero();
    if (a.g_y && cact) S = ldg<V>(a.S + b * a.E + d0);
    for (int vfb = 0; vfb < a.VF; vfb += GPW * KI) {
      // ---- load phase: every load of the batch is issued before the first store ----
      long long rowv[KI];
      FV<V> gdv[KI], eev[KI], t1v[KI], t2v[KI];
#pragma unroll
      for (int it = 0; it < KI; ++it) {
        const int vf = vfb + it * GPW + g;
        const int f = (K == RSB_KIND_QR_CAT && vf >= a.F) ? vf - a.F : vf;
        rowv[it] = (bvalid && vf < a.VF) ? __ldg(a.rows_in + b * a.F + f) : 0;
      }
#pragma unroll
      for (int it = 0; it < KI; ++it) {
        const int vf = vfb + it * GPW + g;
        const bool act = bvalid && (vf < a.VF) && cact;
        const long long o = (b * a.VF + vf) * (long long)a.E + d0;
        gdv[it] = FV<V>::zero();
        eev[it] = FV<V>::zero();
        t1v[it] = FV<V>::zero();
        t2v[it] = FV<V>::zero();
        if (act) {
          if (a.g_deep) gdv[it] = ldg<V>(a.g_deep + o);
          // QR mult: e = e1 * e2 is recomputed from the two (L2-resident) tables with the forward's own
          // rounding instead of streaming the saved [B,F,D] output back from HBM (a third of the kernel's traffic)
          if (a.g_y && K != RSB_KIND_QR_MULT) eev[it] = ldg<V>(a.emb + o);
          if (K == RSB_KIND_QR_MULT) {
            long long i1, i2;
            qr_split(a, rowv[it], i1, i2);
            t1v[it] = ldg<V>(a.table1 + i1 * a.E + d0);
            t2v[it] = ldg<V>(main_row(a, locate(a, i2)) + d0);
#pragma unroll
            for (int i = 0; i < V; ++i) eev[it].v[i] = __fmul_rn(t1v[it].v[i], t2v[it].v[i]);
          } else if (K == RSB_KIND_PEP || K == RSB_KIND_OPTEMBED) {
            t1v[it] = ldg<V>(main_row(a, locate(a, rowv[it])) + d0);
          }
        }
      }
      // ---- compute + store phase ----
#pragma unroll
      for (int it = 0; it < KI; ++it) {
      const int vf = vfb + it * GPW + g;
      const bool vact = bvalid && vf < a.VF;
      const bool act = vact && cact;
      const int f = (K == RSB_KIND_QR_CAT && vf >= a.F) ? vf - a.F : vf;
      const long long row = rowv[it];
      const long long p = (b * a.F + f) * (long long)a.E + d0;    // offset into [n,E] per-lookup arrays
      FV<V> go = gdv[it];
      if (act && a.g_y) {
#pragma unroll
        for (int i = 0; i < V; ++i) go.v[i] = fmaf(gy, S.v[i] - eev[it].v[i], go.v[i]);
      }
      if (K == RSB_KIND_VANILLA) {
        if (act) st<V>(a.rg_main + p, go);
      } else if (K == RSB_KIND_MASK) {
        if (act) {
          const RowLoc l = locate(a, row);
          const unsigned char* m = reinterpret_cast<const unsigned char*>(aux_base(a, l)) + l.lrow * a.E + d0;
          FV<V> r;
#pragma unroll
          for (int i = 0; i < V; ++i) r.v[i] = go.v[i] * (float)(m[i] != 0);
          st<V>(a.rg_main + p, r);
        }
      } else if (K == RSB_KIND_QR_MULT) {
        if (act) {
          long long i1, i2;
          qr_split(a, row, i1, i2);
          const FV<V> e1 = t1v[it], e2 = t2v[it];
          FV<V> r1, r2;
#pragma unroll
          for (int i = 0; i < V; ++i) {
            r2.v[i] = go.v[i] * e1.v[i];
            r1.v[i] = go.v[i] * e2.v[i];
          }
          st<V>(a.rg_main + p, r2);
          if (TINY) {
            const int k1 = (int)i1;
#pragma unroll
            for (int r = 0; r < TR; ++r) {
#pragma unroll
              for (int i = 0; i < V; ++i) tacc[TINY ? r : 0].v[i] += (k1 == r) ? r1.v[i] : 0.f;
            }
          } else {
            st<V>(a.rg_aux + p, r1);
          }
        }
      } else if (K == RSB_KIND_QR_ADD) {
        if (act) {
          st<V>(a.rg_main + p, go);
          if (TINY) {
            long long i1, i2;
            qr_split(a, row, i1, i2);
            const int k1 = (int)i1;
#pragma unroll
            for (int r = 0; r < TR; ++r) {
#pragma unroll
              for (int i = 0; i < V; ++i) tacc[TINY ? r : 0].v[i] += (k1 == r) ? go.v[i] : 0.f;
            }
          } else if (a.rg_aux && a.rg_aux != a.rg_main) {
            st<V>(a.rg_aux + p, go);
          }
        }
      } else if (K == RSB_KIND_QR_CAT) {
        if (act) st<V>((vf < a.F ? a.rg_aux : a.rg_main) + p, go);
      } else if (K == RSB_KIND_PEP) {
        if (act) {
          const FV<V> w = t1v[it];
          const RowLoc l = locate(a, row);
          FV<V> rw, rs;
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float sg = sigmoidf_exact(pep_s(a, l, d0 + i));
            bool keep = (fabsf(w.v[i]) - sg) > 0.0f;
            float sgn = (w.v[i] > 0.f) ? 1.f : ((w.v[i] < 0.f) ? -1.f : 0.f);
            rw.v[i] = keep ? go.v[i] * sgn * sgn : 0.f;
            rs.v[i] = keep ? -(go.v[i] * sgn) * (sg * (1.0f - sg)) : 0.f;
          }
          st<V>(a.rg_main + p, rw);
          if (a.rg_aux) st<V>(a.rg_aux + p, rs);
        }
      } else if (K == RSB_KIND_OPTEMBED) {
        const FV<V> w = t1v[it];
        long long k = (a.mask_d != nullptr && vact) ? __ldg(a.mask_d + b * a.F + f) : (long long)a.E;
        FV<V> u;
#pragma unroll
        for (int i = 0; i < V; ++i) u.v[i] = ((long long)(d0 + i) <= k) ? go.v[i] : 0.f;
        FV<V> r = u;
        if (a.aux != nullptr) {
          float pn = 0.f, pd = 0.f;
#pragma unroll
          for (int i = 0; i < V; ++i) {
            pn += (a.aux_mode == 2) ? w.v[i] * w.v[i] : fabsf(w.v[i]);
            pd = fmaf(u.v[i], w.v[i], pd);
          }
          float nrm = group_sum<LPR>(pn);
          float gme = group_sum<LPR>(pd);
          if (a.aux_mode == 2) nrm = sqrtf(nrm);
          float t = vact ? __ldg(reinterpret_cast<const float*>(a.aux) + f) : 0.f;
          float z = nrm - t;
          float me = (z > 0.f) ? 1.f : 0.f;
          float gz = gme * binary_step_grad(z);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float dn;
            if (a.aux_mode == 2) dn = (nrm > 0.f) ? w.v[i] / nrm : 0.f;
            else dn = (w.v[i] > 0.f) ? 1.f : ((w.v[i] < 0.f) ? -1.f : 0.f);
            r.v[i] = fmaf(gz, dn, u.v[i] * me);
          }
          if (a.rg_aux && vact && c == 0) a.rg_aux[b * a.F + f] = gz;
        }
        if (act) st<V>(a.rg_main + p, r);
      }
      }  // it
    }
  }
  if constexpr (TINY) {
    // lane groups -> warp (shuffles) -> CTA (shared memory, fixed order) -> one partial table per CTA
    __shared__ float tred[8][TR][LPR * V];
    const int wib = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < TR; ++r) {
#pragma unroll
      for (int off = LPR; off < kWarp; off <<= 1) {
        FV<V> o = shfl_xor<V>(tacc[r], off);
#pragma unroll
        for (int i = 0; i < V; ++i) tacc[r].v[i] += o.v[i];
      }
      if (g == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) tred[wib][r][c * V + i] = tacc[r].v[i];
      }
    }
    __syncthreads();
    const int tot = (int)a.divider * a.E;   // rows >= divider are never hit
    float* out = a.tiny_partials + (long long)blockIdx.x * tot;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) {
      const int r = i / a.E, d = i - r * a.E;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += tred[w][r][d];
      out[i] = sum;
    }
  }
}

// 0: one sample per warp, 5 iterations in flight (groups x 5 slots: the 39 Criteo fields in one round at D = 16);
// 1: one sample per warp, 3 iterations (Avazu shape: 22 fields in 24 slots);
// 2: two samples per warp (16 lanes each), 3 iterations (KDD shape: 11 fields in 12 slots).
// Picks the mapping with the fewest lane-slots per sample.
static int pick_mapping(int fields, int groups) {
  auto cost = [&](int g, int ki, int w) {
    if (g < 1) return 1 << 30;
    const int slots = g * ki;
    return ((fields + slots - 1) / slots) * ki * w;
  };
  const int c0 = cost(groups, kIter, 32), c1 = cost(groups, 3, 32), c2 = cost(groups / 2, 3, 16);
  if (c2 < c0 && c2 < c1) return 2;
  return c1 < c0 ? 1 : 0;
}

template <int K>
static int launch_fwd(const LookupArgs& a, RowShape sh, cudaStream_t stream) {
  const int threads = 256;
  long long warps_needed = a.B;
  long long blocks = (warps_needed * 32 + threads - 1) / threads;
  long long cap = (long long)sm_count() * 32;  // grid-stride beyond this
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // lanes per sample / iterations in flight for this field count (see lookup_fwd_kernel)
  const int groups = kWarp / sh.LPR;
  const int variant = pick_mapping(a.VF, groups);
  blocks = (warps_needed / (variant == 2 ? 2 : 1) * 32 + threads - 1) / threads;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
#define CALL(VV, LL)                                                                                                  \
  if (variant == 2 && LL <= 16) lookup_fwd_kernel<K, VV, LL, (LL <= 16 ? 16 : 32), 3><<<(unsigned)blocks, threads, 0, stream>>>(a); \
  else if (variant == 1) lookup_fwd_kernel<K, VV, LL, 32, 3><<<(unsigned)blocks, threads, 0, stream>>>(a);              \
  else lookup_fwd_kernel<K, VV, LL><<<(unsigned)blocks, threads, 0, stream>>>(a)
  RSB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

static long long bwd_blocks(long long B) {
  long long blocks = (B * 32 + 255) / 256;
  long long cap = (long long)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return blocks;
}

// the register-accumulating QR variant writes one partial table per CTA: a few resident CTAs per SM
static int tune(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static long long tiny_blocks(long long B) {
  static const int per_sm = tune("RSB_TINY_CTAS_PER_SM", 4);
  long long blocks = bwd_blocks(B);
  const long long cap = (long long)sm_count() * per_sm;
  return blocks > cap ? cap : blocks;
}

template <int K>
static int launch_bwd(const LookupArgs& a, RowShape sh, cudaStream_t stream) {
  const int threads = 256;
  const long long blocks = bwd_blocks(a.B);
  if constexpr (K == RSB_KIND_QR_MULT || K == RSB_KIND_QR_ADD) {
    if (a.tiny_partials != nullptr) {
      const long long tblocks = tiny_blocks(a.B);
      static const int kt = tune("RSB_TINY_KI", 1);
      const bool five = a.modulus <= 5 && a.divider <= 5;   // rows >= modulus are never hit: fewer accumulators, fewer predicated adds
#define CALLT(VV, LL)                                                                                              \
  if (five && kt == 1) lookup_bwd_rows_kernel<K, VV, LL, true, 1, 5><<<(unsigned)tblocks, threads, 0, stream>>>(a); \
  else if (five) lookup_bwd_rows_kernel<K, VV, LL, true, 2, 5><<<(unsigned)tblocks, threads, 0, stream>>>(a);       \
  else if (kt == 1) lookup_bwd_rows_kernel<K, VV, LL, true, 1><<<(unsigned)tblocks, threads, 0, stream>>>(a);       \
  else lookup_bwd_rows_kernel<K, VV, LL, true, 2><<<(unsigned)tblocks, threads, 0, stream>>>(a)
      RSB_DISPATCH_SHAPE(sh, CALLT);
#undef CALLT
      RSB_CHECK_LAUNCH();
      note_launch(1);
      return RSB_OK;
    }
  }
  const int variant = pick_mapping(a.VF, kWarp / sh.LPR);
  long long vblocks = bwd_blocks(variant == 2 ? (a.B + 1) / 2 : a.B);
#define CALL(VV, LL)                                                                                                       \
  if (variant == 2 && LL <= 16)                                                                                            \
    lookup_bwd_rows_kernel<K, VV, LL, false, 2, kTinyRows, (LL <= 16 ? 16 : 32), 3><<<(unsigned)vblocks, threads, 0, stream>>>(a); \
  else if (variant == 1) lookup_bwd_rows_kernel<K, VV, LL, false, 2, kTinyRows, 32, 3><<<(unsigned)vblocks, threads, 0, stream>>>(a); \
  else lookup_bwd_rows_kernel<K, VV, LL><<<(unsigned)blocks, threads, 0, stream>>>(a)
  RSB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

static int launch_fc_grad(const long long* rows, const float* g_y, long long B, int F, float* fc_grad,
                          cudaStream_t stream) {
  const size_t smem = (size_t)32 * F * sizeof(long long);
  if (smem > 200 * 1024) return RSB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) cudaFuncSetAttribute(fc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  fc_grad_kernel<<<(unsigned)((B + 31) / 32), 256, smem, stream>>>(rows, g_y, B, F, fc_grad);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

static int fill_common(LookupArgs& a, int kind, long long B, int F, int D, const float* table, long long n_rows,
                       long long n_global, const float* table1, long long divider, const void* aux, int aux_mode,
                       RowShape& sh, bool extra_aligned) {
  if (B < 0 || F <= 0 || D <= 0 || table == nullptr) return RSB_ERR_BAD_ARG;
  if (kind < RSB_KIND_VANILLA || kind > RSB_KIND_OPTEMBED) return RSB_ERR_BAD_ARG;
  a.B = B;
  a.F = F;
  a.E = D;
  a.VF = F;
  if (kind == RSB_KIND_QR_CAT) {
    if (D % 2) return RSB_ERR_BAD_ARG;
    a.E = D / 2;
    a.VF = 2 * F;
  }
  if (kind >= RSB_KIND_QR_MULT && kind <= RSB_KIND_QR_CAT) {
    if (table1 == nullptr || divider <= 0) return RSB_ERR_BAD_ARG;
  }
  if ((kind == RSB_KIND_PEP || kind == RSB_KIND_MASK) && aux == nullptr) return RSB_ERR_BAD_ARG;
  if (kind == RSB_KIND_PEP && (aux_mode < RSB_PEP_GLOBAL || aux_mode > RSB_PEP_FEATURE_DIM)) return RSB_ERR_BAD_ARG;
  if (kind == RSB_KIND_OPTEMBED && aux != nullptr && aux_mode != 1 && aux_mode != 2) return RSB_ERR_BAD_ARG;
  a.table = table;
  a.n_rows = n_rows;
  a.n_global = n_global;
  a.table1 = table1;
  a.divider = divider;
  a.modulus = divider;
  if (kind >= RSB_KIND_QR_MULT && kind <= RSB_KIND_QR_CAT && aux_mode > 0) a.modulus = aux_mode;
  a.small32 = (n_global < (1ll << 32) && divider < (1ll << 32)) ? 1 : 0;
  a.fd_div = make_fastdiv((unsigned long long)(divider > 0 ? divider : 1));
  a.fd_mod = make_fastdiv((unsigned long long)(a.modulus > 0 ? a.modulus : 1));
  a.fd_g = make_fastdiv(1);
  a.aux = aux;
  a.aux_mode = aux_mode;
  bool al = aligned16(table) && (table1 == nullptr || aligned16(table1)) && extra_aligned;
  sh = row_shape(a.E, al);
  if (!sh.ok) return RSB_ERR_UNSUPPORTED;
  return RSB_OK;
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int rsb_lookup_fwd(int32_t kind, const void* idx, int32_t idx_is_i32, const int64_t* offsets, int64_t B,
                              int32_t F, int32_t D, const float* table, int64_t n_rows, int64_t n_global,
                              const float* table1, int64_t divider, const void* aux, int32_t aux_mode,
                              const int64_t* mask_d_idx, const float* fc, const float* bias, float* out_emb,
                              float* out_yfm, float* out_sum, int64_t* out_rows, int32_t* err_flag, float* amax_slots,
                              void* stream) {
  LookupArgs a = {};
  RowShape sh;
  if (B == 0) return RSB_OK;  // empty batch: nothing to do (pointers of empty tensors may be NULL)
  if (idx == nullptr || out_emb == nullptr) return RSB_ERR_BAD_ARG;
  bool al = aligned16(out_emb) && (out_sum == nullptr || aligned16(out_sum));
  int rc = fill_common(a, kind, B, F, D, table, n_rows, n_global, table1, divider, aux, aux_mode, sh, al);
  if (rc) return rc;
  if (B == 0) return RSB_OK;
  a.idx = idx;
  a.idx_i32 = idx_is_i32;
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.mask_d = reinterpret_cast<const long long*>(mask_d_idx);
  a.fc = fc;
  a.bias = bias;
  a.out_emb = out_emb;
  a.out_y = out_yfm;
  a.out_sum = out_sum;
  a.out_rows = reinterpret_cast<long long*>(out_rows);
  a.err = err_flag;
  a.amax_slots = amax_slots;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (kind) {
    case RSB_KIND_VANILLA: return launch_fwd<RSB_KIND_VANILLA>(a, sh, s);
    case RSB_KIND_QR_MULT: return launch_fwd<RSB_KIND_QR_MULT>(a, sh, s);
    case RSB_KIND_QR_ADD: return launch_fwd<RSB_KIND_QR_ADD>(a, sh, s);
    case RSB_KIND_QR_CAT: return launch_fwd<RSB_KIND_QR_CAT>(a, sh, s);
    case RSB_KIND_PEP: return launch_fwd<RSB_KIND_PEP>(a, sh, s);
    case RSB_KIND_MASK: return launch_fwd<RSB_KIND_MASK>(a, sh, s);
    default: return launch_fwd<RSB_KIND_OPTEMBED>(a, sh, s);
  }
}

extern "C" RSB_API int64_t rsb_qr_bwd_fused_workspace_bytes(int64_t B, int32_t D);

struct ShardInfo {
  const float* const* table_shards;
  const void* const* aux_shards;
  int G;
};

static int bwd_rows_impl(int32_t kind, const int64_t* rows, int64_t B, int32_t F, int32_t D, const float* table,
                         int64_t n_rows, const float* table1, int64_t divider, const void* aux, int32_t aux_mode,
                         const int64_t* mask_d_idx, const float* emb, const float* S, const float* g_yfm,
                         const float* g_deep, float* rg_main, float* rg_aux, float* fc_grad, float* table1_grad,
                         void* workspace, int64_t workspace_bytes, void* stream, const ShardInfo* shards = nullptr) {
  LookupArgs a = {};
  RowShape sh;
  if (B == 0) return RSB_OK;
  if (rows == nullptr || rg_main == nullptr) return RSB_ERR_BAD_ARG;
  if (g_yfm == nullptr && g_deep == nullptr) return RSB_ERR_BAD_ARG;
  if (g_yfm != nullptr && (emb == nullptr || S == nullptr)) return RSB_ERR_BAD_ARG;
  if ((kind == RSB_KIND_QR_MULT || kind == RSB_KIND_QR_CAT) && rg_aux == nullptr && table1_grad == nullptr)
    return RSB_ERR_BAD_ARG;
  bool al = aligned16(rg_main) && (rg_aux == nullptr || kind == RSB_KIND_OPTEMBED || aligned16(rg_aux)) &&
            (emb == nullptr || aligned16(emb)) && (S == nullptr || aligned16(S)) &&
            (g_deep == nullptr || aligned16(g_deep));
  int rc = fill_common(a, kind, B, F, D, table, n_rows, /*n_global=*/(1ll << 62), table1, divider, aux, aux_mode, sh,
                       al);
  if (rc) return rc;
  if (B == 0) return RSB_OK;
  // small32 must match the forward's index math: rows were validated there
  a.small32 = (divider < (1ll << 32)) ? 1 : 0;
  if (kind >= RSB_KIND_QR_MULT && kind <= RSB_KIND_QR_CAT) {
    // emb2 has n_rows rows, so every id is < n_rows * divider
    long double lim = (long double)n_rows * (long double)divider;
    a.small32 = (lim < 4294967296.0L && divider < (1ll << 32)) ? 1 : 0;
  }
  if (shards != nullptr) {
    // rows of the main table (and of a per-row aux array) are read from the owners' shards
    a.table = nullptr;
    a.table_shards = shards->table_shards;
    a.aux_shards = shards->aux_shards;
    a.G = shards->G;
    a.fd_g = make_fastdiv((unsigned long long)shards->G);
    if (n_rows >= (1ll << 32)) a.small32 = 0;
  }
  a.rows_in = reinterpret_cast<const long long*>(rows);
  a.mask_d = reinterpret_cast<const long long*>(mask_d_idx);
  a.emb = emb;
  a.S = S;
  a.g_y = g_yfm;
  a.g_deep = g_deep;
  a.rg_main = rg_main;
  a.rg_aux = rg_aux;
  a.fc_grad = fc_grad;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (table1_grad != nullptr) {
    // fused small-table gradient: per-CTA register accumulators -> partials in the workspace
    if (!(kind == RSB_KIND_QR_MULT || kind == RSB_KIND_QR_ADD) || divider > kTinyRows) return RSB_ERR_UNSUPPORTED;
    if (workspace == nullptr || workspace_bytes < rsb_qr_bwd_fused_workspace_bytes(B, D)) return RSB_ERR_WORKSPACE;
    a.tiny_partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  }
  switch (kind) {
    case RSB_KIND_VANILLA: rc = launch_bwd<RSB_KIND_VANILLA>(a, sh, s); break;
    case RSB_KIND_QR_MULT: rc = launch_bwd<RSB_KIND_QR_MULT>(a, sh, s); break;
    case RSB_KIND_QR_ADD: rc = launch_bwd<RSB_KIND_QR_ADD>(a, sh, s); break;
    case RSB_KIND_QR_CAT: rc = launch_bwd<RSB_KIND_QR_CAT>(a, sh, s); break;
    case RSB_KIND_PEP: rc = launch_bwd<RSB_KIND_PEP>(a, sh, s); break;
    case RSB_KIND_MASK: rc = launch_bwd<RSB_KIND_MASK>(a, sh, s); break;
    default: rc = launch_bwd<RSB_KIND_OPTEMBED>(a, sh, s); break;
  }
  if (rc) return rc;
  if (table1_grad != nullptr) {
    const int nblk = (int)tiny_blocks(B);
    const int dst_elems = (int)divider * a.E;
    partials_reduce_kernel<<<(dst_elems + 31) / 32, dim3(32, 32), 0, s>>>(a.tiny_partials, nblk, dst_elems,
                                                                          table1_grad);
    RSB_CHECK_LAUNCH();
    note_launch(1);
  }
  if (fc_grad != nullptr && g_yfm != nullptr) {
    rc = launch_fc_grad(reinterpret_cast<const long long*>(rows), g_yfm, B, F, fc_grad, s);
    if (rc) return rc;
  }
  return RSB_OK;
}

extern "C" RSB_API int64_t rsb_qr_bwd_fused_workspace_bytes(int64_t B, int32_t D) {
  if (B < 0 || D <= 0) return 0;
  return (int64_t)bwd_blocks(B) * kTinyRows * D * 4 + 256;
}

extern "C" RSB_API int rsb_lookup_bwd_rows(int32_t kind, const int64_t* rows, int64_t B, int32_t F, int32_t D,
                                   const float* table, int64_t n_rows, const float* table1, int64_t divider,
                                   const void* aux, int32_t aux_mode, const int64_t* mask_d_idx, const float* emb,
                                   const float* S, const float* g_yfm, const float* g_deep, float* rg_main,
                                   float* rg_aux, float* fc_grad, void* stream) {
  return bwd_rows_impl(kind, rows, B, F, D, table, n_rows, table1, divider, aux, aux_mode, mask_d_idx, emb, S, g_yfm,
                       g_deep, rg_main, rg_aux, fc_grad, nullptr, nullptr, 0, stream);
}

extern "C" RSB_API int rsb_qr_bwd_fused(int32_t kind, const int64_t* rows, int64_t B, int32_t F, int32_t D,
                                        const float* table, int64_t n_rows, const float* table1, int64_t divider,
                                        const float* emb, const float* S, const float* g_yfm, const float* g_deep,
                                        float* rg_main, float* table1_grad, float* fc_grad, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
  if (table1_grad == nullptr) return RSB_ERR_BAD_ARG;
  return bwd_rows_impl(kind, rows, B, F, D, table, n_rows, table1, divider, nullptr, 0, nullptr, emb, S, g_yfm, g_deep,
                       rg_main, nullptr, fc_grad, table1_grad, workspace, workspace_bytes, stream);
}

extern "C" RSB_API int rsb_lookup_fwd_sharded(const void* idx, int32_t idx_is_i32, const int64_t* offsets, int64_t B,
                                              int32_t F, int32_t D, const float* const* table_shards,
                                              const float* fc_replicated, int32_t G, int64_t n_global,
                                              const float* bias, const float* hot_table, const int64_t* hot_map,
                                              float* out_emb, float* out_yfm, float* out_sum,
                                              int64_t* out_rows, int32_t* err_flag, float* amax_slots, void* stream) {
  LookupArgs a = {};
  RowShape sh;
  if (B == 0) return RSB_OK;
  if (idx == nullptr || out_emb == nullptr || table_shards == nullptr || G < 1) return RSB_ERR_BAD_ARG;
  if ((hot_map != nullptr) != (hot_table != nullptr)) return RSB_ERR_BAD_ARG;
  if (hot_table != nullptr && !aligned16(hot_table)) return RSB_ERR_UNSUPPORTED;
  bool al = aligned16(out_emb) && (out_sum == nullptr || aligned16(out_sum));
  // shard base pointers come from cudaMalloc (256 B aligned); use a non-null dummy for the common checks
  int rc = fill_common(a, RSB_KIND_VANILLA, B, F, D, reinterpret_cast<const float*>(out_emb), n_global, n_global,
                       nullptr, 0, nullptr, 0, sh, al);
  if (rc) return rc;
  a.table = nullptr;
  a.table_shards = table_shards;
  a.hot_table = hot_table;
  a.hot_map = reinterpret_cast<const long long*>(hot_map);
  a.fc = fc_replicated;
  a.G = G;
  a.fd_g = make_fastdiv((unsigned long long)G);
  a.small32 = (n_global < (1ll << 32)) ? 1 : 0;
  a.idx = idx;
  a.idx_i32 = idx_is_i32;
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.bias = bias;
  a.out_emb = out_emb;
  a.out_y = out_yfm;
  a.out_sum = out_sum;
  a.out_rows = reinterpret_cast<long long*>(out_rows);
  a.err = err_flag;
  a.amax_slots = amax_slots;
  return launch_fwd<RSB_KIND_VANILLA>(a, sh, reinterpret_cast<cudaStream_t>(stream));
}

// Any variant over row-sharded tables: the main table (vanilla / PEP / mask / OptEmbed weight, QR emb2) and the per-row
// aux array (PEP s of the feature / feature_dim kinds, retrain masks) live in G shards; everything small stays local.
static bool aux_is_per_row(int kind, int aux_mode) {
  return kind == RSB_KIND_MASK || (kind == RSB_KIND_PEP && (aux_mode == RSB_PEP_FEATURE || aux_mode == RSB_PEP_FEATURE_DIM));
}

extern "C" RSB_API int rsb_lookup_fwd_sharded_kind(int32_t kind, const void* idx, int32_t idx_is_i32, const int64_t* offsets,
                                                   int64_t B, int32_t F, int32_t D, const float* const* table_shards,
                                                   int32_t G, int64_t n_rows, int64_t n_global, const float* table1,
                                                   int64_t divider, const void* aux, const void* const* aux_shards,
                                                   int32_t aux_mode, const int64_t* mask_d_idx, const float* fc_replicated,
                                                   const float* bias, float* out_emb, float* out_yfm, float* out_sum,
                                                   int64_t* out_rows, int32_t* err_flag, float* amax_slots, void* stream) {
  LookupArgs a = {};
  RowShape sh;
  if (B == 0) return RSB_OK;
  if (idx == nullptr || out_emb == nullptr || table_shards == nullptr || G < 1 || n_rows <= 0) return RSB_ERR_BAD_ARG;
  const bool per_row = aux_is_per_row(kind, aux_mode);
  if (per_row ? (aux_shards == nullptr) : (aux_shards != nullptr)) return RSB_ERR_BAD_ARG;
  bool al = aligned16(out_emb) && (out_sum == nullptr || aligned16(out_sum));
  // shard base pointers come from cudaMalloc (256 B aligned); a non-null dummy stands in for the common checks
  const void* aux_chk = per_row ? reinterpret_cast<const void*>(out_emb) : aux;
  int rc = fill_common(a, kind, B, F, D, reinterpret_cast<const float*>(out_emb), n_rows, n_global, table1, divider, aux_chk,
                       aux_mode, sh, al);
  if (rc) return rc;
  a.table = nullptr;
  a.aux = per_row ? nullptr : aux;
  a.table_shards = table_shards;
  a.aux_shards = aux_shards;
  a.G = G;
  a.fd_g = make_fastdiv((unsigned long long)G);
  if (n_rows >= (1ll << 32)) a.small32 = 0;
  a.idx = idx;
  a.idx_i32 = idx_is_i32;
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.mask_d = reinterpret_cast<const long long*>(mask_d_idx);
  a.fc = fc_replicated;
  a.bias = bias;
  a.out_emb = out_emb;
  a.out_y = out_yfm;
  a.out_sum = out_sum;
  a.out_rows = reinterpret_cast<long long*>(out_rows);
  a.err = err_flag;
  a.amax_slots = amax_slots;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (kind) {
    case RSB_KIND_VANILLA: return launch_fwd<RSB_KIND_VANILLA>(a, sh, s);
    case RSB_KIND_QR_MULT: return launch_fwd<RSB_KIND_QR_MULT>(a, sh, s);
    case RSB_KIND_QR_ADD: return launch_fwd<RSB_KIND_QR_ADD>(a, sh, s);
    case RSB_KIND_QR_CAT: return launch_fwd<RSB_KIND_QR_CAT>(a, sh, s);
    case RSB_KIND_PEP: return launch_fwd<RSB_KIND_PEP>(a, sh, s);
    case RSB_KIND_MASK: return launch_fwd<RSB_KIND_MASK>(a, sh, s);
    default: return launch_fwd<RSB_KIND_OPTEMBED>(a, sh, s);
  }
}

extern "C" RSB_API int rsb_lookup_bwd_rows_sharded(int32_t kind, const int64_t* rows, int64_t B, int32_t F, int32_t D,
                                                   const float* const* table_shards, int32_t G, int64_t n_rows,
                                                   const float* table1, int64_t divider, const void* aux,
                                                   const void* const* aux_shards, int32_t aux_mode,
                                                   const int64_t* mask_d_idx, const float* emb, const float* S,
                                                   const float* g_yfm, const float* g_deep, float* rg_main, float* rg_aux,
                                                   void* stream) {
  if (B == 0) return RSB_OK;
  if (table_shards == nullptr || G < 1 || rg_main == nullptr) return RSB_ERR_BAD_ARG;
  const bool per_row = aux_is_per_row(kind, aux_mode);
  if (per_row ? (aux_shards == nullptr) : (aux_shards != nullptr)) return RSB_ERR_BAD_ARG;
  ShardInfo si = {table_shards, aux_shards, G};
  // non-null, 16-byte aligned stand-ins for the pointers the common checks look at
  return bwd_rows_impl(kind, rows, B, F, D, rg_main, n_rows, table1, divider, per_row ? reinterpret_cast<const void*>(rg_main) : aux,
                       aux_mode, mask_d_idx, emb, S, g_yfm, g_deep, rg_main, rg_aux, nullptr, nullptr, nullptr, 0, stream, &si);
}

extern "C" RSB_API int rsb_fc_grad(const int64_t* rows, const float* g_yfm, int64_t B, int32_t F, float* fc_grad,
                                   void* stream) {
  if (B < 0 || F <= 0) return RSB_ERR_BAD_ARG;
  if (B == 0) return RSB_OK;
  if (!rows || !g_yfm || !fc_grad) return RSB_ERR_BAD_ARG;
  return launch_fc_grad(reinterpret_cast<const long long*>(rows), g_yfm, B, F, fc_grad,
                        reinterpret_cast<cudaStream_t>(stream));
}
