// Backward stage 3: deterministic segmented reduction of per-lookup row gradients over
// the row-sorted lookups, with the consumer fused in (dense grad row write, SparseAdam
// row update, sparse SGD row update).
//
// Work is split by POSITION, not by segment, so hot rows (a 4-value field looked up by
// every sample) cannot unbalance it: every lane group (LPR lanes = one row wide) walks a
// chunk of kChunk consecutive sorted positions and sums runs of equal row ids in
// registers.  A run that lies inside one chunk is final and is applied at once.  A run
// that crosses a chunk boundary leaves one partial per chunk (part_first / part_last);
// a second launch with one lane group per chunk boundary adds the partials of a run in
// chunk order and applies the total.  No atomics, fixed summation order -> bit-reproducible.
//
// Reference semantics: aten embedding_dense_backward (dense grads of nn.Embedding,
// src/models/embeddings/base.py:53-57), torch.optim.SparseAdam (coalesce + row update,
// torch/optim/_functional.py:24-84, used at src/models/deepfm.py:173-184) and sparse SGD
// (src/models/deepfm.py:203-216).
#include "common.cuh"

namespace rsb {

constexpr int kChunk = 32;          // sorted positions per lane group
constexpr int kSegThreads = 128;    // 4 warps per CTA
constexpr unsigned kNoKey = 0xffffffffu;

struct ApplyArgs {
  int mode;
  float* dst;
  float* m;
  float* v;
  float one_minus_b1, one_minus_b2, eps, neg_step_size, neg_lr;
  int E;
  // RSB_APPLY_SHARD_ATOMIC: dst row lives on rank row % G at local row row / G (peer memory)
  float* const* shards;
  int G;
  float scale;
  // source row of sorted position i: perm[i], or perm[i] / src_div when one source row serves src_div consecutive
  // lookups (the first-order gradient: every lookup of sample b carries g_y[b], rsb_fc_grad_sorted)
  int use_src_div;
  FastDiv src_div;
  // RSB_APPLY_SHARD_ATOMIC with replicated small fields: hot_map[f] = (lo, hi, delta), lo ascending (the field
  // offsets); a row inside [lo, hi) of its field is stored (unscaled, one writer) at hot_grad[row + delta]
  const long long* hot_map;
  int n_fields;
  float* hot_grad;
};

// shared-memory slot of sorted position i: one pad word per 32 so that the lane groups of a warp, whose chunks
// start kChunk = 32 positions apart, read different banks
__device__ __forceinline__ int pad32(int i) { return i + (i >> 5); }

__device__ __forceinline__ void atomic_add_row(float* p, const FV<4>& g) {
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(g.v[0], g.v[1], g.v[2], g.v[3]));  // red.global.add.v4.f32
}
__device__ __forceinline__ void atomic_add_row(float* p, const FV<1>& g) { atomicAdd(p, g.v[0]); }

template <int V>
__device__ __forceinline__ void apply_row(const ApplyArgs& ap, unsigned row, int d0, const FV<V>& g) {
  if (ap.mode == RSB_APPLY_SHARD_ATOMIC) {
    if (ap.hot_map != nullptr) {
      // last field whose offset is <= row (39 fields: 6 steps over an L1-resident array, once per unique row)
      int lo = 0, hi = ap.n_fields;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((long long)row >= __ldg(ap.hot_map + 3 * mid)) lo = mid;
        else hi = mid;
      }
      if ((long long)row < __ldg(ap.hot_map + 3 * lo + 1)) {
        st<V>(ap.hot_grad + ((long long)row + __ldg(ap.hot_map + 3 * lo + 2)) * ap.E + d0, g);
        return;
      }
    }
    const unsigned q = row / (unsigned)ap.G;
    const unsigned owner = row - q * (unsigned)ap.G;
    FV<V> s;
#pragma unroll
    for (int i = 0; i < V; ++i) s.v[i] = g.v[i] * ap.scale;
    atomic_add_row(ap.shards[owner] + (long long)q * ap.E + d0, s);
    return;
  }
  const long long o = (long long)row * ap.E + d0;
  if (ap.mode == RSB_APPLY_DENSE) {
    st<V>(ap.dst + o, g);
  } else if (ap.mode == RSB_APPLY_SPARSE_SGD) {
    FV<V> w = ld<V>(ap.dst + o);
#pragma unroll
    for (int i = 0; i < V; ++i) w.v[i] = w.v[i] + ap.neg_lr * g.v[i];
    st<V>(ap.dst + o, w);
  } else {
    // torch/optim/_functional.py:60-84, same operation order
    FV<V> w = ld<V>(ap.dst + o), m = ld<V>(ap.m + o), v = ld<V>(ap.v + o);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float mu = __fmul_rn(__fsub_rn(g.v[i], m.v[i]), ap.one_minus_b1);
      float vu = __fmul_rn(__fsub_rn(__fmul_rn(g.v[i], g.v[i]), v.v[i]), ap.one_minus_b2);
      float mn = __fadd_rn(m.v[i], mu);
      float vn = __fadd_rn(v.v[i], vu);
      float numer = __fadd_rn(mu, m.v[i]);
      float denom = __fadd_rn(sqrtf(__fadd_rn(vu, v.v[i])), ap.eps);
      m.v[i] = mn;
      v.v[i] = vn;
      w.v[i] = __fadd_rn(w.v[i], __fmul_rn(ap.neg_step_size, __fdiv_rn(numer, denom)));
    }
    st<V>(ap.m + o, m);
    st<V>(ap.v + o, v);
    st<V>(ap.dst + o, w);
  }
}

template <int V, int LPR>
__global__ void __launch_bounds__(kSegThreads) seg_chunk_kernel(const unsigned* __restrict__ skeys,
                                                                const unsigned* __restrict__ perm, long long n,
                                                                const float* __restrict__ rg, ApplyArgs ap,
                                                                float* part_first, float* part_last) {
  constexpr int GPW = kWarp / LPR;
  constexpr int WT = GPW * kChunk;  // positions per warp
  constexpr int NW = kSegThreads / 32;
  __shared__ unsigned s_key[NW][WT + 2 + (WT + 2) / 32 + 1];
  __shared__ unsigned s_perm[NW][WT + WT / 32 + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane / LPR, c = lane % LPR;
  const int d0 = c * V;
  const bool cact = d0 < ap.E;
  const long long warp = (long long)blockIdx.x * NW + wib;
  const long long base = warp * WT;
  if (base >= n) return;

  for (int i = lane; i < WT + 2; i += 32) {
    long long p = base - 1 + i;
    s_key[wib][pad32(i)] = (p >= 0 && p < n) ? __ldg(skeys + p) : kNoKey;
  }
  for (int i = lane; i < WT; i += 32) {
    long long p = base + i;
    unsigned src = (p < n) ? __ldg(perm + p) : 0u;
    if (ap.use_src_div) src = fastdiv(src, ap.src_div);
    s_perm[wib][pad32(i)] = src;
  }
  __syncwarp();

  const int start = g * kChunk;
  const long long chunk_id = warp * GPW + g;
  if (base + start >= n) return;
  const unsigned prev_key = s_key[wib][pad32(start)];  // key just before this chunk (slot 0 == base-1)
  unsigned cur = kNoKey;
  bool began0 = false;
  FV<V> acc = FV<V>::zero();

  constexpr int U = 4;
#pragma unroll 1
  for (int i0 = 0; i0 < kChunk; i0 += U) {
    FV<V> val[U];
    unsigned key[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u;
      key[u] = s_key[wib][pad32(1 + start + i)];
      val[u] = FV<V>::zero();
      if (key[u] != kNoKey && cact) val[u] = ldg<V>(rg + (long long)s_perm[wib][pad32(start + i)] * ap.E + d0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u;
      if (key[u] == kNoKey) continue;  // past n (keys are < 2^32-1 only if n_rows < 2^32; see host check)
      if (key[u] != cur) {
        if (cur != kNoKey) {
          // a run ended strictly inside the chunk: it cannot cross the chunk end
          if (began0 && prev_key == cur) {
            if (cact) st<V>(part_first + chunk_id * ap.E + d0, acc);
          } else if (cact) {
            apply_row<V>(ap, cur, d0, acc);
          }
        }
        cur = key[u];
        began0 = (i == 0);
        acc = FV<V>::zero();
      }
#pragma unroll
      for (int k = 0; k < V; ++k) acc.v[k] += val[u].v[k];
    }
  }
  if (cur != kNoKey) {
    const unsigned next_key = s_key[wib][pad32(1 + start + kChunk)];
    const bool crosses_start = began0 && prev_key == cur;
    const bool crosses_end = next_key == cur;
    if (cact) {
      if (crosses_start) st<V>(part_first + chunk_id * ap.E + d0, acc);
      else if (crosses_end) st<V>(part_last + chunk_id * ap.E + d0, acc);
      else apply_row<V>(ap, cur, d0, acc);
    }
  }
}

// One lane group per chunk boundary cb >= 1 (position P = cb * kChunk).  The group that sees
// the FIRST crossing of a run owns it: it finds the run's end by binary search (keys are
// sorted), then either adds the <= kSerialMax partials itself or queues the run for
// seg_long_kernel (a hot row looked up by every sample spans thousands of chunks).
constexpr int kSerialMax = 8;
constexpr int kLongThreads = 256;

template <int V, int LPR>
__global__ void __launch_bounds__(kSegThreads) seg_boundary_kernel(const unsigned* __restrict__ skeys, long long n,
                                                                   long long n_chunks, ApplyArgs ap,
                                                                   const float* __restrict__ part_first,
                                                                   const float* __restrict__ part_last,
                                                                   int* long_count, long long* long_list) {
  constexpr int GPW = kWarp / LPR;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, c = lane % LPR;
  const int d0 = c * V;
  const bool cact = d0 < ap.E;
  const long long grp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32 * GPW + g;
  const long long cb = grp + 1;
  if (cb >= n_chunks) return;
  const long long P = cb * kChunk;
  if (P >= n) return;
  const unsigned row = __ldg(skeys + P);
  if (__ldg(skeys + P - 1) != row) return;                                    // no run crosses here
  if (cb >= 2 && __ldg(skeys + (cb - 1) * kChunk - 1) == row) return;         // not the first crossing
  // last position of the run: invariant key[lo] == row, key[hi] > row (or hi == n)
  long long lo = P, hi = P + (long long)kChunk * kSerialMax;
  if (hi >= n || __ldg(skeys + hi) != row) {
    if (hi > n) hi = n;
  } else {
    lo = hi;
    hi = n;
  }
  while (hi - lo > 1) {
    const long long mid = lo + ((hi - lo) >> 1);
    if (__ldg(skeys + mid) == row) lo = mid;
    else hi = mid;
  }
  const long long c_last = lo / kChunk;
  if (c_last - cb + 1 > kSerialMax) {
    if (c == 0) {
      const int slot = atomicAdd(long_count, 1);
      long_list[2 * slot] = cb;
      long_list[2 * slot + 1] = c_last;
    }
    return;
  }
  if (!cact) return;
  FV<V> tot = ldg<V>(part_last + (cb - 1) * ap.E + d0);
  for (long long cc = cb; cc <= c_last; ++cc) {
    FV<V> p = ldg<V>(part_first + cc * ap.E + d0);
#pragma unroll
    for (int k = 0; k < V; ++k) tot.v[k] += p.v[k];
  }
  apply_row<V>(ap, row, d0, tot);
}

// Long runs: one CTA per run, lane groups stride over the run's chunk partials, then a
// fixed-shape tree in shared memory (deterministic).
template <int V, int LPR>
__global__ void __launch_bounds__(kLongThreads) seg_long_kernel(const unsigned* __restrict__ skeys, ApplyArgs ap,
                                                                const float* __restrict__ part_first,
                                                                const float* __restrict__ part_last,
                                                                const int* __restrict__ long_count,
                                                                const long long* __restrict__ long_list) {
  constexpr int G = kLongThreads / LPR;
  __shared__ float red[G][LPR * V];
  const int g = threadIdx.x / LPR, c = threadIdx.x % LPR;
  const int d0 = c * V;
  const bool cact = d0 < ap.E;
  const int cnt = *long_count;
  for (int w = blockIdx.x; w < cnt; w += gridDim.x) {
    const long long cb = long_list[2 * w], c_last = long_list[2 * w + 1];
    const unsigned row = __ldg(skeys + cb * kChunk);
    FV<V> acc = FV<V>::zero();
    if (cact) {
      if (g == 0) acc = ldg<V>(part_last + (cb - 1) * ap.E + d0);
      for (long long cc = cb + g; cc <= c_last; cc += G) {
        FV<V> p = ldg<V>(part_first + cc * ap.E + d0);
#pragma unroll
        for (int k = 0; k < V; ++k) acc.v[k] += p.v[k];
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) red[g][c * V + k] = acc.v[k];
    __syncthreads();
    for (int s = G / 2; s > 0; s >>= 1) {
      if (g < s) {
#pragma unroll
        for (int k = 0; k < V; ++k) red[g][c * V + k] += red[g + s][c * V + k];
      }
      __syncthreads();
    }
    if (g == 0 && cact) {
      FV<V> tot;
#pragma unroll
      for (int k = 0; k < V; ++k) tot.v[k] = red[0][c * V + k];
      apply_row<V>(ap, row, d0, tot);
    }
    __syncthreads();
  }
}

}  // namespace rsb

using namespace rsb;

static long long seg_align(long long x) { return (x + 255) / 256 * 256; }

extern "C" RSB_API int64_t rsb_segment_workspace_bytes(int64_t n, int32_t E) {
  if (n < 0 || E <= 0) return 0;
  long long n_chunks = (n + kChunk - 1) / kChunk + 1;
  long long max_long = n_chunks / kSerialMax + 2;
  return 2 * seg_align(n_chunks * E * 4) + seg_align(256) + seg_align(max_long * 16) + 256;
}

static int segment_launch(ApplyArgs ap, RowShape sh, const uint32_t* sorted_keys, const uint32_t* perm, int64_t n,
                          const float* row_grads, int32_t E, void* workspace, cudaStream_t s);

extern "C" RSB_API int rsb_segment_scatter_shards(const uint32_t* sorted_keys, const uint32_t* perm, int64_t n,
                                                  const float* row_grads, int32_t E, float* const* grad_shards,
                                                  int32_t G, float scale, const int64_t* hot_map, int32_t n_fields,
                                                  float* hot_grad, void* workspace, int64_t workspace_bytes,
                                                  void* stream) {
  if (n < 0 || E <= 0 || G < 1) return RSB_ERR_BAD_ARG;
  if ((hot_map != nullptr) != (hot_grad != nullptr) || (hot_map != nullptr && n_fields < 1)) return RSB_ERR_BAD_ARG;
  if (n == 0) return RSB_OK;
  if (!sorted_keys || !perm || !row_grads || !grad_shards || !workspace) return RSB_ERR_BAD_ARG;
  if (workspace_bytes < rsb_segment_workspace_bytes(n, E)) return RSB_ERR_WORKSPACE;
  RowShape sh = row_shape(E, aligned16(row_grads));
  if (!sh.ok) return RSB_ERR_UNSUPPORTED;
  ApplyArgs ap = {};
  ap.mode = RSB_APPLY_SHARD_ATOMIC;
  ap.E = E;
  ap.shards = grad_shards;
  ap.G = G;
  ap.scale = scale;
  ap.hot_map = reinterpret_cast<const long long*>(hot_map);
  ap.n_fields = n_fields;
  ap.hot_grad = hot_grad;
  if (hot_grad != nullptr && !aligned16(hot_grad)) return RSB_ERR_UNSUPPORTED;
  return segment_launch(ap, sh, sorted_keys, perm, n, row_grads, E, workspace, reinterpret_cast<cudaStream_t>(stream));
}

// First-order weight gradient fc_grad[row] = sum of g_y[b] over the lookups (b, f) of that row, in sorted order:
// the same three kernels with a 1-wide "row" whose source is g_y[perm / F].  Every touched row has exactly one
// writer and a fixed summation order (the atomic rsb_fc_grad is neither).  fc_grad must be zero-filled.
extern "C" RSB_API int rsb_fc_grad_sorted(const uint32_t* sorted_keys, const uint32_t* perm, int64_t n, const float* g_y,
                                          int32_t F, float* fc_grad, void* workspace, int64_t workspace_bytes,
                                          void* stream) {
  if (n < 0 || F <= 0) return RSB_ERR_BAD_ARG;
  if (n == 0) return RSB_OK;
  if (!sorted_keys || !perm || !g_y || !fc_grad || !workspace) return RSB_ERR_BAD_ARG;
  if (workspace_bytes < rsb_segment_workspace_bytes(n, 1)) return RSB_ERR_WORKSPACE;
  RowShape sh;
  sh.V = 1; sh.LPR = 1; sh.ok = true;
  ApplyArgs ap = {};
  ap.mode = RSB_APPLY_DENSE;
  ap.dst = fc_grad;
  ap.E = 1;
  ap.use_src_div = 1;
  ap.src_div = make_fastdiv((unsigned long long)F);
  return segment_launch(ap, sh, sorted_keys, perm, n, g_y, 1, workspace, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" RSB_API int rsb_segment_reduce_apply(int32_t apply, const uint32_t* sorted_keys, const uint32_t* perm, int64_t n,
                                        const float* row_grads, int32_t E, float* dst, float* exp_avg,
                                        float* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                                        int64_t step, void* workspace, int64_t workspace_bytes, void* stream) {
  if (n < 0 || E <= 0) return RSB_ERR_BAD_ARG;
  if (apply < RSB_APPLY_DENSE || apply > RSB_APPLY_SPARSE_SGD) return RSB_ERR_BAD_ARG;
  if (n == 0) return RSB_OK;
  if (!sorted_keys || !perm || !row_grads || !dst || !workspace) return RSB_ERR_BAD_ARG;
  if (apply == RSB_APPLY_SPARSE_ADAM && (!exp_avg || !exp_avg_sq || step < 1)) return RSB_ERR_BAD_ARG;
  if (workspace_bytes < rsb_segment_workspace_bytes(n, E)) return RSB_ERR_WORKSPACE;
  bool al = aligned16(row_grads) && aligned16(dst) && (!exp_avg || aligned16(exp_avg)) &&
            (!exp_avg_sq || aligned16(exp_avg_sq));
  RowShape sh = row_shape(E, al);
  if (!sh.ok) return RSB_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);

  ApplyArgs ap = {};
  ap.mode = apply;
  ap.dst = dst;
  ap.m = exp_avg;
  ap.v = exp_avg_sq;
  ap.E = E;
  // Scalars exactly as torch computes them (python doubles, then one rounding to fp32)
  ap.one_minus_b1 = (float)(1.0 - (double)beta1);
  ap.one_minus_b2 = (float)(1.0 - (double)beta2);
  ap.eps = eps;
  ap.neg_lr = -lr;
  ap.neg_step_size = 0.f;
  if (apply == RSB_APPLY_SPARSE_ADAM) {
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    ap.neg_step_size = (float)(-((double)lr * sqrt(bc2) / bc1));
  }
  return segment_launch(ap, sh, sorted_keys, perm, n, row_grads, E, workspace, s);
}

static int segment_launch(ApplyArgs ap, RowShape sh, const uint32_t* sorted_keys, const uint32_t* perm, int64_t n,
                          const float* row_grads, int32_t E, void* workspace, cudaStream_t s) {

  const long long n_chunks = (n + kChunk - 1) / kChunk;
  char* w = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  float* part_first = reinterpret_cast<float*>(w);
  float* part_last = reinterpret_cast<float*>(w + seg_align((n_chunks + 1) * E * 4));
  int* long_count = reinterpret_cast<int*>(w + 2 * seg_align((n_chunks + 1) * E * 4));
  long long* long_list = reinterpret_cast<long long*>(w + 2 * seg_align((n_chunks + 1) * E * 4) + seg_align(256));

  const int gpw = 32 / sh.LPR;
  const long long warp_tile = (long long)gpw * kChunk;
  const long long warps = (n + warp_tile - 1) / warp_tile;
  const int wpb = kSegThreads / 32;
  const long long blocks1 = (warps + wpb - 1) / wpb;
  const long long groups2 = n_chunks - 1;
  const long long blocks2 = (groups2 + (long long)gpw * wpb - 1) / ((long long)gpw * wpb);
  long long lb = (n_chunks / kSerialMax) + 1;
  if (lb > 2ll * sm_count()) lb = 2ll * sm_count();
  const unsigned long_blocks = (unsigned)lb;
  if (groups2 > 0) {
    cudaError_t me = cudaMemsetAsync(long_count, 0, sizeof(int), s);
    if (me != cudaSuccess) return (int)me;
  }

#define CALL(VV, LL)                                                                                        \
  seg_chunk_kernel<VV, LL><<<(unsigned)blocks1, kSegThreads, 0, s>>>(sorted_keys, perm, n, row_grads, ap,   \
                                                                     part_first, part_last);                \
  if (groups2 > 0) {                                                                                        \
    seg_boundary_kernel<VV, LL><<<(unsigned)blocks2, kSegThreads, 0, s>>>(                                  \
        sorted_keys, n, n_chunks, ap, part_first, part_last, long_count, long_list);                        \
    if (n_chunks > kSerialMax)                                                                              \
      seg_long_kernel<VV, LL><<<long_blocks, kLongThreads, 0, s>>>(sorted_keys, ap, part_first, part_last,  \
                                                                   long_count, long_list);                  \
  }
  RSB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  RSB_CHECK_LAUNCH();
  note_launch(groups2 > 0 ? (n_chunks > kSerialMax ? 3 : 2) : 1);
  return RSB_OK;
}
