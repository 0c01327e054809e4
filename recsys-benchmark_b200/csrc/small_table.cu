// Gradient of a SMALL table (QR emb1 has `divider` rows: 2/5/20 in the shipped configs,
// int(sqrt(N)) by default; src/models/embeddings/qr_embedding.py:47-59).  All B*F lookups
// collapse onto those few rows, so instead of sorting, every CTA keeps the whole table
// gradient in shared memory, accumulates its slice of the lookups there and writes one
// partial table; a second launch adds the partials in CTA order (deterministic across
// CTAs; the shared-memory adds inside a CTA are float atomics).
#include "common.cuh"

namespace rsb {

constexpr int kSmallThreads = 256;

template <int V, int LPR>
__global__ void __launch_bounds__(kSmallThreads) small_table_partial_kernel(
    const long long* __restrict__ keys, long long n, long long key_div, long long key_mod,
    const float* __restrict__ rg, int E, long long n_rows, float* __restrict__ partials) {
  extern __shared__ float acc[];
  const long long tot = n_rows * E;
  for (long long i = threadIdx.x; i < tot; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  constexpr int GPW = kWarp / LPR;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, c = lane % LPR;
  const int d0 = c * V;
  const bool cact = d0 < E;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32 * GPW + g;
  const long long ngroups = (long long)gridDim.x * blockDim.x / 32 * GPW;
  for (long long p = gid; p < n; p += ngroups) {
    long long k = __ldg(keys + p);
    if (key_div > 1) k = k / key_div;
    if (key_mod > 0) k = k % key_mod;
    if (cact) {
      FV<V> v = ldg<V>(rg + p * E + d0);
#pragma unroll
      for (int i = 0; i < V; ++i) atomicAdd(&acc[k * E + d0 + i], v.v[i]);
    }
  }
  __syncthreads();
  float* out = partials + (long long)blockIdx.x * tot;
  for (long long i = threadIdx.x; i < tot; i += blockDim.x) out[i] = acc[i];
}

// Tiny tables (<= R rows, R = 8 or 32): every lane group keeps one accumulator per table row
// in REGISTERS (the row id only selects which one is added to), so there are no atomics at
// all; lane groups -> warp (shuffles) -> CTA (shared memory, fixed order) -> one partial per CTA.
template <int V, int LPR, int R>
__global__ void __launch_bounds__(kSmallThreads) tiny_table_partial_kernel(
    const long long* __restrict__ keys, long long n, long long key_div, long long key_mod,
    const float* __restrict__ rg, int E, int n_rows, float* __restrict__ partials) {
  constexpr int GPW = kWarp / LPR;
  constexpr int NW = kSmallThreads / 32;
  __shared__ float red[NW][R][LPR * V];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane / LPR, c = lane % LPR;
  const int d0 = c * V;
  const bool cact = d0 < E;
  FV<V> acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = FV<V>::zero();
  // contiguous slice per CTA, groups stride inside it: fixed assignment -> deterministic
  const long long per_cta = (n + gridDim.x - 1) / gridDim.x;
  const long long beg = per_cta * blockIdx.x;
  long long end = beg + per_cta;
  if (end > n) end = n;
  const int gib = wib * GPW + g;            // group index in the CTA
  constexpr int GPB = NW * GPW;
  for (long long p = beg + gib; p < end; p += GPB) {
    long long k = __ldg(keys + p);
    if (key_div > 1) k = k / key_div;
    if (key_mod > 0) k = k % key_mod;
    FV<V> v = FV<V>::zero();
    if (cact) v = ldg<V>(rg + p * E + d0);
    const int kk = (int)k;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool hit = (kk == r);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[r].v[i] += hit ? v.v[i] : 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int off = LPR; off < kWarp; off <<= 1) {
      FV<V> o = shfl_xor<V>(acc[r], off);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[r].v[i] += o.v[i];
    }
    if (g == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) red[wib][r][c * V + i] = acc[r].v[i];
    }
  }
  __syncthreads();
  const long long tot = (long long)n_rows * E;
  float* out = partials + (long long)blockIdx.x * tot;
  for (int i = threadIdx.x; i < tot; i += blockDim.x) {
    const int r = i / E, d = i - r * E;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w][r][d];
    out[i] = s;
  }
}

__global__ void small_table_reduce_kernel(const float* __restrict__ partials, int nblk, long long tot,
                                          float* __restrict__ dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tot) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partials[(long long)b * tot + i];
  dst[i] = s;
}

static int small_blocks(long long n) {
  long long want = (n + 2047) / 2048;
  long long cap = 8ll * sm_count();   // enough resident warps to keep the row-gradient loads in flight
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace rsb

using namespace rsb;

// Upper bound on table bytes kept in shared memory per CTA.
static const long long kSmallTableMaxBytes = 96 * 1024;

extern "C" RSB_API int64_t rsb_small_table_workspace_bytes(int64_t n_rows, int32_t E) {
  if (n_rows <= 0 || E <= 0) return 0;
  if (n_rows * E * 4 > kSmallTableMaxBytes) return -1;  // not a small table: use the sorted path
  return (int64_t)8 * sm_count() * n_rows * E * 4 + 256;
}

extern "C" RSB_API int rsb_small_table_grad(const int64_t* keys, int64_t n, int64_t key_div, int64_t key_mod,
                                    const float* row_grads, int32_t E, int64_t n_rows, float* dst, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  if (n < 0 || E <= 0 || n_rows <= 0 || !dst) return RSB_ERR_BAD_ARG;
  const long long tot = n_rows * E;
  if (tot * 4 > kSmallTableMaxBytes) return RSB_ERR_UNSUPPORTED;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(dst, 0, tot * 4, s);
    return (int)e;
  }
  if (!keys || !row_grads || !workspace) return RSB_ERR_BAD_ARG;
  if (workspace_bytes < rsb_small_table_workspace_bytes(n_rows, E)) return RSB_ERR_WORKSPACE;
  RowShape sh = row_shape(E, aligned16(row_grads));
  if (!sh.ok) return RSB_ERR_UNSUPPORTED;
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  const int nblk = small_blocks(n);
  const size_t smem = (size_t)tot * 4;
  if (n_rows <= 32 && sh.V * sh.LPR <= 32) {
    const long long* k64 = reinterpret_cast<const long long*>(keys);
#define CALLT(VV, LL)                                                                                          \
  if constexpr (VV * LL <= 32) {                                                                               \
    if (n_rows <= 8)                                                                                           \
      tiny_table_partial_kernel<VV, LL, 8><<<nblk, kSmallThreads, 0, s>>>(k64, n, key_div, key_mod,           \
                                                                          row_grads, E, (int)n_rows, partials); \
    else                                                                                                       \
      tiny_table_partial_kernel<VV, LL, 32><<<nblk, kSmallThreads, 0, s>>>(k64, n, key_div, key_mod,          \
                                                                           row_grads, E, (int)n_rows, partials); \
  }
    RSB_DISPATCH_SHAPE(sh, CALLT);
#undef CALLT
    RSB_CHECK_LAUNCH();
    small_table_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(partials, nblk, tot, dst);
    RSB_CHECK_LAUNCH();
    note_launch(2);
    return RSB_OK;
  }
#define CALL(VV, LL)                                                                                          \
  if (smem > 48 * 1024)                                                                                       \
    cudaFuncSetAttribute(small_table_partial_kernel<VV, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                         (int)smem);                                                                          \
  small_table_partial_kernel<VV, LL><<<nblk, kSmallThreads, smem, s>>>(                                       \
      reinterpret_cast<const long long*>(keys), n, key_div, key_mod, row_grads, E, n_rows, partials)
  RSB_DISPATCH_SHAPE(sh, CALL);
#undef CALL
  RSB_CHECK_LAUNCH();
  small_table_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(partials, nblk, tot, dst);
  RSB_CHECK_LAUNCH();
  note_launch(2);
  return RSB_OK;
}
