// Stable LSD radix sort of the lookups by target row (backward stage 2).
//
// keys are row ids < n_rows (so only bit_length(n_rows-1) bits are sorted, in passes of
// <= 11 bits: two passes for tables up to 4 M rows), values are the lookup positions.
// Three launches per pass:
//   histogram (per-CTA digit counts, digit-major)
//   -> row scan (one CTA per digit: exclusive prefix over the CTAs + the digit total)
//   -> stable scatter (digit bases scanned in shared memory, match-any ranking).
// The first pass reads the int64 ids directly and applies the optional QR index
// transform (key = id / key_div or id % key_mod, qr_embedding.py:96-97, bit exact on
// non-negative ids), so no separate key-extraction pass is needed.
#include <cstdlib>

#include "common.cuh"

namespace rsb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // keys per CTA
constexpr int kRadixBitsMax = 11;
constexpr int kRadixMax = 1 << kRadixBitsMax;

struct SortSrc {
  const long long* keys64;  // pass 0 source (or nullptr)
  long long key_div, key_mod;
  FastDiv fd_div, fd_mod;   // for ids < 2^32 (every table the gather can address today)
  const unsigned* keys32;  // later passes
  const unsigned* vals32;
};

__device__ __forceinline__ unsigned sort_key_at(const SortSrc& s, long long i) {
  if (s.keys64) {
    long long k = __ldg(s.keys64 + i);
    if (((unsigned long long)k >> 32) == 0 && ((unsigned long long)(s.key_div | s.key_mod) >> 32) == 0) {
      unsigned r = (unsigned)k;          // exact 32-bit path: a multiply-high instead of an emulated 64-bit divide
      if (s.key_div > 1) r = fastdiv(r, s.fd_div);
      if (s.key_mod > 0) r = r - fastdiv(r, s.fd_mod) * (unsigned)s.key_mod;
      return r;
    }
    if (s.key_div > 1) k = k / s.key_div;
    if (s.key_mod > 0) k = k % s.key_mod;
    return (unsigned)k;
  }
  return __ldg(s.keys32 + i);
}

// Each warp owns a contiguous run of kSortItems*32 keys of the tile, visited in kSortItems
// rounds of 32 consecutive keys: (warp, round, lane) order == memory order, which is what
// makes the ranking below stable.
__device__ __forceinline__ long long sort_pos(long long tile_base, int warp, int round, int lane) {
  return tile_base + (long long)warp * (kSortItems * 32) + round * 32 + lane;
}

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(SortSrc src, long long n, int shift, int radix_bits,
                                                                 unsigned* hist, int nblk) {
  __shared__ unsigned h[kRadixMax];
  const int radix = 1 << radix_bits;
  for (int i = threadIdx.x; i < radix; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 4
  for (int r = 0; r < kSortItems; ++r) {
    long long p = sort_pos(base, warp, r, lane);
    if (p < n) {
      unsigned d = (sort_key_at(src, p) >> shift) & (radix - 1);
      atomicAdd(&h[d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < radix; i += blockDim.x) hist[(long long)i * nblk + blockIdx.x] = h[i];
}

// One CTA per digit d: exclusive prefix of hist[d][0..nblk) in place, digit total to totals[d].
__global__ void __launch_bounds__(256) sort_rowscan_kernel(unsigned* __restrict__ hist, int nblk,
                                                          unsigned* __restrict__ totals) {
  __shared__ unsigned warp_tot[8];
  __shared__ unsigned carry_s;
  unsigned* row = hist + (long long)blockIdx.x * nblk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 256 * 4) {
    const int i0 = base + threadIdx.x * 4;
    unsigned v[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = (i0 + i < nblk) ? row[i0 + i] : 0u;
      sum += v[i];
    }
    unsigned incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned t = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) wbase += (w < warp) ? warp_tot[w] : 0u;
    unsigned run = carry_s + wbase + (incl - sum);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i0 + i < nblk) row[i0 + i] = run;
      run += v[i];
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_s = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(kSortThreads, 3) sort_scatter_kernel(SortSrc src, long long n, int shift,
                                                                    int radix_bits, const unsigned* hist_scanned,
                                                                    const unsigned* totals, int nblk,
                                                                    unsigned* out_keys, unsigned* out_vals) {
  constexpr int NW = kSortThreads / 32;
  extern __shared__ unsigned sort_smem[];
  const int radix = 1 << radix_bits;
  unsigned* dbase = sort_smem;                 // [radix] exclusive scan of the digit totals
  unsigned* cnt = sort_smem + radix;           // [NW][radix] per-warp digit counts (sized by the actual radix)
  __shared__ unsigned wtot[NW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned* wcnt = cnt + warp * radix;
  for (int i = threadIdx.x; i < NW * radix; i += blockDim.x) cnt[i] = 0;
  {
    // digit bases: thread t scans its radix/256 consecutive totals, then a 256-wide block scan
    const int per = (radix + kSortThreads - 1) / kSortThreads;
    unsigned loc[kRadixMax / kSortThreads];
    unsigned sum = 0;
#pragma unroll
    for (int i = 0; i < kRadixMax / kSortThreads; ++i) {
      const int d = threadIdx.x * per + i;
      loc[i] = (i < per && d < radix) ? __ldg(totals + d) : 0u;
      sum += loc[i];
    }
    unsigned incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned t = __shfl_up_sync(kFull, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) wbase += (w < warp) ? wtot[w] : 0u;
    unsigned run = wbase + (incl - sum);
#pragma unroll
    for (int i = 0; i < kRadixMax / kSortThreads; ++i) {
      const int d = threadIdx.x * per + i;
      if (i < per && d < radix) dbase[d] = run;
      run += loc[i];
    }
  }
  __syncthreads();

  const long long base = (long long)blockIdx.x * kSortTile;
  // this CTA's scanned histogram column (a strided, latency-bound read): issued now, consumed after the ranking
  unsigned hs[kRadixMax / kSortThreads];
#pragma unroll
  for (int i = 0; i < kRadixMax / kSortThreads; ++i) {
    const int d = threadIdx.x + i * kSortThreads;
    hs[i] = (d < radix) ? __ldg(hist_scanned + (long long)d * nblk + blockIdx.x) : 0u;
  }
  unsigned key[kSortItems], dr[kSortItems];  // dr = digit << 16 | rank within the warp's digit class (< 512)
  // all loads of the tile are issued before the (serial, shuffle-bound) ranking rounds
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    long long p = sort_pos(base, warp, r, lane);
    key[r] = (p < n) ? sort_key_at(src, p) : 0xffffffffu;
  }
  // Ranking: lanes holding the same digit are found with one ballot per digit bit (plus one for validity)
  // instead of match.any, whose cost grows with the number of distinct values in the warp - with 512 digits
  // nearly all 32 lanes differ, and the kernel was bound by exactly that instruction.
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const bool valid = sort_pos(base, warp, r, lane) < n;
    const unsigned d = (key[r] >> shift) & (radix - 1);
    unsigned peers = __ballot_sync(kFull, valid);
    for (int bit = 0; bit < radix_bits; ++bit) {
      const bool one = (d >> bit) & 1u;
      const unsigned m = __ballot_sync(kFull, one);
      peers &= one ? m : ~m;
    }
    // the class leader reserves the class's ranks with ONE shared-memory atomic whose return value is the
    // count so far; a warp's shared-memory operations complete in issue order, so the rounds need no barrier
    // between them and their atomics / shuffles pipeline (the kernel was bound by that round trip)
    const int leader = valid ? (__ffs(peers) - 1) : lane;
    unsigned pre = 0;
    if (valid && lane == leader) pre = atomicAdd(&wcnt[d], (unsigned)__popc(peers));
    pre = __shfl_sync(kFull, pre, leader);
    dr[r] = (d << 16) | (pre + __popc(peers & ((1u << lane) - 1u)));
  }
  __syncthreads();
  // exclusive scan over warps for each digit, plus this CTA's global base
#pragma unroll
  for (int i = 0; i < kRadixMax / kSortThreads; ++i) {
    const int d = threadIdx.x + i * kSortThreads;
    if (d < radix) {
      unsigned run = dbase[d] + hs[i];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        unsigned t = cnt[w * radix + d];
        cnt[w * radix + d] = run;
        run += t;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    long long p = sort_pos(base, warp, r, lane);
    if (p < n) {
      unsigned dst = wcnt[dr[r] >> 16] + (dr[r] & 0xffffu);
      out_keys[dst] = key[r];
      out_vals[dst] = src.vals32 ? __ldg(src.vals32 + p) : (unsigned)p;
    }
  }
}

static int bit_length(unsigned long long x) {
  int b = 0;
  while (x) {
    ++b;
    x >>= 1;
  }
  return b;
}

static long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int64_t rsb_sort_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  long long nblk = (n + kSortTile - 1) / kSortTile;
  if (nblk < 1) nblk = 1;
  long long bytes = 0;
  bytes += align_up(2 * n * 4, 256);              // ping-pong keys
  bytes += align_up(2 * n * 4, 256);              // ping-pong values
  bytes += align_up(nblk * kRadixMax * 4, 256);   // histogram
  bytes += align_up(kRadixMax * 4, 256);          // digit totals
  return bytes + 256;
}

extern "C" RSB_API int rsb_sort_rows(const int64_t* keys, int64_t n, int64_t n_rows, int64_t key_div, int64_t key_mod,
                             uint32_t* sorted_keys, uint32_t* perm, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  // keys must stay below 0xffffffff: the segmented reduction uses it as its 'no key' sentinel
  if (n < 0 || n_rows <= 0 || n >= (1ll << 32) || n_rows > 0xffffffffll) return RSB_ERR_BAD_ARG;
  if (n == 0) return RSB_OK;
  if (keys == nullptr || sorted_keys == nullptr || perm == nullptr || workspace == nullptr) return RSB_ERR_BAD_ARG;
  if (workspace_bytes < rsb_sort_workspace_bytes(n)) return RSB_ERR_WORKSPACE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long nblk = (n + kSortTile - 1) / kSortTile;

  char* w = reinterpret_cast<char*>(workspace);
  w = reinterpret_cast<char*>(align_up(reinterpret_cast<long long>(w), 256));
  unsigned* kbuf[2];
  unsigned* vbuf[2];
  kbuf[0] = reinterpret_cast<unsigned*>(w);
  kbuf[1] = kbuf[0] + n;
  w += align_up(2 * n * 4, 256);
  vbuf[0] = reinterpret_cast<unsigned*>(w);
  vbuf[1] = vbuf[0] + n;
  w += align_up(2 * n * 4, 256);
  unsigned* hist = reinterpret_cast<unsigned*>(w);
  w += align_up(nblk * kRadixMax * 4, 256);
  unsigned* totals = reinterpret_cast<unsigned*>(w);

  int bits = bit_length((unsigned long long)(n_rows - 1));
  if (bits < 1) bits = 1;
  static const int max_bits = [] {   // tuning knob: digit width cap (<= kRadixBitsMax)
    const char* v = getenv("RSB_SORT_MAX_BITS");
    int m = v ? atoi(v) : kRadixBitsMax;
    return (m < 4 || m > kRadixBitsMax) ? kRadixBitsMax : m;
  }();
  const int passes = (bits + max_bits - 1) / max_bits;
  const int rb = (bits + passes - 1) / passes;  // radix bits per pass (<= 11)
  const size_t scatter_smem_max = (size_t)(kRadixMax + (kSortThreads / 32) * kRadixMax) * sizeof(unsigned);
  const size_t scatter_smem = ((size_t)(1 << rb) * (1 + kSortThreads / 32)) * sizeof(unsigned);
  {
    // function attributes are per device: set once for each device this process launches on
    static bool attr_set[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
    if (dev < 0 || !attr_set[dev]) {
      cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_max);
      if (dev >= 0) attr_set[dev] = true;
    }
  }

  SortSrc src;
  src.keys64 = reinterpret_cast<const long long*>(keys);
  src.key_div = key_div;
  src.key_mod = key_mod;
  src.fd_div = make_fastdiv((unsigned long long)(key_div > 0 ? key_div : 1));
  src.fd_mod = make_fastdiv((unsigned long long)(key_mod > 0 ? key_mod : 1));
  src.keys32 = nullptr;
  src.vals32 = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool last = (p == passes - 1);
    unsigned* ok = last ? sorted_keys : kbuf[p & 1];
    unsigned* ov = last ? perm : vbuf[p & 1];
    const int shift = p * rb;
    sort_hist_kernel<<<(unsigned)nblk, kSortThreads, 0, s>>>(src, n, shift, rb, hist, (int)nblk);
    RSB_CHECK_LAUNCH();
    sort_rowscan_kernel<<<1 << rb, 256, 0, s>>>(hist, (int)nblk, totals);
    RSB_CHECK_LAUNCH();
    sort_scatter_kernel<<<(unsigned)nblk, kSortThreads, scatter_smem, s>>>(src, n, shift, rb, hist, totals, (int)nblk,
                                                                          ok, ov);
    RSB_CHECK_LAUNCH();
    note_launch(3);
    src.keys64 = nullptr;
    src.keys32 = ok;
    src.vals32 = ov;
  }
  return RSB_OK;
}
