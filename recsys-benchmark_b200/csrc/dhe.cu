// Deep hash embedding encoder: the k-dimensional universal-hash code of every looked-up
// id is generated in the kernel instead of being gathered from a cached [N, k] fp32 table
// (src/models/embeddings/dh_embedding.py:194-236 builds that table: 4.4 GB at Criteo shape, k = 1024).
//
//   h      = (slope[j] * (id + prefix + 1) + bias[j]) mod prime[j] mod m      (int64, Python-style mod:
//            the result takes the sign of the divisor, i.e. is non-negative)      dh_embedding.py:205-207
//   code   = float(h) / float(m - 1) * 2 - 1                                   (fp32)  dh_embedding.py:211-213
//
// Bit-exact with the reference: integer arithmetic is exact, the int64 -> fp32 conversions are exact
// (h < m <= 2^24), the division is IEEE round-to-nearest and *2 - 1 cannot be contracted into anything
// that rounds differently (the product by 2 is exact).
//
// The 64-bit modulo is the cost (emulated: ~100 instructions): when |slope|, |bias| < 2^31, prime < 2^31 and
// id + prefix + 1 < 2^30 (the caller states the bound), |x| < 2^62 and the quotient is taken from one fp64
// multiply by 1/prime (error < 1 for these magnitudes), the remainder fixed up exactly in int64.
#include "common.cuh"

namespace rsb {

__device__ __forceinline__ long long python_mod(long long x, long long p) {
  long long r = x % p;
  return (r < 0) ? r + p : r;
}

template <bool FAST>
__global__ void __launch_bounds__(256) dhe_encode_kernel(const void* __restrict__ ids, int ids_i32, long long n,
                                                         long long prefix, const long long* __restrict__ slopes,
                                                         const long long* __restrict__ bias,
                                                         const long long* __restrict__ primes, int k, long long m,
                                                         float* __restrict__ out) {
  // one CTA row of threads walks the code dimension (coalesced stores), grid-stride over the ids
  const float denom = (float)(m - 1);
  for (int j0 = 0; j0 < k; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    const bool jact = j < k;
    const long long a = jact ? __ldg(slopes + j) : 1;
    const long long b = jact ? __ldg(bias + j) : 0;
    const long long p = jact ? __ldg(primes + j) : 1;
    const double inv_p = 1.0 / (double)p;
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
      const long long id = ids_i32 ? (long long)__ldg(reinterpret_cast<const int*>(ids) + i)
                                   : __ldg(reinterpret_cast<const long long*>(ids) + i);
      // two's-complement wrap-around like torch's int64 arithmetic (signed overflow would be undefined in C++)
      const long long x = (long long)((unsigned long long)a * (unsigned long long)(id + prefix + 1) +
                                      (unsigned long long)b);
      long long r;
      if (FAST) {
        const long long q = (long long)((double)x * inv_p);   // |error| < 1: x < 2^62, p >= 2
        r = x - q * p;
        r += (r < 0) ? p : 0;
        r += (r < 0) ? p : 0;
        r -= (r >= p) ? p : 0;
        r -= (r >= p) ? p : 0;
        r = (long long)((unsigned)r % (unsigned)m);           // r < p < 2^31, m <= 2^24
      } else {
        r = python_mod(python_mod(x, p), m);
      }
      if (jact) {
        const float e = __fdiv_rn((float)r, denom);
        out[i * (long long)k + j] = __fsub_rn(__fmul_rn(e, 2.0f), 1.0f);
      }
    }
  }
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int rsb_dhe_encode(const void* ids, int32_t ids_is_i32, int64_t n, int64_t prefix,
                                      const int64_t* slopes, const int64_t* bias, const int64_t* primes, int32_t k,
                                      int64_t m, int32_t small_operands, float* out, void* stream) {
  if (n < 0 || k <= 0 || m < 2 || m > (1ll << 24)) return RSB_ERR_BAD_ARG;
  if (n == 0) return RSB_OK;
  if (!ids || !slopes || !bias || !primes || !out) return RSB_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = (k >= 256) ? 256 : ((k + 31) / 32 * 32);
  long long blocks = n;
  const long long cap = (long long)sms * (2048 / threads) * 4;
  if (blocks > cap) blocks = cap;
  const long long* sl = reinterpret_cast<const long long*>(slopes);
  const long long* bi = reinterpret_cast<const long long*>(bias);
  const long long* pr = reinterpret_cast<const long long*>(primes);
  if (small_operands)
    dhe_encode_kernel<true><<<(unsigned)blocks, threads, 0, s>>>(ids, ids_is_i32, n, prefix, sl, bi, pr, k, m, out);
  else
    dhe_encode_kernel<false><<<(unsigned)blocks, threads, 0, s>>>(ids, ids_is_i32, n, prefix, sl, bi, pr, k, m, out);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}
