// The two ends of the training step that sit right outside the gather / dense tail:
//  * record unpacking: the lmdb caches hold one uint32 record [label, id_0 .. id_{F-1}] per sample
//    (src/dataset/criteo/criteo_torchfm.py:72-93, avazu_fm.py:78-97, kdd_dataset.py:53-74); a staged block of records is
//    turned into the [B,F] int32 id batch the gather reads and the float labels BCEWithLogits wants
//    (src/trainer/deepfm.py:44-52: inputs.to(device), labels.to(device), labels.float()) in one pass on the copy stream.
//  * the dense Adam update of get_optimizers' non-sparse branch (src/models/deepfm.py:155-172): every parameter of the
//    model in ONE launch (chunk table, no per-tensor launches), arithmetic of torch.optim.Adam (L2-style weight decay).
#include "common.cuh"

namespace rsb {

__global__ void __launch_bounds__(256) records_unpack_kernel(const unsigned* __restrict__ rec, long long B, int F,
                                                             int* __restrict__ ids, float* __restrict__ labels,
                                                             FastDiv fd_f) {
  // one thread per id: coalesced on both sides (the record stride F+1 and the id stride F differ by one word)
  const long long total = B * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long b;
    int f;
    if (total < (1ll << 32)) {
      const unsigned q = fastdiv((unsigned)i, fd_f);
      b = q;
      f = (int)((unsigned)i - q * (unsigned)F);
    } else {
      b = i / F;
      f = (int)(i - b * F);
    }
    const unsigned v = __ldg(rec + b * (F + 1) + 1 + f);
    ids[i] = (int)v;
    if (f == 0 && labels) labels[b] = (float)__ldg(rec + b * (F + 1));
  }
}

struct AdamScalars {
  float lr_over_bc1;     // lr / (1 - beta1^step)
  float bc2_sqrt;        // sqrt(1 - beta2^step)
  float beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s) {
  // torch/optim/adam.py::_single_tensor_adam, operation by operation:
  //   grad = grad.add(param, alpha=weight_decay); exp_avg.lerp_(grad, 1 - beta1);
  //   exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2);
  //   denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps); param.addcdiv_(exp_avg, denom, value=-step_size)
  g = fmaf(s.weight_decay, p, g);
  m = fmaf(s.one_minus_beta1, g - m, m);
  v = fmaf(s.one_minus_beta2 * g, g, v * s.beta2);
  const float denom = __fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt) + s.eps;
  p = fmaf(-s.lr_over_bc1, __fdiv_rn(m, denom), p);
}

constexpr int kAdamChunk = RSB_ADAM_CHUNK;

// the descriptors travel in the kernel's parameter space (like torch's multi_tensor_apply): gradient buffers are
// re-allocated by every backward, so a device-resident table would need a host->device copy per step
struct AdamTensors {
  rsb_adam_tensor t[RSB_ADAM_MAX_TENSORS];
};

__global__ void __launch_bounds__(256) adam_dense_kernel(const __grid_constant__ AdamTensors tensors,
                                                         const int2* __restrict__ block_map, AdamScalars s) {
  const int2 bm = block_map[blockIdx.x];                 // (tensor, chunk)
  const rsb_adam_tensor& t = tensors.t[bm.x];
  const long long start = (long long)bm.y * kAdamChunk;
  long long n = t.numel - start;
  if (n > kAdamChunk) n = kAdamChunk;
  float* p = t.param + start;
  const float* g = t.grad + start;
  float* m = t.exp_avg + start;
  float* v = t.exp_avg_sq + start;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15u) == 0;
  if (vec) {
    const int n4 = (int)(n >> 2);
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 mm = reinterpret_cast<float4*>(m)[i];
      float4 vv = reinterpret_cast<float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, s);
      adam_one(pp.y, gg.y, mm.y, vv.y, s);
      adam_one(pp.z, gg.z, mm.z, vv.z, s);
      adam_one(pp.w, gg.w, mm.w, vv.w, s);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) adam_one(p[i], __ldg(g + i), m[i], v[i], s);
  } else {
    for (int i = threadIdx.x; i < n; i += 256) adam_one(p[i], __ldg(g + i), m[i], v[i], s);
  }
}

}  // namespace rsb

using namespace rsb;

extern "C" RSB_API int rsb_records_unpack(const void* records, int64_t B, int32_t F, int32_t* ids_out, float* labels_out,
                                          void* stream) {
  if (B < 0 || F <= 0) return RSB_ERR_BAD_ARG;
  if (B == 0) return RSB_OK;
  if (!records || !ids_out) return RSB_ERR_BAD_ARG;
  const long long total = B * F;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  records_unpack_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned*>(records), B, F, ids_out, labels_out, make_fastdiv((unsigned long long)F));
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_adam_dense(const rsb_adam_tensor* h_tensors, int32_t n_tensors, const int32_t* block_map,
                                      int64_t n_blocks, double lr, double beta1, double beta2, double eps,
                                      double weight_decay, int64_t step, void* stream) {
  if (n_blocks < 0 || n_tensors < 0 || step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0))
    return RSB_ERR_BAD_ARG;
  if (n_blocks == 0 || n_tensors == 0) return RSB_OK;
  if (!h_tensors || !block_map || n_blocks > 0x7fffffffll) return RSB_ERR_BAD_ARG;
  if (n_tensors > RSB_ADAM_MAX_TENSORS) return RSB_ERR_UNSUPPORTED;
  AdamTensors tensors;
  for (int i = 0; i < n_tensors; ++i) {
    tensors.t[i] = h_tensors[i];
    if (!tensors.t[i].param || !tensors.t[i].grad || !tensors.t[i].exp_avg || !tensors.t[i].exp_avg_sq || tensors.t[i].numel < 0)
      return RSB_ERR_BAD_ARG;
  }
  for (int i = n_tensors; i < RSB_ADAM_MAX_TENSORS; ++i) tensors.t[i] = rsb_adam_tensor{nullptr, nullptr, nullptr, nullptr, 0};
  // the scalars exactly as the Python optimizer forms them (double), rounded once
  AdamScalars s;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  s.lr_over_bc1 = (float)(lr / bc1);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.beta2 = (float)beta2;
  s.one_minus_beta1 = (float)(1.0 - beta1);
  s.one_minus_beta2 = (float)(1.0 - beta2);
  s.eps = (float)eps;
  s.weight_decay = (float)weight_decay;
  adam_dense_kernel<<<(unsigned)n_blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      tensors, reinterpret_cast<const int2*>(block_map), s);
  RSB_CHECK_LAUNCH();
  note_launch(1);
  return RSB_OK;
}
