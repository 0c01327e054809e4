// fp32-accurate GEMM on the 5th-generation tensor cores, hand-written for sm_100a (TMA + tcgen05 + TMEM).
//
// What it replaces: the true-fp32 cuBLAS SGEMMs the reference runs for nn.Linear in the dense tails
// (src/models/deepfm.py:55-66, src/models/dcn.py:56-66) and for the DCN-Mix expert projections
// (src/models/layer_dcn.py:20-23) - TF32 is off there, so a plain bf16 / TF32 MMA would break the 1e-5 gate.
//
// Arithmetic: every fp32 operand x is held as three bf16 "planes" x0 + x1 + x2 (x0 = bf16(x), x1 = bf16(x - x0),
// x2 = bf16(x - x0 - x1): 3 x 8 mantissa bits).  A product a*b is the sum of the six plane products >= 2^-16 |a||b|
//      band 3: a0 b2, a1 b1, a2 b0      band 2: a0 b1, a1 b0      band 1: a0 b0
// (the three dropped ones are <= 2^-24 |a||b|).  bf16 x bf16 products are exact in fp32; what costs accuracy is the
// accumulation: the tensor core truncates the running fp32 sum once per MMA.  Therefore
//   * the planes are produced ONCE by whoever writes the operand (the split kernel below, or a producer's epilogue),
//     not re-split by every CTA for every tile, and arrive by TMA as plain bf16 tiles;
//   * within a 32-deep K group the 12 MMAs are issued smallest band first into one TMEM accumulator (the small
//     bands are added while the sum is still small), and after every K group the accumulator is drained into fp32
//     REGISTERS with round-to-nearest adds by the epilogue warps (two TMEM buffers: the MMAs of group g+1 run
//     while group g is drained).  Only two truncations of full-size sums per 32 k: error vs fp64 ~1e-7.
//
// Kernel: persistent, one CTA per SM, 12 warps:
//   warp 0      TMA producer (one lane): 3-plane boxes of A and B per stage -> smem ring, mbarrier complete_tx
//   warp 1      MMA issuer (one lane): tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), tcgen05.commit
//   warp 2      TMEM allocation
//   warps 4-11  drain TMEM -> registers every K group (tcgen05.ld 32x32b), final epilogue (alpha, beta*C, bias)
// Operands may be K-major ([rows, K] row-major: activations for the forward / dX GEMMs, nn.Linear weights) with the
// 64-byte swizzle, or MN-major ([K, rows] row-major: both operands of a weight-gradient GEMM, whose reduction runs
// over the batch) with the 128-byte swizzle; the smem / instruction descriptors carry the majorness.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace pg {

constexpr int kBlockM = 128;        // rows of D per CTA tile = TMEM lanes
constexpr int kBlockK = 32;         // K per stage = one drain group (2 MMA K steps of 16)
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;        // first epilogue warp
constexpr int kTmemCols = 512;
constexpr int kTmemBufStride = 256; // columns between the two accumulator buffers
constexpr int kMaxStages = 6;
constexpr uint32_t kSmemLimit = 227 * 1024;

struct Operand {
  int mn_major;          // 0: K-major [rows, K]; 1: MN-major [K, rows]
  int batch_row_step;    // added to the ROW coordinate of the stored matrix per batch index
  int batch_col_step;    // added to the COLUMN coordinate per batch index
};

struct Params {
  int M, N, K, batch, splits;
  int n_tile, m_tiles, n_tiles;
  int k_blocks;          // ceil(K / 32)
  int stages;
  int drain;             // K groups (of 32) accumulated in TMEM between two drains into registers
  int prefetch;          // A boxes requested into L2 this many K groups ahead of the shared-memory loads (0 = off)
  long long* dbg;        // RSB_GEMM_DEBUG: per-CTA cycle counters (development aid)
  int tma_store;         // fp32 epilogue through shared memory + TMA stores (map_d); 0: direct stores
  uint32_t epi_offset;   // byte offset of the epilogue staging buffers in dynamic shared memory
  int np;                // planes per operand: 3 (bf16) or 2 (fp16, scaled)
  rsb::PlaneFmt fmt_a, fmt_b;              // FP16X2: the operands hold A * sa, B * sb; the epilogue divides by sa * sb
  uint32_t a_stage_bytes, b_stage_bytes;   // shared-memory footprint of one stage in one CTA
  uint32_t tx_bytes;                       // bytes the TMA loads of one CTA deliver per stage
  Operand a, b;
  // epilogue
  float* D;
  long long ldd, d_batch_stride;
  const float* C;
  const float* bias;
  float alpha, beta;
  float* partial;        // split-K: [splits][batch][M][N] raw sums (then splitk_reduce_kernel)
  // fused epilogues (rsb_gemm_epilogue)
  int epi_mode;
  __nv_bfloat16* out_planes;
  long long out_ld, out_plane_stride;
  int ones_col;
  unsigned char* mask;
  float drop_scale;
  float* d_amax;         // fp32 outputs: *d_amax is raised to max |D| (the bound an FP16X2 consumer of D needs)
  float* bn_part;        // TMA-store epilogue: per 32-row group shifted column sums of D, [row_groups][3][N] (see below)
};

// 8 consecutive values of one row -> three 16-byte bf16 plane stores (the fused plane epilogues write BF16X3)
__device__ __forceinline__ void store_planes8(__nv_bfloat16* dst, long long plane_stride, const float* v) {
  rsb::store_planes<8>(dst, plane_stride, v, rsb::kPlanesBf16x3, 1.f);
}

// ------------------------------------------------------------------------------------------- PTX wrappers ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA prefetch of a box into L2 (no shared memory, no barrier): the later cp.async.bulk.tensor load of the same box hits L2
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// TMA store of a [rows x 16 fp32] box staged in shared memory (64-byte swizzle); rows / columns past the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1)
               : "memory");
}

// shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"): start address, leading / stride byte
// offsets (all >> 4), descriptor version 1 (Blackwell) at bit 46, swizzle mode at bits 61-63
constexpr uint64_t kSwizzle128 = 2, kSwizzle64 = 4;
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t swizzle) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (swizzle << 61);
}

// instruction descriptor for kind::f16: D fp32, A / B bf16, majorness bits, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ kernel ---
struct Tile {
  int m_blk, n_blk, batch, split;
  int kb0, kb1;   // K-block range of this split
};

__device__ __forceinline__ bool get_tile(const Params& p, int idx, Tile& t) {
  const int per_split = p.m_tiles * p.n_tiles * p.batch;
  if (idx >= per_split * p.splits) return false;
  t.split = idx / per_split;
  int r = idx - t.split * per_split;
  // N fastest: the N tiles of one M block run at the same time on neighbouring SMs, so the A block (the big operand:
  // activations) comes from DRAM once and from L2 for the other N tiles; B (weights) is small and L2-resident anyway.
  // (M fastest re-read all of A from DRAM once per N tile: 508 MB instead of 245 MB on 65536 x 400 x 624.)
  t.n_blk = r % p.n_tiles;
  r /= p.n_tiles;
  t.m_blk = r % p.m_tiles;
  t.batch = r / p.m_tiles;
  const int base = p.k_blocks / p.splits, rem = p.k_blocks % p.splits;
  t.kb0 = t.split * base + (t.split < rem ? t.split : rem);
  t.kb1 = t.kb0 + base + (t.split < rem ? 1 : 0);
  return true;
}

// ---- cluster helpers (TWO = true: a CTA pair works on one 256-row tile with cta_group::2 MMAs) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <bool TWO>
__device__ __forceinline__ void tma_load_3d_to(const CUtensorMap* map, uint32_t bar_addr, void* dst, int c0, int c1, int c2) {
  if constexpr (TWO) {
    // cta_group::2: the data lands in THIS CTA's shared memory, the bytes are counted on the leader CTA's barrier
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  }
}
template <bool TWO>
__device__ __forceinline__ void mma_bf16_g(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TWO) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    mma_bf16(tmem_d, desc_a, desc_b, idesc, accumulate);
  }
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far are done
template <bool TWO>
__device__ __forceinline__ void mma_commit_g(uint64_t* bar) {
  if constexpr (TWO) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
  } else {
    mma_commit(bar);
  }
}

template <int N_TILE, bool TWO>
__global__ void __launch_bounds__(kThreads, 1)
planes_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_d, const Params p) {
  // accumulator columns handled by the two epilogue warps that share a TMEM lane quarter
  constexpr int kHalf0 = ((N_TILE + 31) / 32) * 16;
  constexpr int kHalf1 = N_TILE - kHalf0;
  constexpr int NC = kHalf0;
  constexpr int kTileM = TWO ? 2 * kBlockM : kBlockM;      // rows of D per scheduling unit (CTA or CTA pair)
  constexpr int kBRows = TWO ? N_TILE / 2 : N_TILE;        // rows / columns of the B tile staged by THIS CTA
  static_assert(N_TILE % 16 == 0 && N_TILE >= 16 && N_TILE <= 256, "N tile");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = TWO ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int unit0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int units = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;   // staged by ONE CTA

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);     // the leader's arrive.expect_tx (the TMA bytes of both CTAs are counted on it)
      mbar_init(&empty_bar[s], 1);    // one tcgen05.commit (multicast to both CTAs of a pair)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], TWO ? 16 : 8);   // one arrive per epilogue warp of every CTA that drains this buffer
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Register budget: the producer / MMA / allocator warps need a handful of registers, the drain warps hold up to 128
  // accumulators + 64 staging registers per thread -> hand the first warpgroup's registers to the other two.
  // (each setmaxnreg sits at the top of its role's branch so that ptxas allocates that region against the new limit)
  if (warp >= kEpiWarp0) goto epilogue_role;
  asm volatile("setmaxnreg.dec.sync.aligned.u32 48;" ::: "memory");
  if (warp == 0) {
    // ================================ TMA producer (every CTA stages its own rows) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      Tile t;
      for (int idx = unit0; get_tile(p, idx, t); idx += units) {
        const int m0 = t.m_blk * kTileM + (int)cta_rank * kBlockM;
        const int n0 = t.n_blk * N_TILE + (int)cta_rank * (TWO ? kBRows : 0);
        // Optional (RSB_GEMM_PREFETCH = n, default off): request the A boxes into L2 n K groups ahead of the ring's own
        // loads.  Measured on 65536 x 400 x 624: 161 -> 151 TFLOP/s - the MMA thread waits on the full barrier only 11 %
        // of the time, so DRAM latency is not what paces this kernel (operand fetch of the SS-mode MMA).
        auto prefetch_a = [&](int kb) {
          const int k0 = kb * kBlockK;
          if (!p.a.mn_major) {
            tma_prefetch_3d(&map_a, k0 + t.batch * p.a.batch_col_step, m0 + t.batch * p.a.batch_row_step, 0);
          } else {
            for (int j = 0; j < kBlockM / 64; ++j)
              tma_prefetch_3d(&map_a, m0 + j * 64 + t.batch * p.a.batch_col_step, k0 + t.batch * p.a.batch_row_step, 0);
          }
        };
        if (p.prefetch > 0) {
          for (int kb = t.kb0; kb < t.kb1 && kb < t.kb0 + p.prefetch; ++kb) prefetch_a(kb);
        }
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          if (p.prefetch > 0 && kb + p.prefetch < t.kb1) prefetch_a(kb + p.prefetch);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + p.a_stage_bytes;
          if (leader) mbar_expect_tx(&full_bar[stage], TWO ? 2 * p.tx_bytes : p.tx_bytes);
          const uint32_t bar = TWO ? map_to_rank(smem_u32(&full_bar[stage]), 0) : smem_u32(&full_bar[stage]);
          const int k0 = kb * kBlockK;
          if (!p.a.mn_major) {
            tma_load_3d_to<TWO>(&map_a, bar, sa, k0 + t.batch * p.a.batch_col_step, m0 + t.batch * p.a.batch_row_step, 0);
          } else {
#pragma unroll 1
            for (int j = 0; j < kBlockM / 64; ++j)
              tma_load_3d_to<TWO>(&map_a, bar, sa + j * (p.np * kBlockK * 128), m0 + j * 64 + t.batch * p.a.batch_col_step,
                                  k0 + t.batch * p.a.batch_row_step, 0);
          }
          if (!p.b.mn_major) {
            tma_load_3d_to<TWO>(&map_b, bar, sb, k0 + t.batch * p.b.batch_col_step, n0 + t.batch * p.b.batch_row_step, 0);
          } else {
#pragma unroll 1
            for (int j = 0; j < (kBRows + 63) / 64; ++j)
              tma_load_3d_to<TWO>(&map_b, bar, sb + j * (p.np * kBlockK * 128), n0 + j * 64 + t.batch * p.b.batch_col_step,
                                  k0 + t.batch * p.b.batch_row_step, 0);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (one lane of the leader CTA) ================================
    if (lane == 0 && leader) {
      // kind::f16 instruction descriptor: D fp32 (bit 4), A / B format at bits 7 / 10 (1 = bf16, 0 = fp16), majorness,
      // N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t ab_fmt = p.np == 3 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((uint32_t)p.a.mn_major << 15) |
                             ((uint32_t)p.b.mn_major << 16) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      // plane strides / descriptor geometry inside a stage
      const uint32_t a_plane = p.a.mn_major ? kBlockK * 128 : kBlockM * 64;
      const uint32_t b_plane = p.b.mn_major ? kBlockK * 128 : kBRows * 64;
      const uint32_t a_kstep = p.a.mn_major ? 2048 : 32;   // 16 k: two 8-row groups of 1 KiB / 32 bytes in a row
      const uint32_t b_kstep = p.b.mn_major ? 2048 : 32;
      int stage = 0, buf = 0;
      uint32_t phase = 0, tphase[2] = {0, 0};
      long long dbg_tempty = 0, dbg_full = 0, dbg_t0 = p.dbg ? clock64() : 0;
      Tile t;
      for (int idx = unit0; get_tile(p, idx, t); idx += units) {
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          const bool group_start = ((kb - t.kb0) % p.drain) == 0;
          const bool group_end = ((kb - t.kb0) % p.drain) == p.drain - 1 || kb == t.kb1 - 1;
          long long c0 = 0, c1 = 0, c2 = 0;
          if (p.dbg) c0 = clock64();
          if (group_start) mbar_wait(&tempty_bar[buf], tphase[buf] ^ 1);
          if (p.dbg) c1 = clock64();
          mbar_wait(&full_bar[stage], phase);
          if (p.dbg) {
            c2 = clock64();
            dbg_tempty += c1 - c0;
            dbg_full += c2 - c1;
          }
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + p.a_stage_bytes;
          const uint32_t d_tmem = tmem_base + buf * kTmemBufStride;
          uint64_t da[3], db[3];
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            da[pl] = p.a.mn_major ? make_desc(sa + pl * a_plane, p.np * kBlockK * 128, 1024, kSwizzle128)
                                  : make_desc(sa + pl * a_plane, 16, 512, kSwizzle64);
            db[pl] = p.b.mn_major ? make_desc(sb + pl * b_plane, p.np * kBlockK * 128, 1024, kSwizzle128)
                                  : make_desc(sb + pl * b_plane, 16, 512, kSwizzle64);
          }
          // smallest band first; the first MMA of the group overwrites the (drained) accumulator
          uint32_t acc = group_start ? 0u : 1u;
#define PG_MMA(PA, PB)                                                                                                \
  _Pragma("unroll") for (int ks = 0; ks < kBlockK / 16; ++ks) {                                                       \
    mma_bf16_g<TWO>(d_tmem, da[PA] + (uint64_t)((ks * a_kstep) >> 4), db[PB] + (uint64_t)((ks * b_kstep) >> 4), idesc, acc); \
    acc = 1;                                                                                                          \
  }
          if (p.np == 3) {
            PG_MMA(0, 2) PG_MMA(1, 1) PG_MMA(2, 0)   // band 3
            PG_MMA(0, 1) PG_MMA(1, 0)                // band 2
            PG_MMA(0, 0)                             // band 1
          } else {
            PG_MMA(0, 1) PG_MMA(1, 0)                // fp16 planes: h0 h1', h1 h0' (2^-11), then h0 h0'
            PG_MMA(0, 0)
          }
#undef PG_MMA
          mma_commit_g<TWO>(&empty_bar[stage]);    // smem stage reusable once these MMAs have read it
          if (group_end) {
            mma_commit_g<TWO>(&tfull_bar[buf]);    // accumulator of this drain group complete
            tphase[buf] ^= 1;
            buf ^= 1;
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 4 + 0] = dbg_tempty;
        p.dbg[blockIdx.x * 4 + 1] = dbg_full;
        p.dbg[blockIdx.x * 4 + 2] = clock64() - dbg_t0;
      }
    }
  }
  goto teardown;
epilogue_role:
  {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;" ::: "memory");
    // ================================ drain + epilogue (every CTA: its own 128 rows) ================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - kEpiWarp0) >> 2;     // which column half of the tile
    const int col0 = half ? kHalf0 : 0;
    const int ncols = half ? kHalf1 : kHalf0;     // multiple of 16 (0 possible only if N_TILE == 16)
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t tempty_addr[2];
#pragma unroll
    for (int b = 0; b < 2; ++b)
      tempty_addr[b] = TWO ? map_to_rank(smem_u32(&tempty_bar[b]), 0) : smem_u32(&tempty_bar[b]);
    int buf = 0;
    uint32_t tphase[2] = {0, 0};
    uint32_t box_seq = 0;          // TMA-store boxes issued by this warp so far (selects the staging buffer)
    // FP16X2 operands hold A * sa and B * sb (powers of two): undo both with the caller's alpha
    const float alpha = p.alpha / (p.fmt_a.scale() * p.fmt_b.scale());
    Tile t;
    for (int idx = unit0; get_tile(p, idx, t); idx += units) {
      float acc[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[i] = 0.f;
      for (int kb = t.kb0; kb < t.kb1; kb += p.drain) {
        long long w0 = 0;
        if (p.dbg && warp == kEpiWarp0 && lane == 0) w0 = clock64();
        mbar_wait(&tfull_bar[buf], tphase[buf]);
        if (p.dbg && warp == kEpiWarp0 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&p.dbg[blockIdx.x * 4 + 3]), (unsigned long long)(clock64() - w0));
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + buf * kTmemBufStride + col0;
        // TMEM -> registers in chunks of up to 64 columns: every load of a chunk is issued before the one wait
        // (a load -> wait -> add chain per 16 columns made this loop, not the MMAs, the pace of the kernel)
#pragma unroll
        for (int c = 0; c < NC; c += 64) {
          if (c < ncols) {
            float v[64];
#pragma unroll
            for (int j = 0; j < 64; j += 16) {
              if (c + j < NC && c + j < ncols) tmem_ld16(taddr + c + j, v + j);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; j += 16) {
              if (c + j < NC && c + j < ncols) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[c + j + i] += v[j + i];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (TWO) mbar_arrive_cluster(tempty_addr[buf]); else mbar_arrive(&tempty_bar[buf]);
        }
        tphase[buf] ^= 1;
        buf ^= 1;
      }
      // ---- epilogue: this thread owns row (m0 + 32 q + lane), columns [n0 + col0, n0 + col0 + ncols) ----
      const long long row = (long long)t.m_blk * kTileM + (long long)cta_rank * kBlockM + q * 32 + lane;
      const int n0 = t.n_blk * N_TILE + col0;
      if (p.tma_store) {
        // fp32 result through shared memory: a thread owns one ROW of the accumulator, so direct stores write 16-byte
        // pieces of 32 different rows per instruction (half-used sectors, ~58 clk per store instruction: 6 us per tile
        // that the next tile's MMAs cannot hide).  Instead each warp stages [32 rows x 16 columns] boxes (64-byte
        // swizzle: conflict-free 128-bit shared stores) and one lane hands them to the TMA unit, which writes full lines
        // asynchronously and clips the rows / columns beyond the matrix.
        uint8_t* stg = smem + p.epi_offset + (warp - kEpiWarp0) * 4096;
        const int row0 = t.m_blk * kTileM + (int)cta_rank * kBlockM + q * 32;
        const unsigned char* mrow = (p.epi_mode == RSB_EPI_MASK_F32 && row < p.M) ? p.mask + row * (long long)p.N : nullptr;
        const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
        float omax = 0.f;
#pragma unroll
        for (int c = 0; c < NC; c += 16) {
          if (c < ncols && n0 + c < p.N) {
            uint8_t* buf = stg + (box_seq & 1) * 2048;
            ++box_seq;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store two boxes back has read its buffer
            __syncwarp();
            // the 16 mask bytes of this box in one 128-bit load when the row pitch allows it
            uint4 m16 = make_uint4(0u, 0u, 0u, 0u);
            const bool wide_mask = mrow != nullptr && (p.N & 15) == 0;
            if (wide_mask) m16 = *reinterpret_cast<const uint4*>(mrow + n0 + c);
#pragma unroll
            for (int h = 0; h < 16; h += 4) {
              float4 o = make_float4(acc[c + h] * alpha, acc[c + h + 1] * alpha, acc[c + h + 2] * alpha,
                                     acc[c + h + 3] * alpha);
              if (n0 + c + h < p.N) {
                if (p.bias) {
                  const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + h));
                  o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                }
                if (p.epi_mode == RSB_EPI_MASK_F32) {
                  uchar4 m = make_uchar4(0, 0, 0, 0);
                  if (wide_mask) {
                    const uint32_t w = h == 0 ? m16.x : (h == 4 ? m16.y : (h == 8 ? m16.z : m16.w));
                    m = make_uchar4(w & 0xffu, (w >> 8) & 0xffu, (w >> 16) & 0xffu, (w >> 24) & 0xffu);
                  } else if (mrow) {
                    m = *reinterpret_cast<const uchar4*>(mrow + n0 + c + h);
                  }
                  o.x = m.x ? o.x * p.drop_scale : 0.f; o.y = m.y ? o.y * p.drop_scale : 0.f;
                  o.z = m.z ? o.z * p.drop_scale : 0.f; o.w = m.w ? o.w * p.drop_scale : 0.f;
                }
                omax = fmaxf(fmaxf(omax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
              }
              *reinterpret_cast<float4*>(buf + lane * 64 + ((((uint32_t)h >> 2) ^ sw) << 4)) = o;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_d, buf, n0 + c, row0);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (p.bn_part != nullptr) {
              // BatchNorm statistics of the rows this warp just produced (the batch statistics pass over D disappears):
              // shifted sums  S1 = sum (d - k),  S2 = sum (d - k)^2  over the warp's valid rows, k = the group's first
              // row, per column - read back from the box just staged in shared memory (lane l sums column l & 15 over
              // rows 16 (l >> 4) .. +15: 17 shared loads and 2 shuffles per box; reducing the accumulator registers
              // over the rows with shuffles took 48 per box and showed as +0.02 ms per GEMM).
              const int cl = lane & 15, rh = (lane >> 4) * 16;
              const float kshift = *reinterpret_cast<const float*>(buf + ((cl >> 2) << 4) + (cl & 3) * 4);   // row 0: no swizzle
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int rr = 0; rr < 16; ++rr) {
                const int r = rh + rr;
                const float zv = *reinterpret_cast<const float*>(buf + r * 64 + ((((uint32_t)cl >> 2) ^ (((uint32_t)r >> 1) & 3u)) << 4) +
                                                                 (cl & 3) * 4);
                const float dv = (row0 + r < p.M) ? zv - kshift : 0.f;
                s1 += dv;
                s2 = fmaf(dv, dv, s2);
              }
              s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
              s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
              const int col = n0 + c + cl;
              const long long rg = (long long)t.m_blk * (kTileM / 32) + q;
              float* part = p.bn_part + rg * 3 * (long long)p.N;
              if (lane < 16 && col < p.N) {
                part[col] = s1;
                part[p.N + col] = s2;
                part[2 * (long long)p.N + col] = kshift;
              }
            }
          }
        }
        if (p.d_amax) {                      // one atomic per warp and tile (rows past M hold alpha * 0 + bias: harmless)
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, off));
          if (lane == 0 && omax > 0.f && omax < 3.0e38f) rsb::atomic_max_nonneg(p.d_amax, omax);
        }
      } else if (row < p.M) {
        if (p.partial != nullptr) {
          float* dst = p.partial + (((long long)t.split * p.batch + t.batch) * p.M + row) * p.N;
#pragma unroll
          for (int c = 0; c < NC; c += 4) {
            if (c < ncols && n0 + c < p.N)
              *reinterpret_cast<float4*>(dst + n0 + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
          }
        } else if (p.epi_mode == RSB_EPI_LINEAR || p.epi_mode == RSB_EPI_MASK_F32) {
          float* dst = p.D + (long long)t.batch * p.d_batch_stride + row * p.ldd;
          const float* csrc = p.C ? p.C + (long long)t.batch * p.d_batch_stride + row * p.ldd : nullptr;
          const unsigned char* mrow = p.epi_mode == RSB_EPI_MASK_F32 ? p.mask + row * (long long)p.N : nullptr;
          float omax = 0.f;
#pragma unroll
          for (int c = 0; c < NC; c += 4) {
            if (c < ncols && n0 + c < p.N) {
              float4 o = make_float4(acc[c] * alpha, acc[c + 1] * alpha, acc[c + 2] * alpha, acc[c + 3] * alpha);
              if (csrc) {
                const float4 cc = *reinterpret_cast<const float4*>(csrc + n0 + c);
                o.x = fmaf(p.beta, cc.x, o.x); o.y = fmaf(p.beta, cc.y, o.y);
                o.z = fmaf(p.beta, cc.z, o.z); o.w = fmaf(p.beta, cc.w, o.w);
              }
              if (p.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c));
                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
              }
              if (mrow) {   // gradient through dropout(relu(.)): keep-and-positive mask of the forward pass
                const uchar4 m = *reinterpret_cast<const uchar4*>(mrow + n0 + c);
                o.x = m.x ? o.x * p.drop_scale : 0.f; o.y = m.y ? o.y * p.drop_scale : 0.f;
                o.z = m.z ? o.z * p.drop_scale : 0.f; o.w = m.w ? o.w * p.drop_scale : 0.f;
              }
              omax = fmaxf(fmaxf(omax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
              *reinterpret_cast<float4*>(dst + n0 + c) = o;
            }
          }
          if (p.d_amax && omax > 0.f && omax < 3.0e38f) rsb::atomic_max_nonneg(p.d_amax, omax);   // (rows diverge here: per thread)
        } else {
          // RSB_EPI_RELU_DROPOUT_PLANES: y = dropout(relu(acc + bias)) -> planes; `mask` holds the dropout keep bits on
          //                              entry (rsb_dropout_keep_mask: the Philox draw is NOT part of this exposed
          //                              epilogue) and keep && (pre-activation > 0) on exit   (forward of a hidden layer)
          // RSB_EPI_MASK_PLANES        : g = acc * mask / (1 - p)      -> planes               (dX of a hidden layer)
          const bool fwd = p.epi_mode == RSB_EPI_RELU_DROPOUT_PLANES;
          __nv_bfloat16* prow = p.out_planes + row * p.out_ld;
          unsigned char* mrow = p.mask + row * (long long)p.N;
#pragma unroll
          for (int c = 0; c < NC; c += 8) {
            if (c < ncols && n0 + c < p.N) {
              float v[8];
#pragma unroll
              for (int h = 0; h < 8; h += 4) {
                const bool valid = n0 + c + h < p.N;      // N is a multiple of 4, not necessarily of 8
                float4 o = make_float4(acc[c + h] * alpha, acc[c + h + 1] * alpha, acc[c + h + 2] * alpha,
                                       acc[c + h + 3] * alpha);
                uchar4 m = make_uchar4(0, 0, 0, 0);
                if (valid) {
                  m = *reinterpret_cast<const uchar4*>(mrow + n0 + c + h);
                  if (fwd) {
                    if (p.bias) {
                      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + h));
                      o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    m.x = m.x && (o.x > 0.f); m.y = m.y && (o.y > 0.f); m.z = m.z && (o.z > 0.f); m.w = m.w && (o.w > 0.f);
                    *reinterpret_cast<uchar4*>(mrow + n0 + c + h) = m;
                  }
                }
                v[h] = m.x ? o.x * p.drop_scale : 0.f; v[h + 1] = m.y ? o.y * p.drop_scale : 0.f;
                v[h + 2] = m.z ? o.z * p.drop_scale : 0.f; v[h + 3] = m.w ? o.w * p.drop_scale : 0.f;
              }
              store_planes8(prow + n0 + c, p.out_plane_stride, v);
            }
          }
          // "ones" column right after the data: a weight-gradient GEMM over these planes yields the bias gradient
          if (p.ones_col && t.n_blk == p.n_tiles - 1 && half == 0) {
            float v[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_planes8(prow + ((p.N + 7) & ~7), p.out_plane_stride, v);
          }
        }
      }
    }
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged boxes fully written
  }
teardown:
  // ---- teardown ----
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (TWO)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// split-K: D = alpha * sum_s partial[s] + beta * C + bias, fixed summation order
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long batch, long long M, long long N,
                                     float* __restrict__ D, long long ldd, long long d_batch_stride,
                                     const float* __restrict__ C, const float* __restrict__ bias, float alpha, float beta,
                                     rsb::PlaneFmt fmt_a, rsb::PlaneFmt fmt_b) {
  alpha /= fmt_a.scale() * fmt_b.scale();
  const long long n4 = N / 4, total = batch * M * n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c4 = i % n4, r = (i / n4) % M, l = i / (n4 * M);
    const long long src = (l * M + r) * N + c4 * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < splits; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (long long)k * batch * M * N + src));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= alpha; s.y *= alpha; s.z *= alpha; s.w *= alpha;
    const long long o = l * d_batch_stride + r * ldd + c4 * 4;
    if (C) {
      const float4 cc = *reinterpret_cast<const float4*>(C + o);
      s.x = fmaf(beta, cc.x, s.x); s.y = fmaf(beta, cc.y, s.y); s.z = fmaf(beta, cc.z, s.z); s.w = fmaf(beta, cc.w, s.w);
    }
    if (bias) {
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c4 * 4));
      s.x += bb.x; s.y += bb.y; s.z += bb.z; s.w += bb.w;
    }
    *reinterpret_cast<float4*>(D + o) = s;
  }
}

// fp32 [rows, cols] (ld) -> three bf16 planes [3][rows][out_ld]; columns >= cols (up to out_ld) are zero-filled.
// transpose = 1: out plane [cols][out_ld >= rows] = in^T (small matrices: nn.Linear weights for the dX GEMM).
__global__ void split_planes_kernel(const float* __restrict__ in, long long rows, long long cols, long long ld,
                                    uint16_t* __restrict__ out, long long out_ld, long long plane_stride, int transpose,
                                    int ones_col, rsb::PlaneFmt fmt) {
  const float ps = fmt.scale();
  const long long orows = transpose ? cols : rows, ocols8 = out_ld / 8;
  const long long ones_at = ones_col ? (((transpose ? rows : cols) + 7) & ~7ll) : -1;   // first column after the padded data
  const long long total = orows * ocols8;
  const bool small = total < (1ll << 32);
  const rsb::FastDiv row_div = rsb::make_fastdiv((unsigned long long)ocols8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = small ? (long long)rsb::fastdiv((unsigned)i, row_div) : i / ocols8, c0 = (i - r * ocols8) * 8;
    float x[8];
    const long long icols = transpose ? rows : cols;
    if (!transpose && c0 + 8 <= icols && (ld & 3) == 0) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(in + r * ld + c0));
      const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ld + c0 + 4));
      x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long c = c0 + j;
        x[j] = (c < icols) ? (transpose ? __ldg(in + c * ld + r) : __ldg(in + r * ld + c)) : 0.f;
      }
      if (c0 == ones_at) x[0] = 1.f;
    }
    rsb::store_planes<8>(out + r * out_ld + c0, plane_stride, x, fmt.format, ps);
  }
}

// *amax = max(*amax, max |in|) over a [rows, cols] (ld) fp32 matrix: the bound of an FP16X2 split (common.cuh).
// Contiguous 16-byte-aligned matrices are read as one flat float4 stream, four loads in flight per thread.
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ in, long long rows, long long cols, long long ld,
                                                     float* __restrict__ amax) {
  float m = 0.f;
  const bool al = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (al && ld == cols && ((rows * cols) & 3) == 0) {
    const float4* q = reinterpret_cast<const float4*>(in);
    const long long total = rows * cols / 4;
    for (long long i = tid; i < total; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < total) ? __ldg(q + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v[u].x), fabsf(v[u].y))), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
    }
  } else if (al && (cols & 3) == 0 && (ld & 3) == 0) {
    const long long c4n = cols / 4, total = rows * c4n;
    for (long long i = tid; i < total; i += stride) {
      const long long r = i / c4n, c4 = i - r * c4n;
      const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * ld) + c4);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
  } else {
    const long long total = rows * cols;
    for (long long i = tid; i < total; i += stride) {
      const long long r = i / cols, c = i - r * cols;
      m = fmaxf(m, fabsf(__ldg(in + r * ld + c)));
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
    if (m > 0.f && m < 3.0e38f) rsb::atomic_max_nonneg(amax, m);
  }
}

// dropout keep bits of a [.., 4 n4] activation: byte i*4+j = (Philox(offset + i).j >= thr); the same stream position /
// counter mapping as rsb_relu_dropout_fwd, so fused and un-fused paths draw identical masks
__global__ void dropout_keep_mask_kernel(uchar4* __restrict__ mask, long long n4, unsigned thr, unsigned long long seed,
                                         unsigned long long offset, const unsigned long long* __restrict__ offset_dev) {
  rsb::Philox rng{(unsigned)seed, (unsigned)(seed >> 32)};
  if (offset_dev) offset += *offset_dev;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 r = rng(offset + (unsigned long long)i);
    mask[i] = make_uchar4(r.x >= thr, r.y >= thr, r.z >= thr, r.w >= thr);
  }
}

// y = dropout_p(relu(x)) of an fp32 [M, N] activation written as planes (+ ones column) and the keep-and-positive mask:
// rsb_relu_dropout_fwd and rsb_split_planes in one HBM pass (read 4 B, write 6 + 1 B per element), same Philox stream
__global__ void relu_dropout_planes_kernel(const float* __restrict__ x, long long M, int N, long long ldx, float scale,
                                           unsigned thr, unsigned long long seed, unsigned long long offset,
                                           const unsigned long long* __restrict__ offset_dev,
                                           __nv_bfloat16* __restrict__ out, long long out_ld, long long plane_stride,
                                           unsigned char* __restrict__ mask, int ones_col) {
  rsb::Philox rng{(unsigned)seed, (unsigned)(seed >> 32)};
  if (offset_dev) offset += *offset_dev;
  const int n8 = (N + 7) / 8 + (ones_col ? 1 : 0);
  const int n4 = N / 4;
  const long long total = M * n8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n8;
    const int c0 = (int)(i - r * n8) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 >= ((N + 7) & ~7)) {
      v[0] = 1.f;                                     // the ones column
    } else {
#pragma unroll
      for (int h = 0; h < 8; h += 4) {
        if (c0 + h < N) {
          const float4 z = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c0 + h));
          const uint4 q = rng(offset + (unsigned long long)(r * n4 + ((c0 + h) >> 2)));
          uchar4 m;
          m.x = (z.x > 0.f) && (q.x >= thr); m.y = (z.y > 0.f) && (q.y >= thr);
          m.z = (z.z > 0.f) && (q.z >= thr); m.w = (z.w > 0.f) && (q.w >= thr);
          *reinterpret_cast<uchar4*>(mask + r * (long long)N + c0 + h) = m;
          v[h] = m.x ? z.x * scale : 0.f; v[h + 1] = m.y ? z.y * scale : 0.f;
          v[h + 2] = m.z ? z.z * scale : 0.f; v[h + 3] = m.w ? z.w * scale : 0.f;
        }
      }
    }
    store_planes8(out + r * out_ld + c0, plane_stride, v);
  }
}

// gz[r, c] = g[r] * w[c] * mask[r, c] / (1 - p) as planes: backward of dropout(relu(.)) for the rank-1 upstream gradient
// of the MLP's one-output Linear (src/models/deepfm.py:64), written straight in the operand format of the dX / dW GEMMs
__global__ void rank1_mask_planes_kernel(const float* __restrict__ g_row, const float* __restrict__ w_col,
                                         const unsigned char* __restrict__ mask, long long M, int N, float scale,
                                         __nv_bfloat16* __restrict__ out, long long out_ld, long long plane_stride) {
  const int n8 = N / 8;
  const long long total = M * n8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n8;
    const int c0 = (int)(i - r * n8) * 8;
    const float g = __ldg(g_row + r) * scale;
    const uint2 m8 = __ldg(reinterpret_cast<const uint2*>(mask + r * (long long)N + c0));
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w_col + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w_col + c0 + 4));
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned byte = ((j < 4 ? m8.x : m8.y) >> (8 * (j & 3))) & 0xffu;
      v[j] = byte ? g * w[j] : 0.f;
    }
    store_planes8(out + r * out_ld + c0, plane_stride, v);
  }
}

// ------------------------------------------------------------------------------------------------ host ---
// cuTensorMapEncodeTiled is a DRIVER entry point: resolved at first use through the runtime, so that librsb.so has no
// link-time dependency on libcuda (the library must load, and its argument checks run, on a box without a driver)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static bool encode_map(CUtensorMap* map, const void* base, long long cols, long long rows, long long ld, long long plane_stride,
                       int box_cols, int box_rows, CUtensorMapSwizzle swz, int np) {
  // dims (innermost first): columns, rows, planes
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)np};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)np};
  cuuint32_t estr[3] = {1, 1, 1};
  EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr) return false;
  CUresult rc = encode(map, np == 3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base),
                       dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS;
}

template <int N_TILE, bool TWO>
static cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const Params& p, int grid, size_t smem,
                          cudaStream_t st) {
  static bool attr[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr[dev]) {
    cudaError_t e = cudaFuncSetAttribute(planes_gemm_kernel<N_TILE, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kSmemLimit - 2048));   // static: barriers
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = TWO ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, planes_gemm_kernel<N_TILE, TWO>, ma, mb, md, p);
}

}  // namespace pg

using namespace pg;

static int pick_n_tile(long long N, int* n_tiles) {
  int nt = (int)((N + 255) / 256);
  long long per = (N + nt - 1) / nt;
  int tile = (int)((per + 15) / 16 * 16);
  *n_tiles = (int)((N + tile - 1) / tile);
  return tile;
}

// CTA pairs (cta_group::2, 256-row tiles, each CTA stages half of B) whenever there is more than one 128-row block
static bool use_pairs(long long M) {
  static const int mode = [] { const char* v = getenv("RSB_GEMM_PAIRS"); return v ? atoi(v) : 0; }();
  return mode != 0 && M > kBlockM;
}

static int default_splits(long long M, long long N, long long K, long long batch) {
  int n_tiles;
  const int tile = pick_n_tile(N, &n_tiles);
  (void)tile;
  const bool two = use_pairs(M);
  const int tile_m = two ? 2 * kBlockM : kBlockM;
  const long long tiles = ((M + tile_m - 1) / tile_m) * n_tiles * batch;
  const long long kb = (K + kBlockK - 1) / kBlockK;
  const int units = two ? rsb::sm_count() / 2 : rsb::sm_count();
  if (tiles >= units || kb < 16) return 1;
  long long s = units / tiles;               // fill one wave
  if (s > kb / 8) s = kb / 8;                // at least 8 K groups (256 k) per split
  return s < 1 ? 1 : (int)s;
}

extern "C" RSB_API int64_t rsb_gemm_planes_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch, int32_t split_k) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return -1;
  int s = split_k > 0 ? split_k : default_splits(M, N, K, batch);
  return s > 1 ? (int64_t)s * batch * M * N * 4 + 256 : 256;
}

// *amax_out = max(*amax_out, mul * max |g| * max |w|): bound on the rank-1 head gradient g[r] * w[c] * mask / (1 - p)
__global__ void __launch_bounds__(1024) rank1_bound_kernel(const float* __restrict__ g, long long m, const float* __restrict__ w,
                                                           long long n, float mul, float* __restrict__ amax) {
  // every CTA: its slice of g, all of w (short); max over the CTAs of mg_i * mw * mul is the bound
  float mg = 0.f, mw = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
    mg = fmaxf(mg, fabsf(__ldg(g + i)));
  for (long long i = threadIdx.x; i < n; i += blockDim.x) mw = fmaxf(mw, fabsf(__ldg(w + i)));
  __shared__ float sg[32], sw[32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, off));
    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, off));
  }
  if ((threadIdx.x & 31) == 0) { sg[threadIdx.x >> 5] = mg; sw[threadIdx.x >> 5] = mw; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { mg = fmaxf(mg, sg[i]); mw = fmaxf(mw, sw[i]); }
    const float b = mg * mw * mul;
    if (b > 0.f && b < 3.0e38f) rsb::atomic_max_nonneg(amax, b);
  }
}

// Bytes of the [row_groups][3][N] statistics buffer a GEMM with rsb_gemm_epilogue.bn_partials writes, or -1 when this
// shape / format cannot carry the statistics (the fp32 TMA-store epilogue must fit beside >= 3 pipeline stages; K-major
// operands; the driver's tensor-map encoder must be there).  Mirrors the stage arithmetic of rsb_gemm_planes.
extern "C" RSB_API int64_t rsb_gemm_bn_partials_bytes(int64_t M, int64_t N, int64_t K, int32_t format) {
  if (M <= 0 || N <= 0 || K <= 0 || N % 4 || use_pairs(M)) return -1;
  static const int want = [] { const char* v = getenv("RSB_GEMM_TMA_STORE"); return v ? atoi(v) : 1; }();
  if (!want || encode_tiled_fn() == nullptr) return -1;
  int n_tiles;
  const int n_tile = pick_n_tile(N, &n_tiles);
  const uint32_t np = format == RSB_PLANES_FP16X2 ? 2u : 3u;
  const uint32_t a_bytes = np * kBlockM * kBlockK * 2;
  const uint32_t b_bytes = ((np * (uint32_t)n_tile * kBlockK * 2) + 1023u) & ~1023u;
  if ((size_t)3 * (a_bytes + b_bytes) + 8u * 4096u + 1024 > kSmemLimit - 2048) return -1;
  const int64_t row_groups = (M + kBlockM - 1) / kBlockM * (kBlockM / 32);
  return row_groups * 3 * N * 4;
}

extern "C" RSB_API int rsb_rank1_absmax(const float* g_row, int64_t M, const float* w_col, int64_t N, float mul, float* amax_out,
                                        void* stream) {
  if (!g_row || !w_col || !amax_out || M < 0 || N < 0) return RSB_ERR_BAD_ARG;
  if (M == 0 || N == 0) return RSB_OK;
  long long blocks = (M + 4095) / 4096;
  if (blocks > 64) blocks = 64;
  rank1_bound_kernel<<<(unsigned)blocks, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(g_row, M, w_col, N, mul, amax_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_absmax(const float* in, int64_t rows, int64_t cols, int64_t ld, float* amax_out, void* stream) {
  if (!in || !amax_out || rows < 0 || cols < 0 || ld < cols) return RSB_ERR_BAD_ARG;
  if (rows == 0 || cols == 0) return RSB_OK;
  long long blocks = (rows * cols / 16 + 255) / 256;
  const long long cap = (long long)rsb::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  absmax_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, rows, cols, ld, amax_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_split_planes(const float* in, int64_t rows, int64_t cols, int64_t ld, int32_t transpose, int32_t ones_col,
                                        void* out_planes, int64_t out_ld, int64_t plane_stride, const rsb_planes_format* fmt,
                                        void* stream) {
  if (!in || !out_planes || rows <= 0 || cols <= 0 || ld < (transpose ? cols : cols) || !rsb::plane_fmt_ok(fmt)) return RSB_ERR_BAD_ARG;
  const int64_t ocols = transpose ? rows : cols, orows = transpose ? cols : rows;
  if (out_ld % 8 || out_ld < ocols + (ones_col ? 8 : 0) || plane_stride < orows * out_ld || plane_stride % 8 ||
      (reinterpret_cast<uintptr_t>(out_planes) & 15))
    return RSB_ERR_UNSUPPORTED;
  const long long total = orows * (out_ld / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)rsb::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  split_planes_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, rows, cols, ld, reinterpret_cast<uint16_t*>(out_planes), out_ld, plane_stride, transpose, ones_col, rsb::plane_fmt(fmt));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_relu_dropout_planes(const float* x, int64_t M, int32_t N, int64_t ldx, float p, uint64_t seed, uint64_t offset,
                                               const uint64_t* offset_dev, int32_t ones_col, void* out_planes, int64_t out_ld,
                                               int64_t plane_stride, uint8_t* mask, void* stream) {
  if (!x || !out_planes || !mask || M < 0 || N <= 0 || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  if (M == 0) return RSB_OK;
  if (N % 4 || ldx % 4 || out_ld % 8 || plane_stride % 8 || out_ld < ((N + 7) / 8) * 8 + (ones_col ? 8 : 0) ||
      (reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(out_planes) & 15u) ||
      (reinterpret_cast<uintptr_t>(mask) & 3u))
    return RSB_ERR_UNSUPPORTED;
  const long long total = M * ((N + 7) / 8 + (ones_col ? 1 : 0));
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)rsb::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  relu_dropout_planes_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, M, N, ldx, 1.0f / (1.0f - p), (unsigned)(p * 4294967296.0), seed, offset,
      reinterpret_cast<const unsigned long long*>(offset_dev), reinterpret_cast<__nv_bfloat16*>(out_planes), out_ld, plane_stride,
      mask, ones_col);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_dropout_keep_mask(uint8_t* mask, int64_t numel, float p, uint64_t seed, uint64_t offset,
                                             const uint64_t* offset_dev, void* stream) {
  if (!mask || numel < 0 || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  if (numel == 0) return RSB_OK;
  if (numel % 4 || (reinterpret_cast<uintptr_t>(mask) & 3u)) return RSB_ERR_UNSUPPORTED;
  long long blocks = (numel / 4 + 255) / 256;
  const long long cap = (long long)rsb::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  dropout_keep_mask_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uchar4*>(mask), numel / 4, (unsigned)(p * 4294967296.0), seed, offset,
      reinterpret_cast<const unsigned long long*>(offset_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_rank1_mask_planes(const float* g_row, const float* w_col, const uint8_t* mask, int64_t M, int32_t N, float p,
                                             void* out_planes, int64_t out_ld, int64_t plane_stride, void* stream) {
  if (!g_row || !w_col || !mask || !out_planes || M < 0 || N <= 0 || p < 0.f || p >= 1.f) return RSB_ERR_BAD_ARG;
  if (M == 0) return RSB_OK;
  if (N % 8 || out_ld % 8 || out_ld < N || plane_stride % 8 || (reinterpret_cast<uintptr_t>(out_planes) & 15u) ||
      (reinterpret_cast<uintptr_t>(mask) & 7u) || (reinterpret_cast<uintptr_t>(w_col) & 15u))
    return RSB_ERR_UNSUPPORTED;
  const long long total = M * (N / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)rsb::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rank1_mask_planes_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g_row, w_col, mask, M, N, 1.0f / (1.0f - p), reinterpret_cast<__nv_bfloat16*>(out_planes), out_ld, plane_stride);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  return RSB_OK;
}

extern "C" RSB_API int rsb_gemm_planes(const rsb_planes_operand* A, const rsb_planes_operand* B, int64_t M, int64_t N, int64_t K,
                                       int64_t batch, int32_t split_k, const float* C, float* D, int64_t ldd,
                                       int64_t d_batch_stride, const float* bias, float alpha, float beta,
                                       const rsb_gemm_epilogue* epi, void* workspace, int64_t workspace_bytes, void* stream) {
  const int mode = epi ? epi->mode : RSB_EPI_LINEAR;
  const bool to_planes = mode == RSB_EPI_RELU_DROPOUT_PLANES || mode == RSB_EPI_MASK_PLANES;
  if (!A || !B || !A->planes || !B->planes || (!D && !to_planes) || M <= 0 || N <= 0 || K <= 0 || batch <= 0) return RSB_ERR_BAD_ARG;
  if (mode < RSB_EPI_LINEAR || mode > RSB_EPI_MASK_F32) return RSB_ERR_BAD_ARG;
  if (A->fmt.format != B->fmt.format || !rsb::plane_fmt_ok(&A->fmt) || !rsb::plane_fmt_ok(&B->fmt)) return RSB_ERR_BAD_ARG;
  if (epi && epi->d_amax && (to_planes || split_k > 1)) return RSB_ERR_UNSUPPORTED;   // max |D| is formed by the fp32 epilogue
  if (epi && epi->d_amax) split_k = 1;
  if (epi && epi->bn_partials) {
    // column statistics ride the TMA-store epilogue of one plain un-batched, un-split GEMM (see rsb_gemm_bn_partials_bytes)
    if (mode != RSB_EPI_LINEAR || batch != 1 || C != nullptr || split_k > 1 || use_pairs(M)) return RSB_ERR_UNSUPPORTED;
    split_k = 1;
  }
  if (mode != RSB_EPI_LINEAR) {
    // fused epilogues: one un-batched, un-split GEMM whose output feeds the next GEMM
    if (batch != 1 || beta != 0.f || !epi->mask || epi->p < 0.f || epi->p >= 1.f) return RSB_ERR_BAD_ARG;
    if (to_planes && (!epi->out_planes || epi->out_ld % 8 || epi->out_plane_stride % 8 ||
                      epi->out_ld < ((N + 7) / 8) * 8 + (epi->ones_col ? 8 : 0) ||
                      (reinterpret_cast<uintptr_t>(epi->out_planes) & 15u) || (reinterpret_cast<uintptr_t>(epi->mask) & 3u)))
      return RSB_ERR_UNSUPPORTED;
    split_k = 1;
  }
  if (beta != 0.f && !C) return RSB_ERR_BAD_ARG;
  if (M > 0x7fffffff || N > 0x7fffffff || K > 0x7fffffff || batch > 65535) return RSB_ERR_BAD_ARG;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (!al16(A->planes) || !al16(B->planes) || (D && !al16(D)) || (C && !al16(C)) || (bias && !al16(bias))) return RSB_ERR_UNSUPPORTED;
  if (N % 4 || ldd % 4 || d_batch_stride % 4 || A->ld % 8 || B->ld % 8 || A->plane_stride % 8 || B->plane_stride % 8)
    return RSB_ERR_UNSUPPORTED;
  // a batch that advances along K inside one stored matrix must not let a K tail read its neighbour
  const bool a_k_step = A->mn_major ? A->batch_row_step != 0 : A->batch_col_step != 0;
  const bool b_k_step = B->mn_major ? B->batch_row_step != 0 : B->batch_col_step != 0;
  if (batch > 1 && (a_k_step || b_k_step) && K % kBlockK) return RSB_ERR_UNSUPPORTED;
  {
    // cuTensorMapEncodeTiled is a driver call: it needs a context current on this (possibly autograd worker) thread
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaSetDevice(dev);
    if (e != cudaSuccess) return (int)e;
  }
  Params p = {};
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.batch = (int)batch;
  const bool two = use_pairs(M);
  const int tile_m = two ? 2 * kBlockM : kBlockM;
  p.n_tile = pick_n_tile(N, &p.n_tiles);
  p.m_tiles = (int)((M + tile_m - 1) / tile_m);
  p.k_blocks = (int)((K + kBlockK - 1) / kBlockK);
  p.splits = split_k > 0 ? split_k : default_splits(M, N, K, batch);
  if (p.splits > p.k_blocks) p.splits = p.k_blocks;
  p.a.mn_major = A->mn_major; p.a.batch_row_step = (int)A->batch_row_step; p.a.batch_col_step = (int)A->batch_col_step;
  p.b.mn_major = B->mn_major; p.b.batch_row_step = (int)B->batch_row_step; p.b.batch_col_step = (int)B->batch_col_step;
  p.np = A->fmt.format == RSB_PLANES_FP16X2 ? 2 : 3;
  p.fmt_a = rsb::plane_fmt(&A->fmt);
  p.fmt_b = rsb::plane_fmt(&B->fmt);
  const uint32_t np = (uint32_t)p.np;
  p.a_stage_bytes = np * kBlockM * kBlockK * 2;                                           // both majors: 8 KiB per plane
  const int b_rows = two ? p.n_tile / 2 : p.n_tile;                                       // staged by one CTA
  p.b_stage_bytes = B->mn_major ? (uint32_t)((b_rows + 63) / 64) * np * kBlockK * 128 : np * b_rows * kBlockK * 2;
  p.tx_bytes = p.a_stage_bytes + p.b_stage_bytes;
  p.b_stage_bytes = (p.b_stage_bytes + 1023u) & ~1023u;                                   // keep every stage 1 KiB aligned
  const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  p.stages = (int)((kSmemLimit - 4096) / stage_bytes);   // 1 KiB alignment slack + static shared memory
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 2) return RSB_ERR_UNSUPPORTED;
  {
    static const int drain = [] { const char* v = getenv("RSB_GEMM_DRAIN"); int d = v ? atoi(v) : 1; return d < 1 ? 1 : d; }();
    // FP16X2 issues 3 MMAs per 16 k instead of 6, so the TMEM reads of the drain (64 B/clk: 1664 clk per 128 x 208 tile)
    // weigh twice as much against the MMA time: two K groups (64 k) per drain.  Measured on 65536 x 400 x 624, error vs
    // fp64 / time: 1 group 2.2e-7 / 0.163 ms, 2: 3.0e-7 / 0.150, 3: 4.3e-7 / 0.137, 4: 5.5e-7 / 0.135 (cuBLAS fp32 SGEMM,
    // what the reference runs, errs 1.2e-6 on the same data); with 3 the DCN-Mix parity test sees more Adam sign flips of
    // rounding-level gradients than it allows, so 2 it is.
    static const int drain16 = [] { const char* v = getenv("RSB_GEMM_DRAIN_FP16"); int d = v ? atoi(v) : 2; return d < 1 ? 1 : d; }();
    p.drain = p.np == 2 ? drain16 : drain;
    static const int prefetch = [] { const char* v = getenv("RSB_GEMM_PREFETCH"); int d = v ? atoi(v) : 0; return d < 0 ? 0 : d; }();
    p.prefetch = prefetch;
  }
  {
    static long long* dbg_buf = [] {
      long long* q = nullptr;
      if (getenv("RSB_GEMM_DEBUG")) { cudaMalloc(&q, sizeof(long long) * 4096); }
      return q;
    }();
    p.dbg = dbg_buf;
    if (p.dbg) cudaMemsetAsync(p.dbg, 0, sizeof(long long) * 4096, reinterpret_cast<cudaStream_t>(stream));
  }
  p.epi_mode = mode;
  if (epi) {
    p.out_planes = reinterpret_cast<__nv_bfloat16*>(epi->out_planes);
    p.out_ld = epi->out_ld; p.out_plane_stride = epi->out_plane_stride; p.ones_col = epi->ones_col;
    p.mask = epi->mask;
    p.drop_scale = 1.0f / (1.0f - epi->p);
    p.d_amax = epi->d_amax;
    p.bn_part = epi->bn_partials;
  }
  p.D = D; p.ldd = ldd; p.d_batch_stride = d_batch_stride; p.C = C; p.bias = bias; p.alpha = alpha; p.beta = beta;
  if (p.splits > 1) {
    const int64_t need = (int64_t)p.splits * batch * M * N * 4 + 256;
    if (!workspace || workspace_bytes < need) return RSB_ERR_WORKSPACE;
    p.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  }
  CUtensorMap ma, mb;
  bool ok = A->mn_major ? encode_map(&ma, A->planes, A->cols, A->rows, A->ld, A->plane_stride, 64, kBlockK, CU_TENSOR_MAP_SWIZZLE_128B, p.np)
                        : encode_map(&ma, A->planes, A->cols, A->rows, A->ld, A->plane_stride, kBlockK, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B, p.np);
  ok = ok && (B->mn_major ? encode_map(&mb, B->planes, B->cols, B->rows, B->ld, B->plane_stride, 64, kBlockK, CU_TENSOR_MAP_SWIZZLE_128B, p.np)
                          : encode_map(&mb, B->planes, B->cols, B->rows, B->ld, B->plane_stride, kBlockK, b_rows, CU_TENSOR_MAP_SWIZZLE_64B, p.np));
  if (!ok) return RSB_ERR_UNSUPPORTED;
  // fp32 results of the big un-batched, un-split GEMMs leave through shared memory + TMA stores when the staging
  // buffers (8 warps x 2 x 2 KiB) fit beside the pipeline stages
  CUtensorMap md = ma;
  {
    static const int want = [] { const char* v = getenv("RSB_GEMM_TMA_STORE"); return v ? atoi(v) : 1; }();
    const uint32_t epi_bytes = 8u * 4096u;
    const bool f32_out = (mode == RSB_EPI_LINEAR || mode == RSB_EPI_MASK_F32) && p.partial == nullptr && batch == 1 && C == nullptr;
    // a deep ring must not crowd the staging buffers out: beyond 3 stages the ring only hides DRAM latency the MMA thread
    // does not wait for (8 % of its time), while a direct-store epilogue exposes ~6 us per tile
    if (want && f32_out) {
      while (p.stages > 3 && (size_t)p.stages * stage_bytes + epi_bytes + 1024 > kSmemLimit - 2048) --p.stages;
    }
    if (want && f32_out && (size_t)p.stages * stage_bytes + epi_bytes + 1024 <= kSmemLimit - 2048) {
      EncodeTiledFn encode = encode_tiled_fn();
      cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
      cuuint64_t strides[1] = {(cuuint64_t)ldd * 4};
      cuuint32_t box[2] = {16, 32};
      cuuint32_t estr[2] = {1, 1};
      if (encode != nullptr &&
          encode(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, D, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        p.tma_store = 1;
        p.epi_offset = (uint32_t)((size_t)p.stages * stage_bytes);
      }
    }
  }
  if (p.bn_part != nullptr && !p.tma_store) return RSB_ERR_UNSUPPORTED;
  const long long tiles = (long long)p.m_tiles * p.n_tiles * p.batch * p.splits;
  int grid = two ? rsb::sm_count() / 2 : rsb::sm_count();     // scheduling units: CTA pairs or CTAs
  if (tiles < grid) grid = (int)tiles;
  if (two) grid *= 2;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + (p.tma_store ? 8 * 4096 : 0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaErrorInvalidValue;
  switch (p.n_tile) {
#define PG_CASE(NT) case NT: e = two ? launch<NT, true>(ma, mb, md, p, grid, smem, st) : launch<NT, false>(ma, mb, md, p, grid, smem, st); break;
    PG_CASE(16) PG_CASE(32) PG_CASE(48) PG_CASE(64) PG_CASE(80) PG_CASE(96) PG_CASE(112) PG_CASE(128)
    PG_CASE(144) PG_CASE(160) PG_CASE(176) PG_CASE(192) PG_CASE(208) PG_CASE(224) PG_CASE(240) PG_CASE(256)
#undef PG_CASE
    default: return RSB_ERR_UNSUPPORTED;
  }
  if (e != cudaSuccess) return (int)e;
  rsb::note_launch(1);
  if (p.dbg) {
    cudaStreamSynchronize(st);
    static long long host[4096];
    cudaMemcpy(host, p.dbg, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
    double a = 0, b = 0, c = 0, d = 0;
    for (int i = 0; i < grid; ++i) { a += host[4 * i]; b += host[4 * i + 1]; c += host[4 * i + 2]; d += host[4 * i + 3]; }
    const long long groups = (long long)p.k_blocks / p.splits * ((tiles + grid - 1) / grid);
    fprintf(stderr, "[rsb gemm dbg] M=%d N=%d K=%d tiles=%lld grid=%d: MMA thread per CTA: total %.0f clk, wait tempty %.0f, wait full %.0f "
            "(~%lld K groups -> %.0f clk/group); drain warp waits tfull %.0f clk\n", p.M, p.N, p.K, tiles, grid, c / grid, a / grid, b / grid,
            groups, c / grid / (double)groups, d / grid);
  }
  if (p.splits > 1) {
    const long long total = batch * M * (N / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)rsb::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(p.partial, p.splits, batch, M, N, D, ldd, d_batch_stride, C, bias, alpha, beta,
                                                           p.fmt_a, p.fmt_b);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    rsb::note_launch(1);
  }
  return RSB_OK;
}
