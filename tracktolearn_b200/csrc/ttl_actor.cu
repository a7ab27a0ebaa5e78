// SAC actor forward for B200 (sm_100a): the only dense contraction on the tracking path.
//
// Reference: algorithms/shared/offpolicy.py:94-140 (MaxEntropyActor.forward) over the
// nn.Sequential built by algorithms/shared/utils.py:41-51 (Linear+ReLU x3, Linear), which the
// reference runs as four cuBLAS sgemm calls plus ~15 small elementwise launches.
//
// Here:
//   pack_state   fp32 state rows [n][ld] -> bf16 [n][K0 padded to 64]
//   dense layers C = relu(A . W^T + b), bf16 operands staged by TMA (128B swizzle) into a
//                4-stage shared-memory ring, tcgen05.mma (cta_group::1, M=128, N=256, K=16)
//                issued by one thread, fp32 accumulators double-buffered in TMEM (2 x 256
//                columns), epilogue warps read them back with tcgen05.ld, add bias, ReLU,
//                convert to bf16 and store; persistent CTAs walk the tile list.
//   head         last (6-wide) layer on CUDA cores in fp32 + clamp/exp/tanh/log-prob policy
//                head fused in one kernel (a 6-column GEMM has no tensor-core shape).
// A CUDA-core fp32 tier (TTL_PRECISION_FP32) reproduces the reference's fp32 arithmetic to
// ~1e-6 for parity tests and users who want it.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ttl_common.cuh"
#include "ttl_tc.cuh"

namespace {
using namespace ttl_tc;

// ==========================================================================================
// Dense layer: C[m][ldc] = act(A[m][k] . W[n][k]^T + bias)
// ==========================================================================================
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4, ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = BN * BK * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GEMM_THREADS = 256;             // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps4-7 epilogue
constexpr int TMEM_COLS = ACC_STAGES * BN;    // 512: all of tensor memory
constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 + 256;

// UMMA shared-memory descriptor: K-major operand tile, rows of 64 bf16 (128 B) under the
// 128-byte swizzle TMA wrote; 8-row groups are 1024 B apart (SBO); descriptor version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;                             // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256.
__device__ __forceinline__ constexpr uint32_t umma_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// HEAD_OUT > 0 fuses the network's last (HEAD_OUT-wide) linear layer into this layer's epilogue:
// instead of storing relu(A.W^T+b) the epilogue threads (one accumulator row each) contract their
// fp32 activations with head_w [HEAD_OUT][n] held in shared memory and write one partial result
// per (row, n-tile) to head_partial [m][n_tiles][8]; head_finish_kernel sums the partials in a
// fixed order, so results are deterministic.
constexpr int EXTRA_SMEM_BIAS = 4096;           // bias vector staged for n_pad <= 1024
constexpr int HEAD_OUT_FUSED = 6;               // SAC actor: 3 means + 3 log-stds
constexpr int HEAD_SMEM_MAX = HEAD_OUT_FUSED * 1024 * 4;

template <int HEAD_OUT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
dense_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ C, int ldc,
                  const int* __restrict__ m_dev, int m_max, int n_pad, int k_pad, int relu,
                  const float* __restrict__ head_w, int head_k, float* __restrict__ head_partial) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ttl_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty = [&](int s) { return bar0 + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int m = m_dev ? *m_dev : m_max;
  m = min(m, m_max);
  const int n_m = (m + BM - 1) / BM, n_n = (n_pad + BN - 1) / BN;
  const int total = n_m * n_n, kblocks = k_pad / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias (and the fused head's weights) staged once per CTA
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar0 + 256u - raw));
  float* s_head = s_bias + EXTRA_SMEM_BIAS / 4;
  const bool bias_in_smem = n_pad <= EXTRA_SMEM_BIAS / 4;
  if (bias_in_smem)
    for (int t = threadIdx.x; t < n_pad; t += GEMM_THREADS) s_bias[t] = bias[t];
  if (HEAD_OUT > 0) {
    for (int t = threadIdx.x; t < HEAD_OUT * n_pad; t += GEMM_THREADS) {
      const int o = t / n_pad, c = t - o * n_pad;
      s_head[t] = c < head_k ? head_w[(size_t)o * head_k + c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m_blk = tile / n_n, n_blk = tile - m_blk * n_n;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full(stage), STAGE_BYTES);
          const uint32_t sa = base + stage * STAGE_BYTES;
          tma_load_2d(sa, &tma_a, full(stage), kb * BK, m_blk * BM);
          tma_load_2d(sa + A_BYTES, &tma_b, full(stage), kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      constexpr uint32_t idesc = umma_idesc();
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t da = umma_desc(sa), db = umma_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per K=16 slice inside the 128-byte swizzle row: +2 in 16-byte units
            tc_mma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                        (uint32_t)((kb | k) != 0));
          }
          tc_commit(empty(stage));   // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull(acc));       // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {  // ===== epilogue: TMEM -> registers -> bias/ReLU -> bf16 -> HBM =====
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int m_blk = tile / n_n, n_blk = tile - m_blk * n_n;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < m;
      __nv_bfloat16* crow = C + (size_t)row * ldc;
      float hp[HEAD_OUT > 0 ? HEAD_OUT : 1];
#pragma unroll
      for (int o = 0; o < (HEAD_OUT > 0 ? HEAD_OUT : 1); ++o) hp[o] = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int col0 = n_blk * BN + ch * 32;
        if (col0 >= n_pad) break;  // warp-uniform
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + ch * 32), v);
        tc_wait_ld();
        float x[32];
        if (bias_in_smem) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col0 + 4 * j);
            x[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b4.x;
            x[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            x[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
            x[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __ldg(bias + col0 + j);
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (HEAD_OUT > 0) {
#pragma unroll
          for (int o = 0; o < HEAD_OUT; ++o) {
            const float* wrow = s_head + o * n_pad + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + 4 * j);
              hp[o] = fmaf(x[4 * j + 0], w4.x, hp[o]);
              hp[o] = fmaf(x[4 * j + 1], w4.y, hp[o]);
              hp[o] = fmaf(x[4 * j + 2], w4.z, hp[o]);
              hp[o] = fmaf(x[4 * j + 3], w4.w, hp[o]);
            }
          }
        } else if (row_ok) {
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(crow + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
      if (HEAD_OUT > 0 && row_ok) {
        float o8[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) o8[o] = o < HEAD_OUT ? hp[o < HEAD_OUT ? o : 0] : 0.f;
        float4* dst = reinterpret_cast<float4*>(head_partial + ((size_t)row * n_n + n_blk) * 8);
        dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
        dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}


// ==========================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 tile.
// Each CTA stages ITS 128 rows of A and ITS 128 of the 256 W rows per k-block (32 KB instead
// of 48 KB), the leader CTA's elected thread issues tcgen05.mma.cta_group::2 (M = 256), each
// CTA's tensor core accumulates its own 128 rows into its own TMEM, and each CTA's epilogue
// warps drain their half.  Shared-memory traffic per SM drops from 96+96 to 64+64 bytes/clk
// at full MMA rate, which is what holds the 1-CTA kernel to ~58 % tensor-pipe utilisation.
//
// Barrier protocol (same smem offsets in both CTAs):
//   full[s]    lives in the leader; its producer arms it with arrive.expect_tx for BOTH CTAs'
//              bytes (64 KB); all four TMA loads complete_tx on it (cp.async.bulk.tensor ...
//              .cta_group::2 with the leader's barrier address).  The peer's bytes may land
//              before the leader arms the phase: the tx-count goes transiently negative, which
//              mbarrier permits, and the phase cannot complete before the leader's arrival.
//   empty[s]   one per CTA, released by tcgen05.commit ... multicast to both CTAs
//   tfull[a]   one per CTA, signalled by the same multicast commit after the last k-block
//   tempty[a]  lives in the leader; 8 arrivals (4 epilogue warps x 2 CTAs, the peer's remotely)
// ==========================================================================================
constexpr int STAGES2 = 6;
constexpr int B2_BYTES = (BN / 2) * BK * 2;        // 16 KB: this CTA's half of the W tile
constexpr int STAGE2_BYTES = A_BYTES + B2_BYTES;   // 32 KB
constexpr int GEMM2_SMEM = STAGES2 * STAGE2_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Remote arrive with the default (cta-scope) release: a cluster-scope release compiles to
// MEMBAR.ALL.GPU, which was measured to serialise the pipeline.  The data this barrier guards is
// TMEM drained by tcgen05.wait::ld + tcgen05.fence, not generic-proxy memory.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* map, uint32_t leader_bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Instruction descriptor for the pair: D=f32, A=B=bf16, K-major, M=256, N=256.
__device__ __forceinline__ constexpr uint32_t umma_idesc_2cta() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
}

template <int HEAD_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
dense_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ C, int ldc,
                       const int* __restrict__ m_dev, int m_max, int n_pad, int k_pad, int relu,
                       const float* __restrict__ head_w, int head_k, float* __restrict__ head_partial) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ttl_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar0 = base + STAGES2 * STAGE2_BYTES;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES2 + s); };
  auto tfull = [&](int s) { return bar0 + 8u * (2 * STAGES2 + s); };
  auto tempty = [&](int s) { return bar0 + 8u * (2 * STAGES2 + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES2 + 2 * ACC_STAGES);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();       // 0 = leader
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // Everything above is independent of earlier kernels (barriers, tensor memory, descriptor prefetch)
  // and overlaps the predecessor's tail under programmatic dependent launch; from here on we read
  // what it wrote (row count, activations, freshly packed weights).
  ttl_grid_dep_wait();
  int m = m_dev ? *m_dev : m_max;
  m = min(m, m_max);
  const int n_m = (m + 2 * BM - 1) / (2 * BM), n_n = (n_pad + BN - 1) / BN;
  const int total = n_m * n_n, kblocks = k_pad / BK;

  // relu: bit 0 = apply ReLU; bits 8.. = how many of the fused head's outputs are wanted (0 = all):
  // the deterministic policy reads mu only, half of the 6-wide head
  const int head_n = (relu >> 8) ? min(relu >> 8, HEAD_OUT > 0 ? HEAD_OUT : 1) : (HEAD_OUT > 0 ? HEAD_OUT : 1);
  relu &= 1;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar0 + 256u - raw));
  float* s_head = s_bias + EXTRA_SMEM_BIAS / 4;
  const bool bias_in_smem = n_pad <= EXTRA_SMEM_BIAS / 4;
  if (bias_in_smem)
    for (int t = threadIdx.x; t < n_pad; t += GEMM_THREADS) s_bias[t] = bias[t];
  if (HEAD_OUT > 0) {
    for (int t = threadIdx.x; t < head_n * n_pad; t += GEMM_THREADS) {
      const int o = t / n_pad, c = t - o * n_pad;
      s_head[t] = c < head_k ? head_w[(size_t)o * head_k + c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer (both CTAs).  The whole warp runs the loop and the waits; the elected lane
    //       issues (ttl_tc.cuh, elect_one: no per-instruction ELECT / BRA.U.ANY loop) =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total; tile += n_clusters) {
      const int m_blk = tile / n_n, n_blk = tile - m_blk * n_n;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t leader_full = mapa_shared(full(stage), 0);
          if (cta == 0) mbar_arrive_expect_tx(full(stage), 2 * STAGE2_BYTES);
          const uint32_t sa = base + stage * STAGE2_BYTES;
          tma_load_2d_2cta(sa, &tma_a, leader_full, kb * BK, m_blk * 2 * BM + (int)cta * BM);
          tma_load_2d_2cta(sa + A_BYTES, &tma_b, leader_full, kb * BK, n_blk * BN + (int)cta * (BN / 2));
        }
        __syncwarp();
        if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (cta == 0) {  // ===== MMA issuer (leader CTA only; warp-uniform, elected lane issues) =====
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      constexpr uint32_t idesc = umma_idesc_2cta();
      for (int tile = cluster_id; tile < total; tile += n_clusters) {
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = base + stage * STAGE2_BYTES;
            const uint64_t da = umma_desc(sa), db = umma_desc(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_bf16_2cta(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                               (uint32_t)((kb | k) != 0));
            tc_commit_2cta(empty(stage));   // frees this stage in BOTH CTAs
            if (kb == kblocks - 1) tc_commit_2cta(tfull(acc));       // accumulators complete in both CTAs
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1u; }
        }
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {  // ===== epilogue (both CTAs, own 128 rows) =====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < total; tile += n_clusters) {
      const int m_blk = tile / n_n, n_blk = tile - m_blk * n_n;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * 2 * BM + (int)cta * BM + q * 32 + lane;
      const bool row_ok = row < m;
      __nv_bfloat16* crow = C + (size_t)row * ldc;
      float hp[HEAD_OUT > 0 ? HEAD_OUT : 1];
#pragma unroll
      for (int o = 0; o < (HEAD_OUT > 0 ? HEAD_OUT : 1); ++o) hp[o] = 0.f;
      // Software-pipelined over the 8 column chunks of 32: the tcgen05.ld of chunk ch+1 is in flight
      // while chunk ch gets its bias / ReLU / bf16 packing and leaves as 32-byte stores (whole
      // sectors; 16-byte stores left every sector half written until the next instruction).
      const int n_ch = min(BN / 32, (n_pad - n_blk * BN + 31) / 32);   // warp-uniform
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const bool wide_st = ((ldc & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 31) == 0);
      uint32_t va[32], vb[32];   // two named buffers: indexing one array by ch & 1 sent it to local memory
      // one chunk: wait for `cur`, start the load of the next chunk into `nxt`, then convert and store
      auto do_chunk = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int ch) {
        const int col0 = n_blk * BN + ch * 32;
        tc_wait_ld_regs(cur);
        if (ch + 1 < n_ch) tc_ld32(t_row + (uint32_t)((ch + 1) * 32), nxt);
        float x[32];
        if (bias_in_smem) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + col0 + 4 * j);
            x[4 * j + 0] = __uint_as_float(cur[4 * j + 0]) + b4.x;
            x[4 * j + 1] = __uint_as_float(cur[4 * j + 1]) + b4.y;
            x[4 * j + 2] = __uint_as_float(cur[4 * j + 2]) + b4.z;
            x[4 * j + 3] = __uint_as_float(cur[4 * j + 3]) + b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(cur[j]) + __ldg(bias + col0 + j);
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (HEAD_OUT > 0) {
#pragma unroll
          for (int o = 0; o < HEAD_OUT; ++o) {
            if (o >= head_n) break;   // warp-uniform
            const float* wrow = s_head + o * n_pad + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + 4 * j);
              hp[o] = fmaf(x[4 * j + 0], w4.x, hp[o]);
              hp[o] = fmaf(x[4 * j + 1], w4.y, hp[o]);
              hp[o] = fmaf(x[4 * j + 2], w4.z, hp[o]);
              hp[o] = fmaf(x[4 * j + 3], w4.w, hp[o]);
            }
          }
        } else if (row_ok) {
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (wide_st) {
            st_global_v8(crow + col0, packed);
            st_global_v8(crow + col0 + 16, packed + 8);
          } else {
            uint4* dst = reinterpret_cast<uint4*>(crow + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        }
      };
      if (n_ch > 0) tc_ld32(t_row, va);
#pragma unroll(HEAD_OUT > 0 ? 1 : BN / 64)
      for (int ch = 0; ch < BN / 32; ch += 2) {
        if (ch >= n_ch) break;
        do_chunk(va, vb, ch);
        if (ch + 1 >= n_ch) break;
        do_chunk(vb, va, ch + 1);
      }
      if (HEAD_OUT > 0 && row_ok) {
        float o8[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) o8[o] = o < HEAD_OUT ? hp[o < HEAD_OUT ? o : 0] : 0.f;
        float4* dst = reinterpret_cast<float4*>(head_partial + ((size_t)row * n_n + n_blk) * 8);
        dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
        dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta == 0) mbar_arrive(tempty(acc));
        else mbar_arrive_remote(mapa_shared(tempty(acc), 0));
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA leaves (or frees TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ==========================================================================================
// Packing kernels
// ==========================================================================================
__global__ void pack_weight_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                        int n_out, int n_in, int n_pad, int k_pad) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_pad * k_pad) return;
  const int r = (int)(t / k_pad), c = (int)(t - (long long)r * k_pad);
  const float x = (r < n_out && c < n_in) ? w[(size_t)r * n_in + c] : 0.f;
  out[t] = __float2bfloat16_rn(x);
}
// first-layer weights for the channel-padded input layout: packed column q reads original column
// p*C + ch for q = p*CP + ch (ch < C), n_points*C + (q - n_points*CP) behind the SH block, else 0
__global__ void pack_weight_layout_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                          int n_out, int n_in, int n_pad, int k_pad, int C, int CP, int n_pts) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_pad * k_pad) return;
  const int r = (int)(t / k_pad), q = (int)(t - (long long)r * k_pad);
  int c = -1;
  if (q < n_pts * CP) {
    const int p = q / CP, ch = q - p * CP;
    if (ch < C) c = p * C + ch;
  } else {
    c = n_pts * C + (q - n_pts * CP);
  }
  const float x = (r < n_out && c >= 0 && c < n_in) ? w[(size_t)r * n_in + c] : 0.f;
  out[t] = __float2bfloat16_rn(x);
}
__global__ void pack_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int n_out, int n_pad) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_pad) out[t] = t < n_out ? b[t] : 0.f;
}
// fp32 state rows -> bf16 rows padded to k_pad (8 outputs = one 16-byte store per thread)
__global__ void __launch_bounds__(256) pack_state_bf16_kernel(const float* __restrict__ state, int ld,
                                                              int width, const int* __restrict__ n_dev,
                                                              int n_max, __nv_bfloat16* __restrict__ out,
                                                              int k_pad) {
  int n = n_dev ? *n_dev : n_max;
  n = min(n, n_max);
  const int per_row = k_pad >> 3;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * per_row) return;
  const int r = (int)(t / per_row), g = (int)(t - (long long)r * per_row);
  const float* s = state + (size_t)r * ld + g * 8;
  uint32_t p[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = g * 8 + 2 * j;
    const float x0 = c < width ? s[2 * j] : 0.f;
    const float x1 = c + 1 < width ? s[2 * j + 1] : 0.f;
    __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    p[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(out + (size_t)r * k_pad + g * 8) = make_uint4(p[0], p[1], p[2], p[3]);
}

// ==========================================================================================
// Head: last linear layer (<= 8 outputs) in fp32 + SAC policy head (offpolicy.py:116-140)
// ==========================================================================================
constexpr int HEAD_MAX_OUT = 8;

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

__device__ __forceinline__ float softplus_f(float x) {  // F.softplus, threshold 20
  return x > 20.f ? x : log1pf(expf(x));
}

template <typename T>
__global__ void __launch_bounds__(256) head_kernel(const T* __restrict__ h, int ldh, int k,
                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                   int n_out, const int* __restrict__ n_dev, int n_max,
                                                   float prob, const float* __restrict__ eps,
                                                   float* __restrict__ action, float* __restrict__ logp,
                                                   float* __restrict__ pre) {
  extern __shared__ float s_w[];  // [n_out][k]
  for (int t = threadIdx.x; t < n_out * k; t += blockDim.x) s_w[t] = w[t];
  __syncthreads();
  int n = n_dev ? *n_dev : n_max;
  n = min(n, n_max);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int A = n_out >> 1;
  for (int r = blockIdx.x * (blockDim.x >> 5) + warp; r < n; r += warps_total) {
    float acc[HEAD_MAX_OUT];
#pragma unroll
    for (int o = 0; o < HEAD_MAX_OUT; ++o) acc[o] = 0.f;
    const T* row = h + (size_t)r * ldh;
    for (int c = lane; c < k; c += 32) {
      const float x = to_f32<T>(row[c]);
#pragma unroll
      for (int o = 0; o < HEAD_MAX_OUT; ++o)
        if (o < n_out) acc[o] = fmaf(x, s_w[o * k + c], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < HEAD_MAX_OUT; ++o)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
    if (lane == 0) {
      float lp = 0.f;
#pragma unroll
      for (int o = 0; o < HEAD_MAX_OUT; ++o)
        if (o < n_out) {
          acc[o] += b[o];
          if (pre) pre[(size_t)r * n_out + o] = acc[o];
        }
#pragma unroll
      for (int a = 0; a < HEAD_MAX_OUT / 2; ++a) {
        if (a >= A) break;
        const float mu = acc[a];
        const float log_std = fminf(fmaxf(acc[A + a], -20.f), 2.f);
        const float std = expf(log_std) * prob;
        const float e = eps ? eps[(size_t)r * A + a] : 0.f;
        const float pi = eps ? fmaf(std, e, mu) : mu;
        if (logp) {
          // Normal(mu,std).log_prob(pi) and the tanh correction of offpolicy.py:131-135
          const float z = pi - mu;
          lp += -(z * z) / (2.f * std * std) - logf(std) - 0.9189385332046727f;
          lp -= 2.f * (0.6931471805599453f - pi - softplus_f(-2.f * pi));
        }
        action[(size_t)r * A + a] = tanhf(pi);
      }
      if (logp) logp[r] = lp;
    }
  }
}

// Sums the per-n-tile partials of the fused head in tile order, adds the bias and applies the
// policy head.  One thread per row.
__global__ void __launch_bounds__(256) head_finish_kernel(const float* __restrict__ partial, int n_tiles,
                                                          const float* __restrict__ b, int n_out,
                                                          const int* __restrict__ n_dev, int n_max,
                                                          float prob, const float* __restrict__ eps,
                                                          float* __restrict__ action,
                                                          float* __restrict__ logp, float* __restrict__ pre) {
  int n = n_dev ? *n_dev : n_max;
  n = min(n, n_max);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  for (int t = 0; t < n_tiles; ++t) {
    const float4* p = reinterpret_cast<const float4*>(partial + ((size_t)r * n_tiles + t) * 8);
    const float4 a = p[0], c = p[1];
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    acc[4] += c.x; acc[5] += c.y; acc[6] += c.z; acc[7] += c.w;
  }
  const int A = n_out >> 1;
  float lp = 0.f;
#pragma unroll
  for (int o = 0; o < 8; ++o)
    if (o < n_out) {
      acc[o] += b[o];
      if (pre) pre[(size_t)r * n_out + o] = acc[o];
    }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (a >= A) break;
    const float mu = acc[a];
    float ls = 0.f;
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o == A + a) ls = acc[o];
    const float log_std = fminf(fmaxf(ls, -20.f), 2.f);
    const float std = expf(log_std) * prob;
    const float e = eps ? eps[(size_t)r * A + a] : 0.f;
    const float pi = eps ? fmaf(std, e, mu) : mu;
    if (logp) {
      const float z = pi - mu;
      lp += -(z * z) / (2.f * std * std) - logf(std) - 0.9189385332046727f;
      lp -= 2.f * (0.6931471805599453f - pi - softplus_f(-2.f * pi));
    }
    action[(size_t)r * A + a] = tanhf(pi);
  }
  if (logp) logp[r] = lp;
}

// ==========================================================================================
// fp32 tier: plain tiled SGEMM on CUDA cores (reference precision)
// ==========================================================================================
constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256) dense_f32_kernel(const float* __restrict__ A, int lda,
                                                        const float* __restrict__ W,
                                                        const float* __restrict__ bias,
                                                        float* __restrict__ C, int ldc, int m, int n,
                                                        int k, int relu) {
  __shared__ float sA[SG_K][SG_T + 1];
  __shared__ float sB[SG_K][SG_T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += SG_K) {
    for (int t = threadIdx.x; t < SG_T * SG_K; t += 256) {
      const int r = t / SG_K, c = t - r * SG_K;
      sA[c][r] = (m0 + r < m && k0 + c < k) ? A[(size_t)(m0 + r) * lda + k0 + c] : 0.f;
      sB[c][r] = (n0 + r < n && k0 + c < k) ? W[(size_t)(n0 + r) * k + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; bb[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + ty * 4 + i, c = n0 + tx * 4 + j;
      if (r < m && c < n) {
        float x = acc[i][j] + bias[c];
        if (relu) x = fmaxf(x, 0.f);
        C[(size_t)r * ldc + c] = x;
      }
    }
}

// ==========================================================================================
// Host side: TMA descriptors, plan
// ==========================================================================================
// bf16 row-major [rows][cols] tensor, box = [box_rows][64 cols], 128-byte swizzle.
int make_tmap(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return TTL_ERR_DRIVER;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : TTL_ERR_DRIVER;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

bool use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TTL_DENSE_1CTA");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// tb2: TMA map of W with a 128-row box (2-CTA kernel); tb: 256-row box (1-CTA kernel)
int launch_dense_bf16(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tb2, const float* bias,
                      __nv_bfloat16* C, int ldc, const int* m_dev, int m_max, int n_pad, int k_pad, int relu,
                      cudaStream_t s, const float* head_w = nullptr, int head_k = 0,
                      float* head_partial = nullptr) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(dense_bf16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GEMM_SMEM + EXTRA_SMEM_BIAS);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_bf16_kernel<HEAD_OUT_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GEMM_SMEM + EXTRA_SMEM_BIAS + HEAD_SMEM_MAX);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_bf16_2cta_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GEMM2_SMEM + EXTRA_SMEM_BIAS);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(dense_bf16_2cta_kernel<HEAD_OUT_FUSED>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GEMM2_SMEM + EXTRA_SMEM_BIAS + HEAD_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  if (use_2cta()) {
    const int tiles = ttl_div_up(m_max, 2 * BM) * ttl_div_up(n_pad, BN);
    int clusters = num_sms() / 2;
    if (tiles < clusters) clusters = tiles;
    if (clusters <= 0) return 0;
    const int grid = 2 * clusters;
    if (head_w) {
      TTL_LAUNCH("dense_bf16_head_kernel", s,
                 ttl_launch_chain(dense_bf16_2cta_kernel<HEAD_OUT_FUSED>, grid, GEMM_THREADS,
                                  GEMM2_SMEM + EXTRA_SMEM_BIAS + HEAD_OUT_FUSED * n_pad * 4, s, ta, tb2, bias, C, ldc,
                                  m_dev, m_max, n_pad, k_pad, relu, head_w, head_k, head_partial));
    } else {
      TTL_LAUNCH("dense_bf16_kernel", s,
                 ttl_launch_chain(dense_bf16_2cta_kernel<0>, grid, GEMM_THREADS, GEMM2_SMEM + EXTRA_SMEM_BIAS, s, ta,
                                  tb2, bias, C, ldc, m_dev, m_max, n_pad, k_pad, relu, (const float*)nullptr, 0,
                                  (float*)nullptr));
    }
    TTL_CHECK_LAST();
    return 0;
  }
  const int tiles = ttl_div_up(m_max, BM) * ttl_div_up(n_pad, BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (grid <= 0) return 0;
  if (head_w) {
    TTL_LAUNCH("dense_bf16_head_kernel", s,
               dense_bf16_kernel<HEAD_OUT_FUSED><<<grid, GEMM_THREADS,
                                                   GEMM_SMEM + EXTRA_SMEM_BIAS + HEAD_OUT_FUSED * n_pad * 4, s>>>(
                   ta, tb, bias, C, ldc, m_dev, m_max, n_pad, k_pad, relu, head_w, head_k, head_partial));
  } else {
    TTL_LAUNCH("dense_bf16_kernel", s,
               dense_bf16_kernel<0><<<grid, GEMM_THREADS, GEMM_SMEM + EXTRA_SMEM_BIAS, s>>>(
                   ta, tb, bias, C, ldc, m_dev, m_max, n_pad, k_pad, relu, nullptr, 0, nullptr));
  }
  TTL_CHECK_LAST();
  return 0;
}

}  // namespace

struct ttl_actor_plan {
  ttl_actor_weights w;
  int max_rows;
  int k_pad[TTL_ACTOR_MAX_LAYERS];   // padded fan-in of layer i  (multiple of 64)
  int n_pad[TTL_ACTOR_MAX_LAYERS];   // padded fan-out of layer i (= k_pad[i+1])
  __nv_bfloat16* wq[TTL_ACTOR_MAX_LAYERS];
  float* bq[TTL_ACTOR_MAX_LAYERS];
  __nv_bfloat16* act[2];             // ping-pong activations [max_rows][max_kpad]
  float* f32[2];                     // fp32-tier scratch [F32_CHUNK][max_width]
  float* head_partial;               // fused head partials [max_rows][4][8]
  bool fuse_head;
  int max_kpad, max_width;
  CUtensorMap map_w[TTL_ACTOR_MAX_LAYERS];
  CUtensorMap map_w2[TTL_ACTOR_MAX_LAYERS];  // 128-row box for the 2-CTA kernel
  __nv_bfloat16* w0_alt;                     // first-layer weights for bf16_layout 1
  CUtensorMap map_w0_alt, map_w0_alt2;
  bool has_alt;
  CUtensorMap map_a[TTL_ACTOR_MAX_LAYERS];  // A operand of layer i
  // first-layer operand maps for caller-owned bf16 state buffers (ttl_actor_forward_packed)
  struct ExtMap { const void* ptr; int rows; CUtensorMap map; };
  std::vector<ExtMap> ext_maps;
};

namespace {
constexpr int F32_CHUNK = 8192;

struct Layout {
  int64_t total;
  int64_t off_w[TTL_ACTOR_MAX_LAYERS], off_b[TTL_ACTOR_MAX_LAYERS], off_act[2], off_f32[2], off_hp, off_w0_alt;
  int k_pad[TTL_ACTOR_MAX_LAYERS], n_pad[TTL_ACTOR_MAX_LAYERS];
  int max_kpad, max_width;
};

int plan_layout(const ttl_actor_weights* w, int max_rows, Layout* L) {
  if (!w || w->n_layers < 2 || w->n_layers > TTL_ACTOR_MAX_LAYERS || max_rows <= 0) return TTL_ERR_BAD_ARG;
  if (w->out_dim[w->n_layers - 1] > HEAD_MAX_OUT || (w->out_dim[w->n_layers - 1] & 1)) return TTL_ERR_UNSUPPORTED;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += (bytes + 1023) / 1024 * 1024; return o; };
  L->max_kpad = 0;
  L->max_width = 0;
  for (int i = 0; i < w->n_layers; ++i) {
    if (i > 0 && w->in_dim[i] != w->out_dim[i - 1]) return TTL_ERR_BAD_ARG;
    L->k_pad[i] = round_up(w->in_dim[i], BK);
    L->n_pad[i] = round_up(w->out_dim[i], BK);
    if (L->k_pad[i] > L->max_kpad) L->max_kpad = L->k_pad[i];
    if (w->in_dim[i] > L->max_width) L->max_width = w->in_dim[i];
    if (w->out_dim[i] > L->max_width) L->max_width = w->out_dim[i];
  }
  for (int i = 0; i < w->n_layers - 1; ++i) {
    L->off_w[i] = take((int64_t)L->n_pad[i] * L->k_pad[i] * 2);
    L->off_b[i] = take((int64_t)round_up(L->n_pad[i], BN) * 4);
  }
  for (int j = 0; j < 2; ++j) L->off_act[j] = take((int64_t)max_rows * L->max_kpad * 2);
  for (int j = 0; j < 2; ++j) L->off_f32[j] = take((int64_t)F32_CHUNK * L->max_width * 4);
  L->off_hp = take((int64_t)max_rows * 4 * 8 * 4);
  L->off_w0_alt = take((int64_t)L->n_pad[0] * L->k_pad[0] * 2);
  L->total = off;
  return 0;
}
}  // namespace

namespace {
// Hidden layers on tensor cores (the last one carries the fused head when possible) + head.
// map_a0: TMA map of the first layer's bf16 operand [rows][k_pad[0]].
int run_bf16_layers(ttl_actor_plan* p, const CUtensorMap& map_a0, const int32_t* n_rows_dev,
                    int32_t n_rows_max, float probabilistic, const float* eps, float* action, float* logp,
                    float* pre, cudaStream_t s, bool alt_w0 = false) {
  const ttl_actor_weights& w = p->w;
  const int nl = w.n_layers;
  const int n_out = w.out_dim[nl - 1], k_last = w.in_dim[nl - 1];
  const size_t head_smem = (size_t)n_out * k_last * sizeof(float);
  const int head_grid = num_sms() * 2;
  static bool head_attr = false;
  if (!head_attr && head_smem > 48 * 1024) {
    cudaFuncSetAttribute(head_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    head_attr = true;
  }
  for (int i = 0; i < nl - 1; ++i) {
    // layer i: act[i&1] (pitch k_pad[i]) -> act[(i+1)&1] (pitch k_pad[i+1] = n_pad[i])
    const bool fused = p->fuse_head && i == nl - 2;
    const bool alt = alt_w0 && i == 0;
    // only mu is needed when nothing reads log_std: deterministic policy, no log-prob, no raw output
    const int head_n = (fused && probabilistic == 0.f && !logp && !pre) ? n_out / 2 : 0;
    int rc = launch_dense_bf16(i == 0 ? map_a0 : p->map_a[i], alt ? p->map_w0_alt : p->map_w[i],
                               alt ? p->map_w0_alt2 : p->map_w2[i], p->bq[i], p->act[(i + 1) & 1],
                               p->n_pad[i], n_rows_dev, n_rows_max, p->n_pad[i], p->k_pad[i], 1 | (head_n << 8), s,
                               fused ? w.w[nl - 1] : nullptr, k_last, fused ? p->head_partial : nullptr);
    if (rc) return rc;
  }
  if (p->fuse_head && !action) return 0;   // the caller reads the partial sums (ttl_actor_head_partial)
  if (p->fuse_head) {
    TTL_LAUNCH("head_finish_kernel", s,
               head_finish_kernel<<<ttl_div_up(n_rows_max, 256), 256, 0, s>>>(
                   p->head_partial, ttl_div_up(p->n_pad[nl - 2], BN), w.b[nl - 1], n_out, n_rows_dev,
                   n_rows_max, probabilistic, eps, action, logp, pre));
  } else {
    TTL_LAUNCH("head_kernel_bf16", s, head_kernel<__nv_bfloat16><<<head_grid, 256, head_smem, s>>>(
        p->act[(nl - 1) & 1], p->k_pad[nl - 1], k_last, w.w[nl - 1], w.b[nl - 1], n_out, n_rows_dev,
        n_rows_max, probabilistic, eps, action, logp, pre));
  }
  TTL_CHECK_LAST();
  return 0;
}
}  // namespace

extern "C" {

int64_t ttl_actor_workspace_bytes(const ttl_actor_weights* w, int32_t max_rows) {
  Layout L;
  if (plan_layout(w, max_rows, &L)) return -1;
  return L.total;
}

int ttl_actor_plan_create(ttl_actor_plan** out, const ttl_actor_weights* w, int32_t max_rows,
                          void* workspace, int64_t workspace_bytes, void* stream) {
  if (!out || !workspace) return TTL_ERR_BAD_ARG;
  Layout L;
  int rc = plan_layout(w, max_rows, &L);
  if (rc) return rc;
  if (workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return TTL_ERR_BAD_ARG;
  ttl_actor_plan* p = new (std::nothrow) ttl_actor_plan();
  if (!p) return TTL_ERR_BAD_ARG;
  p->w = *w;
  p->max_rows = max_rows;
  p->max_kpad = L.max_kpad;
  p->max_width = L.max_width;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  for (int j = 0; j < 2; ++j) {
    p->act[j] = reinterpret_cast<__nv_bfloat16*>(ws + L.off_act[j]);
    p->f32[j] = reinterpret_cast<float*>(ws + L.off_f32[j]);
  }
  const int nl = w->n_layers;
  for (int i = 0; i < nl; ++i) { p->k_pad[i] = L.k_pad[i]; p->n_pad[i] = L.n_pad[i]; }
  p->head_partial = reinterpret_cast<float*>(ws + L.off_hp);
  p->w0_alt = reinterpret_cast<__nv_bfloat16*>(ws + L.off_w0_alt);
  p->has_alt = false;
  // the last hidden layer can carry the head when it is 6 wide and the layer fits 4 n-tiles
  p->fuse_head = w->out_dim[nl - 1] == HEAD_OUT_FUSED && L.n_pad[nl - 2] <= 1024;
  for (int i = 0; i < nl - 1; ++i) {  // hidden layers run on tensor cores
    p->wq[i] = reinterpret_cast<__nv_bfloat16*>(ws + L.off_w[i]);
    p->bq[i] = reinterpret_cast<float*>(ws + L.off_b[i]);
    const long long tot = (long long)L.n_pad[i] * L.k_pad[i];
    TTL_LAUNCH("pack_weight_bf16_kernel", s, pack_weight_bf16_kernel<<<ttl_div_up(tot, 256), 256, 0, s>>>(w->w[i], p->wq[i], w->out_dim[i],
                                                                w->in_dim[i], L.n_pad[i], L.k_pad[i]));
    const int bp = round_up(L.n_pad[i], BN);
    TTL_LAUNCH("pack_bias_kernel", s, pack_bias_kernel<<<ttl_div_up(bp, 256), 256, 0, s>>>(w->b[i], p->bq[i], w->out_dim[i], bp));
    rc = make_tmap(&p->map_w[i], p->wq[i], (uint64_t)L.n_pad[i], (uint64_t)L.k_pad[i], BN);
    if (rc) { delete p; return rc; }
    rc = make_tmap(&p->map_w2[i], p->wq[i], (uint64_t)L.n_pad[i], (uint64_t)L.k_pad[i], BN / 2);
    if (rc) { delete p; return rc; }
    // A operand of layer i lives in act[i & 1] with row pitch k_pad[i]
    rc = make_tmap(&p->map_a[i], p->act[i & 1], (uint64_t)max_rows, (uint64_t)L.k_pad[i], BM);
    if (rc) { delete p; return rc; }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { delete p; return (int)e; }
  *out = p;
  return 0;
}

void ttl_actor_plan_destroy(ttl_actor_plan* plan) { delete plan; }

int ttl_actor_forward(ttl_actor_plan* p, const float* state, int32_t ld_state, const int32_t* n_rows_dev,
                      int32_t n_rows_max, float probabilistic, const float* eps, float* action,
                      float* logp, float* pre, int32_t precision, void* stream) {
  if (!p || !state || !action || n_rows_max > p->max_rows) return TTL_ERR_BAD_ARG;
  if (probabilistic != 0.f && !eps) return TTL_ERR_BAD_ARG;
  if (n_rows_max <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const ttl_actor_weights& w = p->w;
  const int nl = w.n_layers;
  const int n_out = w.out_dim[nl - 1], k_last = w.in_dim[nl - 1];
  const size_t head_smem = (size_t)n_out * k_last * sizeof(float);
  const int head_grid = num_sms() * 2;

  if (precision == TTL_PRECISION_BF16) {
    const long long tot = (long long)n_rows_max * (p->k_pad[0] >> 3);
    TTL_LAUNCH("pack_state_bf16_kernel", s, pack_state_bf16_kernel<<<ttl_div_up(tot, 256), 256, 0, s>>>(state, ld_state, w.in_dim[0], n_rows_dev,
                                                              n_rows_max, p->act[0], p->k_pad[0]));
    return run_bf16_layers(p, p->map_a[0], n_rows_dev, n_rows_max, probabilistic, eps, action, logp, pre, s);
  }
  if (precision == TTL_PRECISION_FP32) {
    // Reference-precision tier; needs the row count on the host.
    int n = n_rows_max;
    if (n_rows_dev) {
      cudaError_t e = cudaMemcpyAsync(&n, n_rows_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) return (int)e;
      e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) return (int)e;
      if (n > n_rows_max) n = n_rows_max;
    }
    static bool head_attr32 = false;
    if (!head_attr32 && head_smem > 48 * 1024) {
      cudaFuncSetAttribute(head_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      head_attr32 = true;
    }
    const int A = n_out / 2;
    for (int r0 = 0; r0 < n; r0 += F32_CHUNK) {
      const int m = (n - r0) < F32_CHUNK ? (n - r0) : F32_CHUNK;
      const float* in = state + (size_t)r0 * ld_state;
      int ld_in = ld_state;
      for (int i = 0; i < nl - 1; ++i) {
        float* o = p->f32[i & 1];
        dim3 grid(ttl_div_up(w.out_dim[i], SG_T), ttl_div_up(m, SG_T));
        TTL_LAUNCH("dense_f32_kernel", s, dense_f32_kernel<<<grid, 256, 0, s>>>(in, ld_in, w.w[i], w.b[i], o, p->max_width, m, w.out_dim[i],
                                              w.in_dim[i], 1));
        in = o;
        ld_in = p->max_width;
      }
      TTL_LAUNCH("head_kernel_f32", s, head_kernel<float><<<head_grid, 256, head_smem, s>>>(
          in, ld_in, k_last, w.w[nl - 1], w.b[nl - 1], n_out, nullptr, m, probabilistic,
          eps ? eps + (size_t)r0 * A : nullptr, action + (size_t)r0 * A, logp ? logp + r0 : nullptr,
          pre ? pre + (size_t)r0 * n_out : nullptr));
    }
    TTL_CHECK_LAST();
    return 0;
  }
  return TTL_ERR_UNSUPPORTED;
}

int ttl_actor_forward_packed(ttl_actor_plan* p, const void* state_bf16, int32_t ld, int32_t rows_alloc,
                             const int32_t* n_rows_dev, int32_t n_rows_max, float probabilistic,
                             const float* eps, float* action, float* logp, float* pre, int32_t layout,
                             void* stream) {
  if (!p || !state_bf16 || n_rows_max > p->max_rows || n_rows_max > rows_alloc) return TTL_ERR_BAD_ARG;
  if (!action && (logp || pre || eps || !p->fuse_head)) return TTL_ERR_BAD_ARG;
  if (layout != 0 && !(layout == 1 && p->has_alt)) return TTL_ERR_BAD_ARG;
  if (ld != p->k_pad[0] || (reinterpret_cast<uintptr_t>(state_bf16) & 15)) return TTL_ERR_BAD_ARG;
  if (probabilistic != 0.f && !eps) return TTL_ERR_BAD_ARG;
  if (n_rows_max <= 0) return 0;
  const CUtensorMap* map = nullptr;
  for (auto& e : p->ext_maps)
    if (e.ptr == state_bf16 && e.rows == rows_alloc) map = &e.map;
  if (!map) {
    if (p->ext_maps.size() >= 16) p->ext_maps.clear();
    ttl_actor_plan::ExtMap e;
    e.ptr = state_bf16;
    e.rows = rows_alloc;
    int rc = make_tmap(&e.map, state_bf16, (uint64_t)rows_alloc, (uint64_t)ld, BM);
    if (rc) return rc;
    p->ext_maps.push_back(e);
    map = &p->ext_maps.back().map;
  }
  return run_bf16_layers(p, *map, n_rows_dev, n_rows_max, probabilistic, eps, action, logp, pre,
                         (cudaStream_t)stream, layout == 1);
}

int ttl_actor_head_partial(const ttl_actor_plan* p, const float** partial, int32_t* n_tiles,
                           const float** bias) {
  if (!p || !partial || !n_tiles || !bias) return TTL_ERR_BAD_ARG;
  if (!p->fuse_head) return TTL_ERR_UNSUPPORTED;
  const int nl = p->w.n_layers;
  *partial = p->head_partial;
  *n_tiles = ttl_div_up(p->n_pad[nl - 2], BN);
  *bias = p->w.b[nl - 1];
  return 0;
}

int ttl_actor_plan_set_layout(ttl_actor_plan* p, int32_t C, int32_t CP, int32_t n_points, void* stream) {
  if (!p || C <= 0 || CP < C || n_points <= 0) return TTL_ERR_BAD_ARG;
  const int n_in = p->w.in_dim[0];
  if (n_points * C > n_in || n_points * CP + (n_in - n_points * C) > p->k_pad[0]) return TTL_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const long long tot = (long long)p->n_pad[0] * p->k_pad[0];
  TTL_LAUNCH("pack_weight_layout_kernel", s,
             pack_weight_layout_kernel<<<ttl_div_up(tot, 256), 256, 0, s>>>(p->w.w[0], p->w0_alt, p->w.out_dim[0], n_in,
                                                                          p->n_pad[0], p->k_pad[0], C, CP, n_points));
  int rc = make_tmap(&p->map_w0_alt, p->w0_alt, (uint64_t)p->n_pad[0], (uint64_t)p->k_pad[0], BN);
  if (rc) return rc;
  rc = make_tmap(&p->map_w0_alt2, p->w0_alt, (uint64_t)p->n_pad[0], (uint64_t)p->k_pad[0], BN / 2);
  if (rc) return rc;
  p->has_alt = true;
  TTL_CHECK_LAST();
  return 0;
}

int ttl_actor_plan_refresh(ttl_actor_plan* p, void* stream) {
  // the fp32 weights changed in place (an optimiser step): repack the bf16 copies
  if (!p) return TTL_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const ttl_actor_weights& w = p->w;
  for (int i = 0; i < w.n_layers - 1; ++i) {
    const long long tot = (long long)p->n_pad[i] * p->k_pad[i];
    TTL_LAUNCH("pack_weight_bf16_kernel", s,
               pack_weight_bf16_kernel<<<ttl_div_up(tot, 256), 256, 0, s>>>(w.w[i], p->wq[i], w.out_dim[i], w.in_dim[i],
                                                                          p->n_pad[i], p->k_pad[i]));
    const int bp = round_up(p->n_pad[i], BN);
    TTL_LAUNCH("pack_bias_kernel", s, pack_bias_kernel<<<ttl_div_up(bp, 256), 256, 0, s>>>(w.b[i], p->bq[i], w.out_dim[i], bp));
  }
  p->has_alt = false;   // ttl_actor_plan_set_layout repacks the permuted first layer on demand
  TTL_CHECK_LAST();
  return 0;
}

int ttl_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int32_t m, int32_t n,
                  int32_t k, int32_t ldc, int32_t relu, const int32_t* m_dev, void* stream) {
  if (!A || !W || !bias || !C || (k % BK) || (n % BK) || ldc < n || (ldc % 8)) return TTL_ERR_BAD_ARG;
  if (m <= 0) return 0;
  CUtensorMap ta, tb, tb2;
  int rc = make_tmap(&ta, A, (uint64_t)m, (uint64_t)k, BM);
  if (rc) return rc;
  rc = make_tmap(&tb, W, (uint64_t)n, (uint64_t)k, BN);
  if (rc) return rc;
  rc = make_tmap(&tb2, W, (uint64_t)n, (uint64_t)k, BN / 2);
  if (rc) return rc;
  // bias must be readable up to the tile edge: caller pads it to a multiple of 256 floats
  return launch_dense_bf16(ta, tb, tb2, bias, static_cast<__nv_bfloat16*>(C), ldc, m_dev, m, n, k, relu,
                           (cudaStream_t)stream);
}

}  // extern "C"
