// Persistent multi-layer dense kernel of the SAC actor (sm_100a): every tensor-core layer of
//   h_{l+1} = relu(h_l . W_l^T + b_l)        (algorithms/shared/offpolicy.py:94-140 over the
//                                              nn.Sequential of shared/utils.py:41-51)
// in ONE launch.  A cluster of two CTAs (one TPC) computes 256 x BN output tiles with
// tcgen05.mma.cta_group::2 (M = 256, N = BN, fp32 accumulators double-buffered in tensor memory),
// operands staged by TMA (128-byte swizzle) through a shared-memory ring, the epilogue warps read the
// accumulators back with tcgen05.ld, add bias, ReLU, round to the operand type and store.
//
// Operand kinds (template KIND): bf16 and fp16 (kind::f16, K = 16 per instruction, 64 elements per
// 128-byte row) and tf32 (kind::tf32, K = 8, 32 fp32 words per row; weights and activations are
// rounded to tf32 with round-to-nearest when they are written, so the tensor core's truncation of the
// low 13 bits is a no-op).  fp16 and tf32 both carry 11 significant bits: they meet the 1e-3 tolerance
// that bf16 (8 bits) cannot; fp16 at the bf16 rate within its range (stores saturate at 65504 and raise
// a flag), tf32 at half the rate with fp32's range.
//
// Tile schedule: the tiles of all layers form one list that the clusters walk round-robin.  The
// m-tiles are cut into groups; the list interleaves the layers group by group, software-pipelined by
// one group -- [L0 g0] [L0 g1][L1 g0] [L0 g2][L1 g1][L2 g0] ... -- so that (a) a layer's input rows
// were finished about two tile waves earlier (the dependency wait below practically never blocks),
// (b) activations are re-read while they still sit in L2, and (c) there is no wave quantisation or
// launch ramp between layers.  Dependencies are explicit: the epilogue of a tile publishes
// flags[layer][m-tile] (+1 per CTA per n-tile, release); the TMA producer of a tile of layer l+1 waits
// (acquire) until the count shows every n-tile of both CTAs, then crosses to the async proxy and
// loads.  Every tile depends only on tiles earlier in the list, every cluster walks the list in order
// and the grid never exceeds the machine, so the earliest unfinished tile can always run.  The last
// cluster to finish resets the flags (ticket).
//
// BN is a run-time value (256 / 128 / 64: the UMMA N and the TMA box of W): when few rows are alive the
// host picks narrower tiles so that every cluster has work (the low-occupancy tail of a tracking run).
//
// Fused head: when head_w != NULL the LAST layer of the launch does not store its activations; its
// epilogue contracts them (fp32) with the 6-wide output layer's weights and writes per-tile partial
// sums.  The partials are formed as a fixed tree over 64-column groups so that the final sum does not
// depend on BN: p_g = fma chain over group g's 64 columns; a tile's partial = (p0 + p1) + (p2 + p3)
// over the groups it covers (missing groups = 0); consumers rebuild the same tree per 256-column
// super-tile and add the super-tiles left to right (head_tree_sum below).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ttl_common.cuh"
#include "ttl_tc.cuh"

namespace ttl_mlp {
using namespace ttl_tc;

enum { KIND_BF16 = 0, KIND_F16 = 1, KIND_TF32 = 2 };

constexpr int MLP_MAX_LAYERS = 4;
constexpr int MLP_BM = 128;                 // rows per CTA; the pair computes 256
constexpr int MLP_ROW_BYTES = 128;          // one swizzled smem row: 64 x 16-bit or 32 x 32-bit
constexpr int MLP_A_BYTES = MLP_BM * MLP_ROW_BYTES;        // 16 KB
constexpr int MLP_STAGE_BYTES = 2 * MLP_A_BYTES;           // A + this CTA's half of W (<= 128 rows)
constexpr int MLP_THREADS = 256;            // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps4-7 epilogue
constexpr int MLP_TMEM_COLS = 512;
constexpr int MLP_ACC_STRIDE = 256;         // TMEM columns per accumulator stage
constexpr int MLP_BAR_BYTES = 256;
constexpr int MLP_HEAD = 6;                 // SAC actor: 3 means + 3 log-stds

__host__ __device__ constexpr int kind_esize(int kind) { return kind == KIND_TF32 ? 4 : 2; }
__host__ __device__ constexpr int kind_bk(int kind) { return MLP_ROW_BYTES / kind_esize(kind); }

struct MlpMaps {
  CUtensorMap a[MLP_MAX_LAYERS];   // A operand of layer l: [rows][k_pad], box 128 rows x 128 bytes
  CUtensorMap w[MLP_MAX_LAYERS];   // W of layer l: [n_pad][k_pad], box (bn / 2) rows x 128 bytes
};

struct MlpArgs {
  int n_layers, bn, group_m, n_stages;
  int kblocks[MLP_MAX_LAYERS];     // k_pad / elements per 128-byte row
  int n_pad[MLP_MAX_LAYERS];       // output width, multiple of 64
  int ldc[MLP_MAX_LAYERS];         // output row pitch in elements
  void* C[MLP_MAX_LAYERS];         // output activations (ignored for the fused-head layer)
  const float* bias[MLP_MAX_LAYERS];
  const int* m_dev;
  int m_max;
  int relu_mask;
  int bias_stride;                 // floats between the staged bias vectors (>= max n_pad)
  const float* head_w;             // [HEAD][head_k] fp32 or NULL
  int head_k, head_n;              // head_n: how many of the HEAD outputs are wanted
  float* head_partial;             // [rows][n_tiles][8]
  unsigned* flags;                 // [n_layers - 1][flag_stride] tile-completion counts, then the ticket
  int flag_stride;
  unsigned* overflow;              // set to 1 when an fp16 store saturated
};

// ---- PTX wrappers specific to the pair -----------------------------------------------------
static __device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
static __device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
static __device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Remote arrive with the default (cta-scope) release: a cluster-scope release compiles to
// MEMBAR.ALL.GPU, which was measured to serialise the pipeline.  The data this barrier guards is
// TMEM drained by tcgen05.wait::ld + tcgen05.fence, not generic-proxy memory.
static __device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
static __device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* map,
                                                        uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
static __device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
template <int KIND>
static __device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                   uint32_t idesc, uint32_t accumulate) {
  if (KIND == KIND_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// Instruction descriptor of the pair: D = f32, A and B K-major in the operand format of KIND
// (0 = f16, 1 = bf16, 2 = tf32), M = 256, N = bn.
template <int KIND>
static __device__ __forceinline__ uint32_t umma_idesc_pair(int bn) {
  const uint32_t fmt = KIND == KIND_F16 ? 0u : (KIND == KIND_BF16 ? 1u : 2u);
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((2 * MLP_BM) >> 4) << 24);
}
static __device__ __forceinline__ uint32_t round_tf32(float x) { return ttl_round_tf32(x); }
// two fp32 -> packed 16-bit pair (lo = a, hi = b); fp16 saturates to +-65504 instead of overflowing
template <int KIND>
static __device__ __forceinline__ uint32_t pack_pair(float a, float b) {
  return KIND == KIND_F16 ? ttl_pack_f16x2_sat(a, b) : ttl_pack_bf16x2(a, b);
}
static __device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
static __device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static __device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// The tile list (see the header comment).  Every role of a cluster walks it with its own copy.
// Host-callable too: tests/test_mlp_schedule_cpu.py enumerates it on the CPU and checks that every tile
// appears exactly once and that a tile's inputs always come earlier in the list.
struct Sched {
  int n_layers, n_m, group_m, n_groups;
  int nn0, nn1, nn2, nn3;          // n-tiles per layer (scalars: a register-indexed array would live in local memory)
  int s, l, off;

  __host__ __device__ __forceinline__ int n_n_of(int layer) const {
    return layer == 0 ? nn0 : (layer == 1 ? nn1 : (layer == 2 ? nn2 : nn3));
  }
  __host__ __device__ __forceinline__ int block_tiles() const {
    const int g = s - l;
    if (g < 0 || g >= n_groups) return 0;
    const int rows = n_m - g * group_m;
    return (rows < group_m ? rows : group_m) * n_n_of(l);
  }
  // move to the block that holds tile offset `off` (counted from the current block's start)
  __host__ __device__ __forceinline__ bool normalize() {
    while (s < n_groups + n_layers - 1) {
      const int bt = block_tiles();
      if (off < bt) return true;
      off -= bt;
      if (++l == n_layers) { l = 0; ++s; }
    }
    return false;
  }
  __host__ __device__ __forceinline__ bool start(int first) {
    s = 0; l = 0; off = first;
    return n_groups > 0 && normalize();
  }
  __host__ __device__ __forceinline__ bool advance(int step) {
    off += step;
    return normalize();
  }
  __host__ __device__ __forceinline__ void tile(int& layer, int& m_blk, int& n_blk) const {
    layer = l;
    const int nn = n_n_of(l);
    const int q = off / nn;
    m_blk = (s - l) * group_m + q;
    n_blk = off - q * nn;
  }
};

template <int KIND, int HEAD_OUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(MLP_THREADS, 1)
mlp_pair_kernel(const __grid_constant__ MlpMaps maps, const MlpArgs args) {
  constexpr int ESIZE = kind_esize(KIND);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ttl_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int n_stages = args.n_stages;
  const uint32_t bar0 = base + (uint32_t)n_stages * MLP_STAGE_BYTES;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (n_stages + s); };
  auto tfull = [&](int s) { return bar0 + 8u * (2 * n_stages + s); };
  auto tempty = [&](int s) { return bar0 + 8u * (2 * n_stages + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * n_stages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();       // 0 = leader
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int n_layers = args.n_layers, bn = args.bn;
  if (warp == 0 && lane == 0) {
    for (int l = 0; l < n_layers; ++l) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[l])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.w[l])) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 8); }
    mbar_init(bar0 + 8u * (2 * n_stages + 5), 1);      // bias / head weights staged (see below)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)MLP_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // Everything above is independent of earlier kernels (barriers, tensor memory, descriptor prefetch)
  // and overlaps the predecessor's tail under programmatic dependent launch; from here on we read
  // what it wrote (row count, state rows, freshly packed weights).
  ttl_grid_dep_wait();
  int m = args.m_dev ? *args.m_dev : args.m_max;
  m = min(m, args.m_max);

  Sched sc;
  sc.n_layers = n_layers;
  sc.n_m = (m + 2 * MLP_BM - 1) / (2 * MLP_BM);
  sc.group_m = max(args.group_m, 1);
  sc.n_groups = (sc.n_m + sc.group_m - 1) / sc.group_m;
  sc.nn0 = (args.n_pad[0] + bn - 1) / bn;
  sc.nn1 = n_layers > 1 ? (args.n_pad[1] + bn - 1) / bn : 1;
  sc.nn2 = n_layers > 2 ? (args.n_pad[2] + bn - 1) / bn : 1;
  sc.nn3 = n_layers > 3 ? (args.n_pad[3] + bn - 1) / bn : 1;

  // Bias vectors and the fused head's weights go to shared memory by bulk copies that complete on their own
  // mbarrier: issued here by one thread, awaited by the epilogue warps before their first tile, so the
  // ~30 KB never sit on the critical path (staging them with ordinary loads cost ~4 us at the head of every
  // launch, a quarter of a small-batch step).
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar0 + MLP_BAR_BYTES - raw));
  float* s_head = s_bias + n_layers * args.bias_stride;
  const uint32_t stage_bar = bar0 + 8u * (2 * n_stages + 5);
  const int head_np = args.n_pad[n_layers - 1];
  const bool has_head = HEAD_OUT > 0 && args.head_w != nullptr;
  const int head_n = has_head ? min(max(args.head_n, 1), HEAD_OUT > 0 ? HEAD_OUT : 1) : 0;
  // the head weights can be copied as they are when their rows are as wide as the padded layer
  const bool head_bulk = has_head && args.head_k == head_np && ((reinterpret_cast<uintptr_t>(args.head_w) & 15) == 0);
  if (warp == 3 && lane == 0) {
    uint32_t bytes = 0;
    for (int l = 0; l < n_layers; ++l) bytes += (uint32_t)args.n_pad[l] * 4u;
    if (head_bulk) bytes += (uint32_t)(head_n * head_np) * 4u;
    mbar_arrive_expect_tx(stage_bar, bytes);
    for (int l = 0; l < n_layers; ++l)
      bulk_load(ttl_smem_u32(s_bias + l * args.bias_stride), args.bias[l], (uint32_t)args.n_pad[l] * 4u, stage_bar);
    if (head_bulk) bulk_load(ttl_smem_u32(s_head), args.head_w, (uint32_t)(head_n * head_np) * 4u, stage_bar);
  }
  if (has_head && !head_bulk) {
    for (int t = threadIdx.x; t < head_n * head_np; t += MLP_THREADS) {
      const int o = t / head_np, c = t - o * head_np;
      s_head[t] = c < args.head_k ? args.head_w[(size_t)o * args.head_k + c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t w_bytes = (uint32_t)(bn / 2) * MLP_ROW_BYTES;      // this CTA's half of a W tile

  if (warp == 0) {
    // ===== TMA producer (both CTAs).  The whole warp runs the loop and the waits; the elected lane
    //       issues (ttl_tc.cuh, elect_one: no per-instruction ELECT / BRA.U.ANY loop) =====
    int stage = 0;
    uint32_t phase = 0;
    for (bool ok = sc.start(cluster_id); ok; ok = sc.advance(n_clusters)) {
      int l, m_blk, n_blk;
      sc.tile(l, m_blk, n_blk);
      if (l > 0) {
        // the rows of this m-tile were written by the previous layer's epilogues (generic proxy, other
        // SMs): wait for all of its n-tiles in both CTAs, then order our async-proxy reads after them
        const unsigned* f = args.flags + (size_t)(l - 1) * args.flag_stride + m_blk;
        const unsigned want = 2u * (unsigned)sc.n_n_of(l - 1);
        uint64_t t0 = 0;
        for (uint32_t it = 0; ld_acquire_u32(f) < want; ++it) {
          if ((it & 255u) == 255u) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();   // 2 s: a protocol bug must not hang the GPU
          }
          __nanosleep(32);
        }
        __syncwarp();
        fence_proxy_async_global();
      }
      const CUtensorMap* ma = &maps.a[l];
      const CUtensorMap* mw = &maps.w[l];
      const int kblocks = args.kblocks[l];
      constexpr int BKE = kind_bk(KIND);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty(stage), phase ^ 1u);
        if (elect_one()) {
          const uint32_t leader_full = mapa_shared(full(stage), 0);
          if (cta == 0) mbar_arrive_expect_tx(full(stage), 2u * (MLP_A_BYTES + w_bytes));
          const uint32_t sa = base + stage * MLP_STAGE_BYTES;
          tma_load_2d_pair(sa, ma, leader_full, kb * BKE, m_blk * 2 * MLP_BM + (int)cta * MLP_BM);
          tma_load_2d_pair(sa + MLP_A_BYTES, mw, leader_full, kb * BKE, n_blk * bn + (int)cta * (bn / 2));
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (cta == 0) {  // ===== MMA issuer (leader CTA only; warp-uniform, elected lane issues) =====
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = umma_idesc_pair<KIND>(bn);
      for (bool ok = sc.start(cluster_id); ok; ok = sc.advance(n_clusters)) {
        int l, m_blk, n_blk;
        sc.tile(l, m_blk, n_blk);
        const int kblocks = args.kblocks[l];
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MLP_ACC_STRIDE);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = base + stage * MLP_STAGE_BYTES;
            const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + MLP_A_BYTES);
            // four instructions per 128-byte row (K = 16 halves or 8 tf32 words = 32 bytes each):
            // +32 bytes inside the swizzle row = +2 in the descriptor's 16-byte units
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_pair<KIND>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                (uint32_t)((kb | k) != 0));
            tc_commit_pair(empty(stage));                        // frees this stage in BOTH CTAs
            if (kb == kblocks - 1) tc_commit_pair(tfull(acc));   // accumulators complete in both CTAs
          }
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {  // ===== epilogue (both CTAs, own 128 rows) =====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t sat = 0;      // fp16: running max of the stored magnitudes (packed pair)
    mbar_wait(stage_bar, 0);       // bias vectors / head weights have landed
    for (bool ok = sc.start(cluster_id); ok; ok = sc.advance(n_clusters)) {
      int l, m_blk, n_blk;
      sc.tile(l, m_blk, n_blk);
      const int n_pad = args.n_pad[l];
      const bool relu = (args.relu_mask >> l) & 1;
      const bool head_tile = has_head && l == n_layers - 1;
      const float* bias_l = s_bias + l * args.bias_stride;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * 2 * MLP_BM + (int)cta * MLP_BM + q * 32 + lane;
      const bool row_ok = row < m;
      uint8_t* crow = static_cast<uint8_t*>(args.C[l]) + (size_t)row * args.ldc[l] * ESIZE;
      constexpr int H = HEAD_OUT > 0 ? HEAD_OUT : 1;
      float hp[H], t0[H], t1[H];
#pragma unroll
      for (int o = 0; o < H; ++o) { hp[o] = 0.f; t0[o] = 0.f; t1[o] = 0.f; }
      // Software-pipelined over the column chunks of 32: the tcgen05.ld of chunk ch+1 is in flight
      // while chunk ch gets its bias / ReLU / rounding and leaves as 32-byte stores (whole sectors).
      const int n_ch = min(bn / 32, (n_pad - n_blk * bn + 31) / 32);   // warp-uniform
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MLP_ACC_STRIDE);
      uint32_t va[32], vb[32];   // two named buffers: indexing one array by ch & 1 sent it to local memory
      auto do_chunk = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int ch) {
        const int col0 = n_blk * bn + ch * 32;
        tc_wait_ld_regs(cur);
        if (ch + 1 < n_ch) tc_ld32(t_row + (uint32_t)((ch + 1) * 32), nxt);
        float x[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_l + col0 + 4 * j);
          x[4 * j + 0] = __uint_as_float(cur[4 * j + 0]) + b4.x;
          x[4 * j + 1] = __uint_as_float(cur[4 * j + 1]) + b4.y;
          x[4 * j + 2] = __uint_as_float(cur[4 * j + 2]) + b4.z;
          x[4 * j + 3] = __uint_as_float(cur[4 * j + 3]) + b4.w;
        }
        if (relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (HEAD_OUT > 0 && head_tile) {
#pragma unroll
          for (int o = 0; o < H; ++o) {
            if (o >= head_n) break;   // warp-uniform
            const float* wrow = s_head + o * head_np + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w4 = *reinterpret_cast<const float4*>(wrow + 4 * j);
              hp[o] = fmaf(x[4 * j + 0], w4.x, hp[o]);
              hp[o] = fmaf(x[4 * j + 1], w4.y, hp[o]);
              hp[o] = fmaf(x[4 * j + 2], w4.z, hp[o]);
              hp[o] = fmaf(x[4 * j + 3], w4.w, hp[o]);
            }
          }
          if ((ch & 1) || ch + 1 == n_ch) {     // a 64-column group is complete: fold it into the tree
            const int g = ch >> 1;
#pragma unroll
            for (int o = 0; o < H; ++o) {
              if (g == 0) t0[o] = hp[o];
              else if (g == 1) t0[o] = t0[o] + hp[o];
              else if (g == 2) t1[o] = hp[o];
              else t1[o] = t1[o] + hp[o];
              hp[o] = 0.f;
            }
          }
        } else if (row_ok) {
          if (KIND == KIND_TF32) {
            uint32_t r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = round_tf32(x[j]);
            uint8_t* d = crow + (size_t)col0 * 4;
            st_global_v8(d, r);
            st_global_v8(d + 32, r + 8);
            st_global_v8(d + 64, r + 16);
            st_global_v8(d + 96, r + 24);
          } else {
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) packed[j] = pack_pair<KIND>(x[2 * j], x[2 * j + 1]);
            if (KIND == KIND_F16) {
              // running maximum of what was stored (one HMNMX2 per pair).  After ReLU the values are
              // non-negative; without it the sign bit is dropped first.
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t pj = relu ? packed[j] : (packed[j] & 0x7fff7fffu);
                __half2 mx = __hmax2(*reinterpret_cast<const __half2*>(&pj), *reinterpret_cast<__half2*>(&sat));
                sat = *reinterpret_cast<uint32_t*>(&mx);
              }
            }
            uint8_t* d = crow + (size_t)col0 * 2;
            st_global_v8(d, packed);
            st_global_v8(d + 32, packed + 8);
          }
        }
      };
      if (n_ch > 0) tc_ld32(t_row, va);
#pragma unroll 1
      for (int ch = 0; ch < 8; ch += 2) {
        if (ch >= n_ch) break;
        do_chunk(va, vb, ch);
        if (ch + 1 >= n_ch) break;
        do_chunk(vb, va, ch + 1);
      }
      if (HEAD_OUT > 0 && head_tile && row_ok) {
        float o8[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) o8[o] = o < H ? t0[o < H ? o : 0] + t1[o < H ? o : 0] : 0.f;
        float4* dst = reinterpret_cast<float4*>(args.head_partial + ((size_t)row * sc.n_n_of(l) + n_blk) * 8);
        dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
        dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (cta == 0) mbar_arrive(tempty(acc));
        else mbar_arrive_remote(mapa_shared(tempty(acc), 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      if (l + 1 < n_layers) {
        // publish this CTA's rows of the tile to the next layer's TMA producers
        fence_proxy_async_global();
        asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps
        if (warp == 4 && lane == 0)
          red_release_add_u32(args.flags + (size_t)l * args.flag_stride + m_blk, 1u);
      }
    }
    if (KIND == KIND_F16 && args.overflow) {
      const __half2 mx = *reinterpret_cast<const __half2*>(&sat);
      if (__hge(__hmax(__low2half(mx), __high2half(mx)), __float2half(65504.f))) atomicOr(args.overflow, 1u);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // neither CTA leaves (or frees TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)MLP_TMEM_COLS)
                 : "memory");
  }
  if (n_layers > 1 && cta == 0 && threadIdx.x == 0) {
    // the last cluster to get here resets the dependency counts for the next launch
    unsigned* ticket = args.flags + (size_t)(n_layers - 1) * args.flag_stride;
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ticket) : "memory");
    if (old == (unsigned)n_clusters - 1u) {
      for (int l = 0; l + 1 < n_layers; ++l)
        for (int i = 0; i < sc.n_m; ++i) args.flags[(size_t)l * args.flag_stride + i] = 0u;
      *ticket = 0u;
    }
  }
}

}  // namespace ttl_mlp
