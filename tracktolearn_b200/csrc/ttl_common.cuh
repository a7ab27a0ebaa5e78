// Shared host/device helpers for libttl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "ttl_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libttl_b200 targets sm_100a only"
#endif

extern std::atomic<long long> g_ttl_launches;

#define TTL_LAUNCHED() (g_ttl_launches.fetch_add(1, std::memory_order_relaxed))

// Optional per-kernel timing with CUDA events on the launching stream (ttl_prof_enable).
void ttl_prof_begin(const char* name, cudaStream_t s);
void ttl_prof_end(cudaStream_t s);

// Every kernel launch of the library goes through this macro: it counts the launch and, when
// profiling is on, brackets it with a pair of events.
#define TTL_LAUNCH(name, stream, ...)    \
  do {                                   \
    ttl_prof_begin(name, stream);        \
    __VA_ARGS__;                         \
    ttl_prof_end(stream);                \
    TTL_LAUNCHED();                      \
  } while (0)

#define TTL_CHECK_LAST()                     \
  do {                                       \
    cudaError_t e__ = cudaGetLastError();    \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

static inline int ttl_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL) -------------------------------------------------
// The kernels of one tracking step form a chain on one stream.  Launched with the programmatic
// stream-serialization attribute, kernel N+1 is set up (and its CTAs become resident as kernel N's
// drain) while kernel N is still running; it blocks in ttl_grid_dep_wait() until kernel N has
// completed and its writes are visible.  Every kernel of the chain waits before its first global
// access, so by induction all earlier kernels are complete when a wait returns; dependents are
// released at exit, except by the state kernel (see build_state_kernel).  Both instructions are
// no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void ttl_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ttl_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 1 (default): step kernels are launched with the attribute; ttl_pdl_enable(0) turns it off.
extern std::atomic<int> g_ttl_pdl;

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t ttl_launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                           cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_ttl_pdl.load(std::memory_order_relaxed) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- operand conversions shared by the actor kernels and the env's state kernel --------------------
// fp32 -> tf32 (10-bit mantissa) with round-to-nearest, returned as the fp32 bit pattern
__device__ __forceinline__ uint32_t ttl_round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// two fp32 -> packed fp16 pair (lo = a, hi = b), saturating to +-65504 instead of overflowing to inf
__device__ __forceinline__ uint32_t ttl_pack_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// two fp32 -> packed bf16 pair (lo = a, hi = b)
__device__ __forceinline__ uint32_t ttl_pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// fmt: TTL_OPERAND_BF16 or TTL_OPERAND_FP16 (run-time)
__device__ __forceinline__ uint32_t ttl_pack16(float a, float b, int fmt) {
  return fmt == TTL_OPERAND_FP16 ? ttl_pack_f16x2_sat(a, b) : ttl_pack_bf16x2(a, b);
}

// Sum of the actor's fused-head partials for one output.  The producing kernel (ttl_mlp.cuh) forms
// a tile's partial as a fixed tree over 64-column groups -- p_g = fma chain over group g, tile =
// (p0 + p1) + (p2 + p3) over the groups it covers -- so that the total does not depend on the tile
// width of the launch: the consumer rebuilds the tree per 256-column super-tile and adds the
// super-tiles left to right.  p: this row's partials [n_tiles][8]; tiles_per_256 = 256 / tile width.
__device__ __forceinline__ float ttl_head_tree_sum(const float* __restrict__ p, int n_tiles, int tiles_per_256,
                                                  int o) {
  float acc = 0.f;
  for (int t = 0; t < n_tiles; t += tiles_per_256) {
    float s;
    if (tiles_per_256 == 1) {
      s = __ldg(p + (size_t)t * 8 + o);
    } else if (tiles_per_256 == 2) {
      const float a = __ldg(p + (size_t)t * 8 + o);
      const float b = t + 1 < n_tiles ? __ldg(p + (size_t)(t + 1) * 8 + o) : 0.f;
      s = a + b;
    } else {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = t + j < n_tiles ? __ldg(p + (size_t)(t + j) * 8 + o) : 0.f;
      s = (v[0] + v[1]) + (v[2] + v[3]);
    }
    acc += s;
  }
  return acc;
}

// The first three outputs (the action means) at once, from float4 loads issued together: what the env
// step's propagate kernel needs.  Same additions in the same order as ttl_head_tree_sum.
__device__ __forceinline__ void ttl_head_tree_sum3(const float* __restrict__ p, int n_tiles, int tiles_per_256,
                                                   float& ox, float& oy, float& oz) {
  const float4* q = reinterpret_cast<const float4*>(p);      // tile t: q[2 t] = outputs 0..3
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (tiles_per_256 == 1 && n_tiles <= 4) {                  // the full-batch shape: up to 4 tiles of 256 columns
    float4 v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = t < n_tiles ? __ldg(q + 2 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t < n_tiles) { ax += v[t].x; ay += v[t].y; az += v[t].z; }
  } else {
    for (int t = 0; t < n_tiles; t += tiles_per_256) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = (j < tiles_per_256 && t + j < n_tiles) ? __ldg(q + 2 * (t + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (tiles_per_256 == 1) {
        ax += v[0].x; ay += v[0].y; az += v[0].z;
      } else if (tiles_per_256 == 2) {
        ax += v[0].x + v[1].x; ay += v[0].y + v[1].y; az += v[0].z + v[1].z;
      } else {
        ax += (v[0].x + v[1].x) + (v[2].x + v[3].x);
        ay += (v[0].y + v[1].y) + (v[2].y + v[3].y);
        az += (v[0].z + v[1].z) + (v[2].z + v[3].z);
      }
    }
  }
  ox = ax; oy = ay; oz = az;
}

__device__ __forceinline__ uint32_t ttl_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
