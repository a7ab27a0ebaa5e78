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

__device__ __forceinline__ uint32_t ttl_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
