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

__device__ __forceinline__ uint32_t ttl_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
