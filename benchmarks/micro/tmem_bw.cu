// Microbenchmark: tensor-memory read / write bandwidth per SM on B200 (tcgen05.ld / tcgen05.st,
// shape 32x32b.x32), 1 or 2 CTAs per SM, 4 warps per CTA (one per TMEM lane quarter).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t v[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

template <int MODE>   // 0: loads, 1: stores, 2: load + store pairs
__global__ void __launch_bounds__(128) tmem_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
  uint32_t v[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  st32(base, v);
  st32(base + 32, v);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
      uint32_t w[32];
      ld32(base + (it & 1) * 32, v);
      ld32(base + 128 + (it & 1) * 32, w);      // two loads in flight
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[0] ^ v[31] ^ w[0] ^ w[31];
    }
    if (MODE == 1 || MODE == 2) {
      st32(base + 64 + (it & 1) * 32, v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * 128 + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256u) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* cyc;
  uint32_t* sink;
  cudaMalloc(&cyc, 8 * 1024);
  cudaMalloc(&sink, 4 * 128 * 1024);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode)
    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
      const int grid = sms * per_sm;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_kernel<0><<<grid, 128>>>(iters, cyc, sink);
        if (mode == 1) tmem_kernel<1><<<grid, 128>>>(iters, cyc, sink);
        if (mode == 2) tmem_kernel<2><<<grid, 128>>>(iters, cyc, sink);
        cudaDeviceSynchronize();
      }
      unsigned long long h[1024];
      cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < grid; ++i) avg += (double)h[i];
      avg /= grid;
      // bytes moved per CTA per iteration: 16 KB per tcgen05.ld/st.x32 over 128 lanes; ld modes issue two loads
      const double bytes = 16384.0 * iters * (mode == 0 ? 2 : mode == 2 ? 3 : 1) * per_sm;
      printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"cycles_per_iter\": %.1f, \"bytes_per_cycle_per_sm\": %.1f, \"err\": \"%s\"}\n",
             mode == 0 ? "ld" : mode == 1 ? "st" : "ld+st", per_sm, avg / iters, bytes / avg,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
