// Microbenchmark: cycles per small tcgen05.mma (M=128, K=16, fp16) on B200 when consecutive MMAs
// accumulate into the SAME TMEM columns (dependent chain) or rotate over several accumulators,
// with the A operand in shared memory (SS) or tensor memory (TS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_latency umma_latency.cu && ./umma_latency
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

template <int N, int NACC, int TS>
__global__ void __launch_bounds__(64) k(int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t bb = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint64_t da = desc_sw128(sb), db = desc_sw128(sb + 16384);
    const uint32_t id = idesc(128, N);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tm + 256 + (uint32_t)((it % NACC) * N);
        if (TS) mma_ts(d, tm + (uint32_t)((it & 3) * 8), db + 2 * (it & 3), id, 1u);
        else mma_ss(d, da + 2 * (it & 3), db + 2 * (it & 3), id, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bb) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bb), "r"(0u) : "memory");
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int N, int NACC, int TS>
void run(const char* name, unsigned long long* dout, int grid) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<N, NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792);
  for (int r = 0; r < 2; ++r) {
    k<N, NACC, TS><<<grid, 64, 33792>>>(iters, dout);
    cudaDeviceSynchronize();
  }
  unsigned long long h[296];
  cudaMemcpy(h, dout, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += (double)h[i];
  printf("{\"case\": \"%s\", \"N\": %d, \"accumulators\": %d, \"a_operand\": \"%s\", \"cycles_per_mma\": %.1f, \"err\": \"%s\"}\n", name, N,
         NACC, TS ? "tmem" : "smem", avg / grid / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 8 * 296);
  run<32, 1, 0>("dependent", d, 148);
  run<32, 4, 0>("rotating", d, 148);
  run<32, 1, 1>("dependent", d, 148);
  run<32, 4, 1>("rotating", d, 148);
  run<64, 1, 0>("dependent", d, 148);
  run<64, 2, 0>("rotating", d, 148);
  run<16, 1, 1>("dependent", d, 148);
  run<16, 4, 1>("rotating", d, 148);
  run<256, 1, 0>("dependent", d, 148);
  return 0;
}
