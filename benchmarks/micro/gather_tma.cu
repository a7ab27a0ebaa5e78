// Microbenchmark for north_star (1), "the volume staged through TMA": the 7-point trilinear SH gather of
// the state kernel (environments/env.py:538-541) with the 2x2x2x48 neighbourhood boxes fetched by TMA
// (cp.async.bulk.tensor.4d, one box per neighbourhood point, 7 per streamline) into shared memory,
// against the product's scheme (24 LDG.128 per lane straight into registers, duplicates served by L1).
// Same volume shape (145x174x145x48 fp32), same row count (50 000), tips sorted by voxel like the
// product's alive list; interior tips only (TMA zero-fills out-of-bounds boxes, the reference clamps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_tma gather_tma.cu && ./gather_tma
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

constexpr int X = 145, Y = 174, Z = 145, CP = 48, CP4 = 12;
constexpr int N_ROWS = 50000;
constexpr float RAD = 0.75096f;

struct Corner { int vox[56]; float w[56]; };

__device__ __forceinline__ void corner_table(float tx, float ty, float tz, float* s_w, int* s_vox, int* s_org, int lane) {
  for (int rep = 0; rep < 2; ++rep) {
    const int i = lane + 32 * rep;
    if (i < 56) {
      const int p = i >> 3, c = i & 7;
      float cx = tx, cy = ty, cz = tz;
      if (p == 1) cx += RAD; if (p == 2) cy += RAD; if (p == 3) cz += RAD;
      if (p == 4) cx -= RAD; if (p == 5) cy -= RAD; if (p == 6) cz -= RAD;
      const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
      const float dx = cx - fx, dy = cy - fy, dz = cz - fz;
      const float wx = (c & 4) ? dx : 1.f - dx, wy = (c & 2) ? dy : 1.f - dy, wz = (c & 1) ? dz : 1.f - dz;
      const int xi = (int)fx + ((c >> 2) & 1), yi = (int)fy + ((c >> 1) & 1), zi = (int)fz + (c & 1);
      s_w[i] = wx * wy * wz;
      s_vox[i] = (xi * Y + yi) * Z + zi;
      if (c == 0 && s_org) { s_org[3 * p] = (int)fx; s_org[3 * p + 1] = (int)fy; s_org[3 * p + 2] = (int)fz; }
    }
  }
}

// ---- product scheme: gathers straight into registers ------------------------------------------------
__global__ void __launch_bounds__(256, 6) gather_ldg(const float* __restrict__ vol, const float* __restrict__ tips, int n,
                                                     __nv_bfloat16* __restrict__ out) {
  __shared__ float s_wa[8][64];
  __shared__ int s_va[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= n) return;
  float* s_w = s_wa[warp];
  int* s_vox = s_va[warp];
  corner_table(tips[3 * r], tips[3 * r + 1], tips[3 * r + 2], s_w, s_vox, nullptr, lane);
  __syncwarp();
  const float4* vol4 = reinterpret_cast<const float4*>(vol);
  float4 a[3][8];
  int pp[3], cc[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int item = min(lane + 32 * t, 7 * CP4 - 1);
    const int p = item / CP4, ck = item - p * CP4;
    pp[t] = p; cc[t] = ck;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4* ptr = vol4 + (size_t)s_vox[p * 8 + k] * CP4 + ck;
      asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[t][k].x), "=f"(a[t][k].y), "=f"(a[t][k].z), "=f"(a[t][k].w) : "l"(ptr));
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float w = s_w[pp[t] * 8 + k];
      acc.x = fmaf(w, a[t][k].x, acc.x); acc.y = fmaf(w, a[t][k].y, acc.y);
      acc.z = fmaf(w, a[t][k].z, acc.z); acc.w = fmaf(w, a[t][k].w, acc.w);
    }
    if (lane + 32 * t < 7 * CP4) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
      *reinterpret_cast<uint2*>(out + (size_t)r * 336 + pp[t] * CP + cc[t] * 4) =
          make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
  }
}

// ---- TMA scheme: 7 boxes of 2x2x2 voxels x 48 channels per streamline into shared memory ----------------
constexpr int TMA_WARPS = 4;
constexpr int BOX_FLOATS = 8 * CP;              // 384 floats = 1536 B
constexpr int WARP_SMEM = 7 * BOX_FLOATS * 4;   // 10752 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32 * TMA_WARPS) gather_tma(const __grid_constant__ CUtensorMap map, const float* __restrict__ tips,
                                                             int n, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float s_wa[TMA_WARPS][64];
  __shared__ int s_va[TMA_WARPS][64];
  __shared__ int s_oa[TMA_WARPS][24];
  __shared__ __align__(8) uint64_t bars[TMA_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * TMA_WARPS + warp;
  float* box = reinterpret_cast<float*>(smem + (size_t)warp * WARP_SMEM);
  const uint32_t bar = smem_u32(&bars[warp]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (r >= n) return;
  float* s_w = s_wa[warp];
  int* s_org = s_oa[warp];
  corner_table(tips[3 * r], tips[3 * r + 1], tips[3 * r + 2], s_w, s_va[warp], s_org, lane);
  __syncwarp();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)WARP_SMEM) : "memory");
#pragma unroll
    for (int p = 0; p < 7; ++p) {
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
              smem_u32(box + p * BOX_FLOATS)),
          "l"(reinterpret_cast<uint64_t>(&map)), "r"(bar), "r"(0), "r"(s_org[3 * p + 2]), "r"(s_org[3 * p + 1]), "r"(s_org[3 * p])
          : "memory");
    }
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar) : "memory");
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int item = lane + 32 * t;
    if (item < 7 * CP4) {
      const int p = item / CP4, ck = item - p * CP4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // box layout [x2][y2][z2][48]: corner k = 4 cx + 2 cy + cz is voxel k of the box
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(box + p * BOX_FLOATS + k * CP + ck * 4);
        const float w = s_w[p * 8 + k];
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
      __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y), h1 = __floats2bfloat162_rn(acc.z, acc.w);
      *reinterpret_cast<uint2*>(out + (size_t)r * 336 + p * CP + ck * 4) =
          make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int main() {
  const size_t n_vox = (size_t)X * Y * Z;
  std::vector<float> h_vol(n_vox * CP);
  uint32_t s = 12345u;
  for (size_t i = 0; i < h_vol.size(); ++i) { s = s * 1664525u + 1013904223u; h_vol[i] = ((i % CP) < 45) ? (float)(s >> 8) / 16777216.f - 0.5f : 0.f; }
  // tips: random interior points, then sorted by voxel raster key (the product's alive list is in that order)
  struct Tip { float x, y, z; long key; };
  std::vector<Tip> tips(N_ROWS);
  for (auto& t : tips) {
    s = s * 1664525u + 1013904223u; t.x = 2.f + (float)(s >> 8) / 16777216.f * (X - 5);
    s = s * 1664525u + 1013904223u; t.y = 2.f + (float)(s >> 8) / 16777216.f * (Y - 5);
    s = s * 1664525u + 1013904223u; t.z = 2.f + (float)(s >> 8) / 16777216.f * (Z - 5);
    t.key = ((long)t.x * Y + (long)t.y) * Z + (long)t.z;
  }
  std::sort(tips.begin(), tips.end(), [](const Tip& a, const Tip& b) { return a.key < b.key; });
  std::vector<float> h_tips(3 * N_ROWS);
  for (int i = 0; i < N_ROWS; ++i) { h_tips[3 * i] = tips[i].x; h_tips[3 * i + 1] = tips[i].y; h_tips[3 * i + 2] = tips[i].z; }
  float *d_vol, *d_tips;
  __nv_bfloat16 *d_a, *d_b;
  CK(cudaMalloc(&d_vol, h_vol.size() * 4));
  CK(cudaMalloc(&d_tips, h_tips.size() * 4));
  CK(cudaMalloc(&d_a, (size_t)N_ROWS * 336 * 2));
  CK(cudaMalloc(&d_b, (size_t)N_ROWS * 336 * 2));
  CK(cudaMemcpy(d_vol, h_vol.data(), h_vol.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tips, h_tips.data(), h_tips.size() * 4, cudaMemcpyHostToDevice));
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  CUtensorMap map;
  cuuint64_t dims[4] = {CP, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X};
  cuuint64_t strides[3] = {CP * 4, (cuuint64_t)Z * CP * 4, (cuuint64_t)Y * Z * CP * 4};
  cuuint32_t boxd[4] = {CP, 2, 2, 2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_vol, dims, strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed\n");
    return 1;
  }
  CK(cudaFuncSetAttribute(gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_WARPS * WARP_SMEM));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms_ldg = 0, ms_tma = 0;
  const int reps = 20;
  for (int it = 0; it < 3 + reps; ++it) {
    if (it == 3) cudaEventRecord(e0);
    gather_ldg<<<(N_ROWS + 7) / 8, 256>>>(d_vol, d_tips, N_ROWS, d_a);
  }
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_ldg, e0, e1);
  for (int it = 0; it < 3 + reps; ++it) {
    if (it == 3) cudaEventRecord(e0);
    gather_tma<<<(N_ROWS + TMA_WARPS - 1) / TMA_WARPS, 32 * TMA_WARPS, TMA_WARPS * WARP_SMEM>>>(map, d_tips, N_ROWS, d_b);
  }
  cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms_tma, e0, e1);
  CK(cudaGetLastError());
  std::vector<uint16_t> ha((size_t)N_ROWS * 336), hb((size_t)N_ROWS * 336);
  CK(cudaMemcpy(ha.data(), d_a, ha.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hb.data(), d_b, hb.size() * 2, cudaMemcpyDeviceToHost));
  size_t diff = 0;
  for (size_t i = 0; i < ha.size(); ++i) diff += ha[i] != hb[i];
  int occ_ldg = 0, occ_tma = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ldg, gather_ldg, 256, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_tma, gather_tma, 32 * TMA_WARPS, TMA_WARPS * WARP_SMEM);
  printf("{\"rows\": %d, \"ldg128_us\": %.2f, \"tma_box_us\": %.2f, \"ldg_warps_per_sm\": %d, \"tma_warps_per_sm\": %d, "
         "\"tma_smem_bytes_per_row\": %d, \"mismatching_bf16_values\": %zu}\n",
         N_ROWS, 1000.f * ms_ldg / reps, 1000.f * ms_tma / reps, occ_ldg * 8, occ_tma * TMA_WARPS, WARP_SMEM, diff);
  return 0;
}
