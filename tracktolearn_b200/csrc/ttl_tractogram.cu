// Tractogram output path on the device: what the reference does
// per streamline in Python after tracking (tracking/tracker.py:118-125) -- dipy `length` for the
// min/max length filter and dipy `compress_streamlines` for --compress -- on packed ragged arrays
// [total][3] fp32 + int64 offsets [n+1], one warp per streamline.
#include <math.h>

#include "ttl_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

// x*x + y*y + z*z left to right without fused multiply-adds (numpy's order and rounding)
__device__ __forceinline__ double sumsq3(double x, double y, double z) {
  return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

// dipy.tracking.streamline.length: sum of the segment norms, in double (tracker.py:120).
// Lanes take segments round-robin; the 32 partial sums are combined by a shuffle tree, so the value
// can differ from a sequential sum in the last bits (tests allow 1e-12 relative).
__global__ void __launch_bounds__(kWarpsPerBlock * 32) streamline_length_kernel(
    const float* __restrict__ points, const long long* __restrict__ offsets, int n, double* __restrict__ out) {
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const long long beg = offsets[i], end = offsets[i + 1];
  const float* P = points + beg * 3;
  const int N = (int)(end - beg);
  double acc = 0.0;
  for (int j = lane; j + 1 < N; j += 32) {
    const double dx = (double)P[3 * j + 3] - (double)P[3 * j + 0];
    const double dy = (double)P[3 * j + 4] - (double)P[3 * j + 1];
    const double dz = (double)P[3 * j + 5] - (double)P[3 * j + 2];
    acc += sqrt(dx * dx + dy * dy + dz * dz);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) out[i] = acc;
}

// dipy compress_streamlines (streamlinespeed.pyx c_compress_streamline), restated in
// oracle/ttl_oracle.py::compress_streamline -- same arithmetic: coordinate differences and their
// products in float (no contraction), sums and square roots in double.  The walk over `nxt` is
// sequential (prev depends on the previous decision); the test of the points between prev and nxt
// is spread over the lanes (any point off the chord by more than tol, or a NaN distance, keeps
// point nxt-1).  Output: keep[total] (1 = point survives) and count[n].
__global__ void __launch_bounds__(kWarpsPerBlock * 32) compress_mask_kernel(
    const float* __restrict__ points, const long long* __restrict__ offsets, int n, double tol_error,
    double max_segment_length, uint8_t* __restrict__ keep, int* __restrict__ count) {
  const int i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const long long beg = offsets[i], end = offsets[i + 1];
  const float* P = points + beg * 3;
  uint8_t* K = keep + beg;
  const int N = (int)(end - beg);
  if (N <= 2) {   // copied as is
    if (lane < N) K[lane] = 1;
    if (lane == 0) count[i] = N;
    return;
  }
  for (int j = lane; j < N; j += 32) K[j] = (j == 0 || j == N - 1) ? 1 : 0;
  __syncwarp();   // lane 0 overwrites entries other lanes have just cleared
  int prev = 0, nb = 2;
  float px = P[0], py = P[1], pz = P[2];
  for (int nxt = 2; nxt < N; ++nxt) {
    const float nx = P[3 * nxt + 0], ny = P[3 * nxt + 1], nz = P[3 * nxt + 2];
    const float ax = __fsub_rn(nx, px), ay = __fsub_rn(ny, py), az = __fsub_rn(nz, pz);
    const double seg = sqrt(sumsq3((double)ax, (double)ay, (double)az));
    bool replace = false;
    if (seg < max_segment_length) {
      const double norm2 = sqrt(__dadd_rn(__dadd_rn((double)__fmul_rn(ax, ax), (double)__fmul_rn(ay, ay)),
                                          (double)__fmul_rn(az, az)));
      bool fail = false;
      for (int base = prev + 1; base < nxt; base += 32) {
        const int curr = base + lane;
        if (curr < nxt) {
          const float bx = __fsub_rn(P[3 * curr + 0], nx), by = __fsub_rn(P[3 * curr + 1], ny),
                      bz = __fsub_rn(P[3 * curr + 2], nz);
          const double cx = (double)__fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
          const double cy = (double)__fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
          const double cz = (double)__fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
          const double dist = sqrt(sumsq3(cx, cy, cz)) / norm2;
          fail |= isnan(dist) || dist > tol_error;
        }
      }
      replace = !__any_sync(0xffffffffu, fail);
    }
    if (!replace) {
      prev = nxt - 1;
      px = P[3 * prev + 0]; py = P[3 * prev + 1]; pz = P[3 * prev + 2];
      if (lane == 0) K[prev] = 1;
      ++nb;
    }
  }
  if (lane == 0) count[i] = nb;
}

}  // namespace

extern "C" {

int ttl_streamline_lengths(const float* points, const int64_t* offsets, int32_t n, double* out_lengths,
                           void* stream) {
  if (n <= 0) return 0;
  if (!points || !offsets || !out_lengths) return TTL_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  TTL_LAUNCH("streamline_length_kernel", s,
             streamline_length_kernel<<<ttl_div_up(n, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                 points, (const long long*)offsets, n, out_lengths));
  TTL_CHECK_LAST();
  return 0;
}

int ttl_compress_mask(const float* points, const int64_t* offsets, int32_t n, double tol_error,
                      double max_segment_length, uint8_t* keep, int32_t* count, void* stream) {
  if (n <= 0) return 0;
  if (!points || !offsets || !keep || !count || !(tol_error >= 0.0)) return TTL_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  TTL_LAUNCH("compress_mask_kernel", s,
             compress_mask_kernel<<<ttl_div_up(n, kWarpsPerBlock), kWarpsPerBlock * 32, 0, s>>>(
                 points, (const long long*)offsets, n, tol_error, max_segment_length, keep, count));
  TTL_CHECK_LAST();
  return 0;
}

}  // extern "C"
