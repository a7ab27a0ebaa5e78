// Load-time SH work on the device: the peak extraction the
// reference runs voxel by voxel in Python when a subject is loaded from files
// (environments/env.py:405-432): spherical function on a hemisphere (SF = sh . B^T, the only SH-to-SF
// projection in the reference), scilpy get_maximas -> dipy peak_directions (local maxima over the
// sphere's edges, relative threshold, 25 degree separation), the first five peaks scaled by
// value / first value.  One warp per voxel; the spherical function lives in shared memory in double
// like the reference's (float32 coefficients times a float64 matrix).
#include <math.h>

#include "ttl_common.cuh"

namespace {

constexpr int kPeakWarps = 8;
constexpr int kMaxVerts = 512;      // doubles of shared memory per warp
constexpr int kMaxCoefs = 64;
constexpr int kMaxPeaks = 8;

__global__ void __launch_bounds__(kPeakWarps * 32) peaks_from_sh_kernel(
    const float* __restrict__ sh, long long n_vox, int C, int ld,
    const double* __restrict__ B,          // [V][C]
    const double* __restrict__ verts,      // [V][3]
    const int* __restrict__ nbr, int V, int D,
    double rel_thr, double abs_thr, double sep_cos, int npeaks, float* __restrict__ out) {
  __shared__ double s_sf[kPeakWarps][kMaxVerts];
  __shared__ float s_sh[kPeakWarps][kMaxCoefs];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long vox = (long long)blockIdx.x * kPeakWarps + warp;
  if (vox >= n_vox) return;
  double* sf = s_sf[warp];
  float* c = s_sh[warp];
  float* o = out + vox * (npeaks * 3);
  for (int j = lane; j < npeaks * 3; j += 32) o[j] = 0.f;
  // np.sum(data, axis=-1) != 0 (env.py:416), float32 accumulation
  for (int j = lane; j < C; j += 32) c[j] = sh[vox * ld + j];
  __syncwarp();
  float total = 0.f;
  for (int j = 0; j < C; ++j) total = __fadd_rn(total, c[j]);
  if (total == 0.f) return;       // warp-uniform

  // SF on the sphere, values under the absolute threshold zeroed (scilpy get_maximas)
  double vmin = INFINITY;
  bool has_nan = false;
  for (int v = lane; v < V; v += 32) {
    const double* b = B + (size_t)v * C;
    double acc = 0.0;
    for (int j = 0; j < C; ++j) acc = fma((double)c[j], __ldg(b + j), acc);
    if (acc < abs_thr) acc = 0.0;
    has_nan |= acc != acc;
    sf[v] = acc;
    vmin = fmin(vmin, acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, off));
  if (__any_sync(0xffffffffu, has_nan)) return;    // dipy raises on NaN; no peaks here
  __syncwarp();
  const double odf_min = fmax(vmin, 0.0);

  // local maxima (dipy local_maxima): greater than some neighbour, smaller than none.
  // Each lane keeps a bit mask of its own vertices (vertex v = lane + 32 * t) that are candidates.
  unsigned cand = 0;
  for (int v = lane, t = 0; v < V; v += 32, ++t) {
    const double x = sf[v];
    bool greater = false, less = false;
    for (int k = 0; k < D; ++k) {
      const int u = __ldg(nbr + v * D + k);
      if (u < 0) continue;
      const double y = sf[u];
      greater |= x > y;
      less |= x < y;
    }
    if (greater && !less) cand |= 1u << t;
  }

  // Peaks in descending order (ties: lower index first); stop at the relative threshold; keep a
  // direction only if it is further than the separation angle from every kept one (|cos|).
  double kept[kMaxPeaks][3];
  double first_val = 0.0, first_norm = 0.0;
  int n_kept = 0;
  bool first = true;
  while (n_kept < npeaks) {
    double best = -INFINITY;
    int best_v = 0x7fffffff;
    for (unsigned m = cand; m; m &= m - 1) {
      const int t = __ffs(m) - 1, v = lane + 32 * t;
      const double x = sf[v];
      if (x > best || (x == best && v < best_v)) { best = x; best_v = v; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int ov = __shfl_xor_sync(0xffffffffu, best_v, off);
      if (ob > best || (ob == best && ov < best_v)) { best = ob; best_v = ov; }
    }
    if (best_v == 0x7fffffff) break;            // no candidates left
    if (first && best < 0.0) break;             // peak_directions: values[0] < 0 -> nothing
    const double norm = best - odf_min;
    if (first) { first_norm = norm; first_val = best; }
    else if (!(norm >= rel_thr * first_norm)) break;
    first = false;
    if ((best_v & 31) == lane) cand &= ~(1u << (best_v >> 5));
    const double dx = __ldg(verts + 3 * best_v), dy = __ldg(verts + 3 * best_v + 1), dz = __ldg(verts + 3 * best_v + 2);
    bool similar = false;
    for (int k = 0; k < n_kept; ++k)
      similar |= fabs(dx * kept[k][0] + dy * kept[k][1] + dz * kept[k][2]) > sep_cos;
    if (similar) continue;
    kept[n_kept][0] = dx; kept[n_kept][1] = dy; kept[n_kept][2] = dz;
    // peak_values / peak_values[0] (0 where the first is 0), direction scaled by it (env.py:427-431)
    const double w = first_val != 0.0 ? best / first_val : 0.0;
    if (lane == 0) {
      o[3 * n_kept + 0] = (float)(dx * w);
      o[3 * n_kept + 1] = (float)(dy * w);
      o[3 * n_kept + 2] = (float)(dz * w);
    }
    ++n_kept;
  }
}

}  // namespace

extern "C" {

int ttl_peaks_from_sh(const float* sh, int64_t n_voxels, int32_t C, int32_t ld, const double* basis,
                      const double* vertices, const int32_t* neighbours, int32_t n_vertices,
                      int32_t max_degree, double relative_threshold, double absolute_threshold,
                      double min_separation_deg, int32_t npeaks, float* out_peaks, void* stream) {
  if (!sh || !basis || !vertices || !neighbours || !out_peaks) return TTL_ERR_BAD_ARG;
  if (C <= 0 || C > kMaxCoefs || ld < C || n_vertices <= 0 || n_vertices > kMaxVerts || max_degree <= 0 ||
      npeaks <= 0 || npeaks > kMaxPeaks)
    return TTL_ERR_UNSUPPORTED;
  if (n_voxels <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const double sep_cos = cos(min_separation_deg * 3.14159265358979323846 / 180.0);
  TTL_LAUNCH("peaks_from_sh_kernel", s,
             peaks_from_sh_kernel<<<ttl_div_up(n_voxels, kPeakWarps), kPeakWarps * 32, 0, s>>>(
                 sh, n_voxels, C, ld, basis, vertices, neighbours, n_vertices, max_degree, relative_threshold,
                 absolute_threshold, sep_cos, npeaks, out_peaks));
  TTL_CHECK_LAST();
  return 0;
}

}  // extern "C"
