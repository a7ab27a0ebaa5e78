// TractOracle-Net scoring on the device.
//
// Reference: oracles/oracle.py:39-89 (OracleSingleton.predict: dipy set_number_of_points(s,128),
// np.diff, batches of 4096 through the model under fp16 autocast) and
// oracles/transformer_oracle.py:40-92 (CLS token + Linear(3,32)+ReLU scaled by sqrt(32), sinusoidal
// positional encoding, n_layers post-norm nn.TransformerEncoderLayer(d=32, n_head, ff=2048, relu),
// sigmoid(Linear(32,1)) of token 0).
//
//   oracle_features_kernel  one warp per streamline, bit-identical to the sequential algorithm of dipy's
//                           c_set_number_of_points (segment differences in float, arc lengths and
//                           interpolation in double) fused with the np.diff: no host resampling.
//   oracle_forward_kernel   one CTA per streamline, one thread per token (128).  The residual
//                           stream lives in registers, K/V of the current layer in shared memory,
//                           weights are streamed through shared memory in 16 KB chunks; all layers
//                           in one launch, fp32 throughout (the reference's CPU precision; its CUDA
//                           path autocasts to fp16).  ~147 MFLOP per streamline on the FP32 pipes.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstddef>
#include <cstdlib>
#include <new>

#include "ttl_common.cuh"
#include "ttl_tc.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// features
// ------------------------------------------------------------------------------------------
constexpr int kOraclePts = 128;

// One streamline: P = its N points, D = its 127 output directions.
__device__ void resample_and_diff(const float* __restrict__ P, int N, float* __restrict__ D) {
  if (N <= 0) {
    for (int j = 0; j < (kOraclePts - 1) * 3; ++j) D[j] = 0.f;
    return;
  }
  // pass 1: total arc length, summed sequentially in double like c_arclengths
  double total = 0.0;
  for (int i = 1; i < N; ++i) {
    const double dx = (double)__fsub_rn(P[3 * i], P[3 * i - 3]);
    const double dy = (double)__fsub_rn(P[3 * i + 1], P[3 * i - 2]);
    const double dz = (double)__fsub_rn(P[3 * i + 2], P[3 * i - 1]);
    total = __dadd_rn(total, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))));
  }
  const double step = total / (double)(kOraclePts - 1);
  // pass 2: the while loop of c_set_number_of_points; emit res[i] - res[i-1] as points appear
  double nxt = 0.0, cum_k = 0.0, cum_km1 = 0.0;   // cum[k], cum[k-1]
  int i = 0, k = 0;
  float prev[3] = {0.f, 0.f, 0.f};                // res[i-1]
  float res126[3] = {0.f, 0.f, 0.f};
  auto emit = [&](float x, float y, float z) {
    if (i >= 1 && i <= kOraclePts - 2) {
      D[3 * (i - 1) + 0] = __fsub_rn(x, prev[0]);
      D[3 * (i - 1) + 1] = __fsub_rn(y, prev[1]);
      D[3 * (i - 1) + 2] = __fsub_rn(z, prev[2]);
    }
    if (i == kOraclePts - 2) { res126[0] = x; res126[1] = y; res126[2] = z; }
    prev[0] = x; prev[1] = y; prev[2] = z;
    ++i;
  };
  while (nxt < total && i < kOraclePts) {
    if (nxt == cum_k) {
      emit(P[3 * k], P[3 * k + 1], P[3 * k + 2]);
      nxt += step;
      ++k;
      if (k < N) {
        const double dx = (double)__fsub_rn(P[3 * k], P[3 * k - 3]);
        const double dy = (double)__fsub_rn(P[3 * k + 1], P[3 * k - 2]);
        const double dz = (double)__fsub_rn(P[3 * k + 2], P[3 * k - 1]);
        cum_km1 = cum_k;
        cum_k = __dadd_rn(cum_k, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))));
      }
    } else if (nxt < cum_k) {
      const double ratio = 1.0 - ((cum_k - nxt) / (cum_k - cum_km1));
      float r[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double delta = (double)__fsub_rn(P[3 * k + d], P[3 * (k - 1) + d]);
        r[d] = (float)__dadd_rn((double)P[3 * (k - 1) + d], __dmul_rn(ratio, delta));
      }
      emit(r[0], r[1], r[2]);
      nxt += step;
    } else {
      ++k;
      if (k >= N) break;
      const double dx = (double)__fsub_rn(P[3 * k], P[3 * k - 3]);
      const double dy = (double)__fsub_rn(P[3 * k + 1], P[3 * k - 2]);
      const double dz = (double)__fsub_rn(P[3 * k + 2], P[3 * k - 1]);
      cum_km1 = cum_k;
      cum_k = __dadd_rn(cum_k, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz))));
    }
  }
  // points the loop never produced stay zero (the restated np.zeros initialisation)
  while (i <= kOraclePts - 2) emit(0.f, 0.f, 0.f);
  // the last resampled point is always the original last point
  D[3 * (kOraclePts - 2) + 0] = __fsub_rn(P[3 * (N - 1) + 0], res126[0]);
  D[3 * (kOraclePts - 2) + 1] = __fsub_rn(P[3 * (N - 1) + 1], res126[1]);
  D[3 * (kOraclePts - 2) + 2] = __fsub_rn(P[3 * (N - 1) + 2], res126[2]);
}

// Warp-per-streamline version of the same algorithm, bit-identical to resample_and_diff():
//   1. lanes compute the segment lengths (double) in parallel;
//   2. lane 0 accumulates them in the reference's order (the sums must round identically) and then
//      builds the 128 targets nxt_i by repeated addition of total/127, as the reference loop does;
//   3. every lane resolves 4 of the 128 output points: the loop's state machine reduces to "smallest
//      k with cum[k] >= nxt_i" (emit P[k] when equal, else interpolate in segment k-1..k), found
//      by binary search; targets >= total are never reached by the loop and stay zero;
//   4. differences of consecutive points, the last point pinned to the original last point.
// Streamlines longer than FEAT_CAP points take the sequential path on lane 0.
constexpr int FEAT_CAP = 1024;
constexpr int FEAT_WARPS = 4;
struct FeatSmem {
  double cum[FEAT_CAP];          // cum[0] = 0, cum[k] = arc length up to point k
  double nxt[kOraclePts];
  float res[kOraclePts][3];
};

__device__ __forceinline__ double seg_len(const float* __restrict__ P, int i) {
  const double dx = (double)__fsub_rn(P[3 * i], P[3 * i - 3]);
  const double dy = (double)__fsub_rn(P[3 * i + 1], P[3 * i - 2]);
  const double dz = (double)__fsub_rn(P[3 * i + 2], P[3 * i - 1]);
  return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__device__ void resample_and_diff_warp(const float* __restrict__ P, int N, float* __restrict__ D, FeatSmem& sm, int lane) {
  if (N <= 0 || N > FEAT_CAP) {
    if (lane == 0) resample_and_diff(P, N, D);
    return;
  }
  for (int i = 1 + lane; i < N; i += 32) sm.cum[i] = seg_len(P, i);
  __syncwarp();
  if (lane == 0) {
    double c = 0.0;
    sm.cum[0] = 0.0;
    for (int i = 1; i < N; ++i) {
      c = __dadd_rn(c, sm.cum[i]);
      sm.cum[i] = c;
    }
    const double step = c / (double)(kOraclePts - 1);
    double t = 0.0;
    for (int i = 0; i < kOraclePts; ++i) {
      sm.nxt[i] = t;
      t += step;
    }
  }
  __syncwarp();
  const double total = sm.cum[N - 1];
#pragma unroll
  for (int q = 0; q < kOraclePts / 32; ++q) {
    const int i = lane + 32 * q;
    const double t = sm.nxt[i];
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    if (i == kOraclePts - 1) {
      r0 = P[3 * (N - 1)]; r1 = P[3 * (N - 1) + 1]; r2 = P[3 * (N - 1) + 2];
    } else if (t < total) {
      int lo = 0, hi = N - 1;              // cum[hi] = total > t: the answer exists
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sm.cum[mid] >= t) hi = mid; else lo = mid + 1;
      }
      const int k = lo;
      const double ck = sm.cum[k];
      if (t == ck) {
        r0 = P[3 * k]; r1 = P[3 * k + 1]; r2 = P[3 * k + 2];
      } else {
        const double ratio = 1.0 - ((ck - t) / (ck - sm.cum[k - 1]));
        const float* a = P + 3 * (k - 1);
        r0 = (float)__dadd_rn((double)a[0], __dmul_rn(ratio, (double)__fsub_rn(a[3], a[0])));
        r1 = (float)__dadd_rn((double)a[1], __dmul_rn(ratio, (double)__fsub_rn(a[4], a[1])));
        r2 = (float)__dadd_rn((double)a[2], __dmul_rn(ratio, (double)__fsub_rn(a[5], a[2])));
      }
    }
    sm.res[i][0] = r0; sm.res[i][1] = r1; sm.res[i][2] = r2;
  }
  __syncwarp();
  for (int j = lane; j < (kOraclePts - 1) * 3; j += 32) {
    const int i = j / 3, d = j - 3 * i;
    D[j] = __fsub_rn(sm.res[i + 1][d], sm.res[i][d]);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * FEAT_WARPS) oracle_features_kernel(const float* __restrict__ points,
                                                                          const long long* __restrict__ offsets, int n,
                                                                          float* __restrict__ dirs) {
  extern __shared__ __align__(16) uint8_t feat_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FeatSmem& sm = reinterpret_cast<FeatSmem*>(feat_raw)[warp];
  for (int s = blockIdx.x * FEAT_WARPS + warp; s < n; s += gridDim.x * FEAT_WARPS) {
    const long long o0 = offsets[s];
    resample_and_diff_warp(points + o0 * 3, (int)(offsets[s + 1] - o0), dirs + (size_t)s * (kOraclePts - 1) * 3, sm, lane);
  }
}

// The alive streamlines of a tracking batch, straight from the streamline buffer (what
// OracleStoppingCriterion / OracleReward score every step, stopping_criteria.py:113-154).
__global__ void __launch_bounds__(32 * FEAT_WARPS) oracle_features_rows_kernel(ttl_batch b, int cur, float* __restrict__ dirs) {
  extern __shared__ __align__(16) uint8_t feat_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FeatSmem& sm = reinterpret_cast<FeatSmem*>(feat_raw)[warp];
  const int n = b.ctrl[cur];
  for (int r = blockIdx.x * FEAT_WARPS + warp; r < n; r += gridDim.x * FEAT_WARPS) {
    const int row = b.alive[cur][r];
    resample_and_diff_warp(b.points + (size_t)row * b.max_pts * 3, b.npts[row], dirs + (size_t)r * (kOraclePts - 1) * 3, sm, lane);
  }
}

// ------------------------------------------------------------------------------------------
// transformer
// ------------------------------------------------------------------------------------------
constexpr int D_MODEL = 32;
constexpr int N_TOK = 128;
constexpr int FF_CHUNK = 64;

struct OracleSmem {
  float k[N_TOK][D_MODEL];        // 16 KB keys of the current layer
  float v[N_TOK][D_MODEL];        // 16 KB values
  float w[4096];                  // 16 KB weight staging
  float b[256];                   // bias staging
};

__device__ __forceinline__ void layer_norm32(float x[D_MODEL], const float* __restrict__ g,
                                             const float* __restrict__ b) {
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) mean += x[i];
  mean *= (1.f / D_MODEL);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) { const float d = x[i] - mean; var = fmaf(d, d, var); }
  var *= (1.f / D_MODEL);
  const float inv = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) x[i] = (x[i] - mean) * inv * __ldg(g + i) + __ldg(b + i);
}

// y[i] = sum_k W[i][k] x[k] + bias[i] for i < ROWS, W staged in shared memory as [ROWS][32]
template <int ROWS>
__device__ __forceinline__ void matvec32(const float* __restrict__ sw, const float* __restrict__ sb,
                                         const float x[D_MODEL], float y[ROWS]) {
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    float acc = sb[i];
#pragma unroll
    for (int k4 = 0; k4 < D_MODEL / 4; ++k4) {
      const float4 w4 = *reinterpret_cast<const float4*>(sw + i * D_MODEL + 4 * k4);
      acc = fmaf(w4.x, x[4 * k4 + 0], acc);
      acc = fmaf(w4.y, x[4 * k4 + 1], acc);
      acc = fmaf(w4.z, x[4 * k4 + 2], acc);
      acc = fmaf(w4.w, x[4 * k4 + 3], acc);
    }
    y[i] = acc;
  }
}

__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int n, int tid) {
  for (int t = tid; t < n; t += N_TOK) dst[t] = __ldg(src + t);
}

// softmax(q k^T / sqrt(dh)) v for this thread's query token, head by head; K/V broadcast from
// shared memory.  Two passes per head (max, then exp/sum/weighted values) like torch's softmax.
template <int NH>
__device__ __forceinline__ void attention(const OracleSmem& sm, const float q[D_MODEL], float att[D_MODEL]) {
  constexpr int DH = D_MODEL / NH;
  const float qscale = 1.f / sqrtf((float)DH);
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    float qh[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) qh[c] = q[h * DH + c] * qscale;
    float mx = -INFINITY;
    for (int j = 0; j < N_TOK; ++j) {
      float sdot = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) sdot = fmaf(qh[c], sm.k[j][h * DH + c], sdot);
      mx = fmaxf(mx, sdot);
    }
    float den = 0.f;
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    for (int j = 0; j < N_TOK; ++j) {
      float sdot = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) sdot = fmaf(qh[c], sm.k[j][h * DH + c], sdot);
      const float p = expf(sdot - mx);
      den += p;
#pragma unroll
      for (int c = 0; c < DH; ++c) acc[c] = fmaf(p, sm.v[j][h * DH + c], acc[c]);
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int c = 0; c < DH; ++c) att[h * DH + c] = acc[c] * inv;
  }
}

// The 128 token threads of a CTA synchronise on named barrier 1 (the tensor-core kernel has two
// more warps that must not take part).
__device__ __forceinline__ void token_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// relu(Linear(3,32)) * sqrt(32) + positional encoding for token `tid` of streamline `s`
__device__ __forceinline__ void embed_token(const ttl_oracle_weights& W, const float* __restrict__ dirs, int s,
                                            int tid, float x[D_MODEL]) {
  float t3[3];
  if (tid == 0) {
    t3[0] = __ldg(W.cls_token); t3[1] = __ldg(W.cls_token + 1); t3[2] = __ldg(W.cls_token + 2);
  } else {
    const float* d = dirs + ((size_t)s * (N_TOK - 1) + (tid - 1)) * 3;
    t3[0] = d[0]; t3[1] = d[1]; t3[2] = d[2];
  }
  const float scale = sqrtf((float)D_MODEL);
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) {
    float e = __ldg(W.emb_b + i);
    e = fmaf(__ldg(W.emb_w + 3 * i), t3[0], e);
    e = fmaf(__ldg(W.emb_w + 3 * i + 1), t3[1], e);
    e = fmaf(__ldg(W.emb_w + 3 * i + 2), t3[2], e);
    x[i] = fmaxf(e, 0.f) * scale + __ldg(W.pe + tid * D_MODEL + i);
  }
}

// x <- LayerNorm1(x + out_proj(attention(x))) for layer l; K/V of the layer go through sm.k / sm.v,
// the projection weights through sm.w / sm.b.
__device__ __forceinline__ void attention_block(const ttl_oracle_weights& W, int l, OracleSmem& sm, int tid,
                                                float x[D_MODEL]) {
  float q[D_MODEL];
  token_sync();
  stage(sm.w, W.in_proj_w[l], 3 * D_MODEL * D_MODEL, tid);
  stage(sm.b, W.in_proj_b[l], 3 * D_MODEL, tid);
  token_sync();
  {
    float kv[D_MODEL];
    matvec32<D_MODEL>(sm.w, sm.b, x, q);
    matvec32<D_MODEL>(sm.w + D_MODEL * D_MODEL, sm.b + D_MODEL, x, kv);
#pragma unroll
    for (int i = 0; i < D_MODEL; ++i) sm.k[tid][i] = kv[i];
    matvec32<D_MODEL>(sm.w + 2 * D_MODEL * D_MODEL, sm.b + 2 * D_MODEL, x, kv);
#pragma unroll
    for (int i = 0; i < D_MODEL; ++i) sm.v[tid][i] = kv[i];
  }
  token_sync();
  float att[D_MODEL];
  switch (W.n_head) {
    case 1: attention<1>(sm, q, att); break;
    case 2: attention<2>(sm, q, att); break;
    case 4: attention<4>(sm, q, att); break;
    default: attention<8>(sm, q, att); break;
  }
  token_sync();
  stage(sm.w, W.out_proj_w[l], D_MODEL * D_MODEL, tid);
  stage(sm.b, W.out_proj_b[l], D_MODEL, tid);
  token_sync();
  {
    float y[D_MODEL];
    matvec32<D_MODEL>(sm.w, sm.b, att, y);
#pragma unroll
    for (int i = 0; i < D_MODEL; ++i) x[i] += y[i];
  }
  layer_norm32(x, W.norm1_w[l], W.norm1_b[l]);
}

__device__ __forceinline__ void score_head(const ttl_oracle_weights& W, const float x[D_MODEL], float* out) {
  float y = __ldg(W.head_b);
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) y = fmaf(__ldg(W.head_w + i), x[i], y);
  *out = 1.f / (1.f + expf(-y));
}

// ---- fp32 tier: everything on the FP32 pipes ----
__global__ void __launch_bounds__(N_TOK, 3) oracle_forward_kernel(ttl_oracle_weights W,
                                                                 const float* __restrict__ dirs, int n,
                                                                 float* __restrict__ scores) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  OracleSmem& sm = *reinterpret_cast<OracleSmem*>(smem_raw);
  const int tid = threadIdx.x;   // token index
  for (int s = blockIdx.x; s < n; s += gridDim.x) {
    float x[D_MODEL];
    embed_token(W, dirs, s, tid, x);
    for (int l = 0; l < W.n_layers; ++l) {
      attention_block(W, l, sm, tid, x);
      // ---- feed forward 32 -> d_ff -> 32 in chunks of 64 hidden units ----
      float out[D_MODEL];
#pragma unroll
      for (int i = 0; i < D_MODEL; ++i) out[i] = 0.f;
      const int d_ff = W.d_ff;
      for (int c0 = 0; c0 < d_ff; c0 += FF_CHUNK) {
        token_sync();
        // W1 rows c0..c0+63 ([64][32], contiguous) and W2 columns c0..c0+63 ([32][64])
        stage(sm.w, W.lin1_w[l] + (size_t)c0 * D_MODEL, FF_CHUNK * D_MODEL, tid);
        stage(sm.b, W.lin1_b[l] + c0, FF_CHUNK, tid);
        for (int t = tid; t < D_MODEL * FF_CHUNK; t += N_TOK) {
          const int i = t / FF_CHUNK, j = t - i * FF_CHUNK;
          sm.w[FF_CHUNK * D_MODEL + t] = __ldg(W.lin2_w[l] + (size_t)i * d_ff + c0 + j);
        }
        token_sync();
        float hbuf[FF_CHUNK];
        matvec32<FF_CHUNK>(sm.w, sm.b, x, hbuf);
#pragma unroll
        for (int j = 0; j < FF_CHUNK; ++j) hbuf[j] = fmaxf(hbuf[j], 0.f);
        const float* w2 = sm.w + FF_CHUNK * D_MODEL;
#pragma unroll
        for (int i = 0; i < D_MODEL; ++i) {
          float acc = out[i];
#pragma unroll
          for (int j4 = 0; j4 < FF_CHUNK / 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(w2 + i * FF_CHUNK + 4 * j4);
            acc = fmaf(w4.x, hbuf[4 * j4 + 0], acc);
            acc = fmaf(w4.y, hbuf[4 * j4 + 1], acc);
            acc = fmaf(w4.z, hbuf[4 * j4 + 2], acc);
            acc = fmaf(w4.w, hbuf[4 * j4 + 3], acc);
          }
          out[i] = acc;
        }
      }
#pragma unroll
      for (int i = 0; i < D_MODEL; ++i) x[i] += out[i] + __ldg(W.lin2_b[l] + i);
      layer_norm32(x, W.norm2_w[l], W.norm2_b[l]);
    }
    if (tid == 0) score_head(W, x, scores + s);
  }
}

// ------------------------------------------------------------------------------------------
// fp16 tensor-core tier: the feed-forward block (91 % of the FLOPs) on tcgen05
// ------------------------------------------------------------------------------------------
// The reference's CUDA path runs the model under torch.autocast(fp16) (oracles/oracle.py:9,76):
// linear layers take fp16 operands and accumulate in fp32.  This tier does the same for
// linear1 / linear2; attention, LayerNorm and the embedding stay fp32 on the token threads.
//
// One CTA = one streamline at a time (128 tokens = the M of a 128-row UMMA), two CTAs per SM:
//   warps 0-3  token threads (thread t = token t = TMEM lane t): attention block in fp32, then per
//              64-wide hidden chunk c the epilogue of GEMM1: tcgen05.ld acc1 -> +b1, ReLU, fp16 ->
//              swizzled store into H[c&1], the A operand of GEMM2
//   warp 4     TMA producer: streams the fp16 weights (W1 chunk pair 8 KB + two W2 tiles 2x4 KB per
//              stage) through a 2-stage ring, running ahead across layers and streamlines
//   warp 5     MMA issuer: GEMM1(c) acc1[c&1] = X . W1_c^T (M128 N64 K32), then GEMM2(c-1)
//              acc2 += H_{c-1} . W2_{c-1}^T (M128 N32 K64), so the tensor pipe works on chunk c+1
//              while the token threads are in the epilogue of chunk c.
// TMEM: 256 columns per CTA (acc1 2 x 64, acc2 32).  The hidden activations never leave the SM.
namespace tcgen {
using namespace ttl_tc;
constexpr int CH = 64;                       // hidden units per chunk
constexpr int NST = 2;                       // weight ring stages (one stage = two chunks)
constexpr int STAGE_BYTES = 16384;           // 8 KB W1 pair tile + 2 x 4 KB W2 tiles
constexpr int OFF_XA = 0;                    // [128][128 B] SW128, x (fp16) in the first 64 B of a row
constexpr int OFF_H = 16384;                 // 2 x [128][128 B] SW128; aliases OracleSmem::k / ::v
constexpr int OFF_W = OFF_H + 32768;         // OracleSmem::w / ::b continue here (fp32 staging, b1)
constexpr int OFF_RING = OFF_H + (int)sizeof(OracleSmem);
constexpr int OFF_BAR = OFF_RING + NST * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;   // + alignment slack
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int ACC2_COL = 128;
static_assert(offsetof(OracleSmem, w) == 32768, "k and v must cover exactly the two H buffers");
static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
}  // namespace tcgen

__global__ void __launch_bounds__(tcgen::THREADS, 2)
oracle_forward_tc_kernel(ttl_oracle_weights W, const __grid_constant__ CUtensorMap tma_w1,
                         const __grid_constant__ CUtensorMap tma_w2, const float* __restrict__ dirs, int n,
                         float* __restrict__ scores) {
  using namespace tcgen;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ttl_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - raw);
  OracleSmem& sm = *reinterpret_cast<OracleSmem*>(sbase + OFF_H);
  const uint32_t bar0 = base + OFF_BAR;
  auto full = [&](int s) { return bar0 + 8u * s; };               // TMA -> MMA
  auto empty = [&](int s) { return bar0 + 8u * (NST + s); };      // MMA -> TMA
  auto acc1_full = [&](int b) { return bar0 + 8u * (2 * NST + b); };       // MMA -> tokens
  auto acc1_empty = [&](int b) { return bar0 + 8u * (2 * NST + 2 + b); };  // tokens -> MMA (4 warps)
  auto h_full = [&](int b) { return bar0 + 8u * (2 * NST + 4 + b); };      // tokens -> MMA (4 warps)
  auto h_empty = [&](int b) { return bar0 + 8u * (2 * NST + 6 + b); };     // MMA -> tokens
  const uint32_t x_ready = bar0 + 8u * (2 * NST + 8);                       // tokens -> MMA
  const uint32_t acc2_full = bar0 + 8u * (2 * NST + 9);                     // MMA -> tokens
  const uint32_t tmem_slot = bar0 + 8u * (2 * NST + 10);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sbase + OFF_BAR + 8 * (2 * NST + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_layers = W.n_layers;
  const int n_chunks = W.d_ff / CH;          // even (d_ff % 128 == 0)

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w2)) : "memory");
    for (int s = 0; s < NST; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc1_full(b), 1); mbar_init(acc1_empty(b), 4);
      mbar_init(h_full(b), 4); mbar_init(h_empty(b), 1);
    }
    mbar_init(x_ready, 1);
    mbar_init(acc2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 4) {
    if (lane == 0) {  // ===== TMA producer =====
      uint32_t gp = 0;   // chunk pairs issued so far
      for (int s = blockIdx.x; s < n; s += gridDim.x)
        for (int l = 0; l < n_layers; ++l)
          for (int j = 0; j < n_chunks / 2; ++j, ++gp) {
            const int st = gp % NST;
            mbar_wait(empty(st), ((gp / NST) & 1u) ^ 1u);
            mbar_arrive_expect_tx(full(st), STAGE_BYTES);
            const uint32_t dst = base + OFF_RING + st * STAGE_BYTES;
            tma_load_2d(dst, &tma_w1, full(st), 0, l * (W.d_ff / 2) + 64 * j);
            tma_load_2d(dst + 8192, &tma_w2, full(st), 128 * j, l * D_MODEL);
            tma_load_2d(dst + 12288, &tma_w2, full(st), 128 * j + 64, l * D_MODEL);
          }
    }
  } else if (warp == 5) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc1 = umma_idesc_f16(128, CH, 0);
      constexpr uint32_t idesc2 = umma_idesc_f16(128, D_MODEL, 0);
      const uint64_t desc_x = umma_desc_sw128(base + OFF_XA);
      uint32_t g = 0;       // chunks whose GEMM1 has been issued
      uint32_t n_ffn = 0;
      for (int s = blockIdx.x; s < n; s += gridDim.x)
        for (int l = 0; l < n_layers; ++l, ++n_ffn) {
          mbar_wait(x_ready, n_ffn & 1u);
          tc_fence_after();
          for (int c = 0; c <= n_chunks; ++c) {
            if (c < n_chunks) {
              const uint32_t b = g & 1u, u = g >> 1, gp = g >> 1;
              const int st = gp % NST;
              if ((c & 1) == 0) {
                mbar_wait(full(st), (gp / NST) & 1u);
                tc_fence_after();
              }
              mbar_wait(acc1_empty(b), (u & 1u) ^ 1u);
              tc_fence_after();
              const uint64_t desc_w1 = umma_desc_sw128(base + OFF_RING + st * STAGE_BYTES) + (uint64_t)(4 * (c & 1));
#pragma unroll
              for (int k = 0; k < D_MODEL / 16; ++k)
                tc_mma_bf16(tmem_base + b * CH, desc_x + (uint64_t)(2 * k), desc_w1 + (uint64_t)(2 * k), idesc1,
                            (uint32_t)(k != 0));
              tc_commit(acc1_full(b));
              ++g;
            }
            if (c >= 1) {
              const uint32_t g2 = g - (c < n_chunks ? 2u : 1u);   // global index of chunk c-1
              const uint32_t b = g2 & 1u, u = g2 >> 1, gp = g2 >> 1;
              const int st = gp % NST;
              mbar_wait(h_full(b), u & 1u);
              tc_fence_after();
              const uint64_t desc_h = umma_desc_sw128(base + OFF_H + b * 16384);
              const uint64_t desc_w2 =
                  umma_desc_sw128(base + OFF_RING + st * STAGE_BYTES + 8192 + ((c - 1) & 1) * 4096);
#pragma unroll
              for (int k = 0; k < CH / 16; ++k)
                tc_mma_bf16(tmem_base + ACC2_COL, desc_h + (uint64_t)(2 * k), desc_w2 + (uint64_t)(2 * k), idesc2,
                            (uint32_t)((c - 1) != 0 || k != 0));
              tc_commit(h_empty(b));
              if ((c - 1) & 1) tc_commit(empty(st));
              if (c == n_chunks) tc_commit(acc2_full);
            }
          }
        }
    }
  } else {  // ===== token threads =====
    const int tid = threadIdx.x;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    float* s_b1 = sm.w;        // linear1 bias of the layer, staged after the attention block
    uint32_t g = 0, n_ffn = 0;
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
      float x[D_MODEL];
      embed_token(W, dirs, s, tid, x);
      for (int l = 0; l < n_layers; ++l, ++n_ffn) {
        attention_block(W, l, sm, tid, x);
        // x (fp16) -> A operand of GEMM1; b1 -> shared memory
        token_sync();    // every thread is done with sm.w (out_proj) and sm.k / sm.v
        {
          uint32_t pk[D_MODEL / 2];
#pragma unroll
          for (int i = 0; i < D_MODEL / 2; ++i) {
            __half2 h = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
            pk[i] = *reinterpret_cast<uint32_t*>(&h);
          }
#pragma unroll
          for (int cch = 0; cch < 4; ++cch)
            *reinterpret_cast<uint4*>(sbase + OFF_XA + sw128_offset(tid, cch)) =
                make_uint4(pk[4 * cch], pk[4 * cch + 1], pk[4 * cch + 2], pk[4 * cch + 3]);
        }
        stage(s_b1, W.lin1_b[l], W.d_ff, tid);
        fence_proxy_async();
        token_sync();
        if (tid == 0) mbar_arrive(x_ready);
        for (int c = 0; c < n_chunks; ++c, ++g) {
          const uint32_t b = g & 1u, u = g >> 1;
          mbar_wait(acc1_full(b), u & 1u);
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tc_ld32(lane_base + b * CH, v0);
          tc_ld32(lane_base + b * CH + 32, v1);
          tc_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc1_empty(b));
          uint32_t pk[CH / 2];
          const float* bb = s_b1 + c * CH;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bb + 4 * j);
            __half2 h0 = __floats2half2_rn(fmaxf(__uint_as_float(v0[4 * j + 0]) + b4.x, 0.f),
                                           fmaxf(__uint_as_float(v0[4 * j + 1]) + b4.y, 0.f));
            __half2 h1 = __floats2half2_rn(fmaxf(__uint_as_float(v0[4 * j + 2]) + b4.z, 0.f),
                                           fmaxf(__uint_as_float(v0[4 * j + 3]) + b4.w, 0.f));
            pk[2 * j] = *reinterpret_cast<uint32_t*>(&h0);
            pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bb + 32 + 4 * j);
            __half2 h0 = __floats2half2_rn(fmaxf(__uint_as_float(v1[4 * j + 0]) + b4.x, 0.f),
                                           fmaxf(__uint_as_float(v1[4 * j + 1]) + b4.y, 0.f));
            __half2 h1 = __floats2half2_rn(fmaxf(__uint_as_float(v1[4 * j + 2]) + b4.z, 0.f),
                                           fmaxf(__uint_as_float(v1[4 * j + 3]) + b4.w, 0.f));
            pk[16 + 2 * j] = *reinterpret_cast<uint32_t*>(&h0);
            pk[16 + 2 * j + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
          mbar_wait(h_empty(b), (u & 1u) ^ 1u);   // GEMM2 of the chunk that used H[b] before has retired
          uint8_t* hrow = sbase + OFF_H + b * 16384;
#pragma unroll
          for (int cch = 0; cch < 8; ++cch)
            *reinterpret_cast<uint4*>(hrow + sw128_offset(tid, cch)) =
                make_uint4(pk[4 * cch], pk[4 * cch + 1], pk[4 * cch + 2], pk[4 * cch + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(h_full(b));
        }
        // linear2 output: acc2 -> registers, residual, LayerNorm 2
        mbar_wait(acc2_full, n_ffn & 1u);
        tc_fence_after();
        uint32_t o[32];
        tc_ld32(lane_base + ACC2_COL, o);
        tc_wait_ld();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < D_MODEL; ++i) x[i] += __uint_as_float(o[i]) + __ldg(W.lin2_b[l] + i);
        layer_norm32(x, W.norm2_w[l], W.norm2_b[l]);
      }
      if (tid == 0) score_head(W, x, scores + s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// fp16 tensor-core tier, full version: every matrix product of the encoder layer on tcgen05
// ------------------------------------------------------------------------------------------
// Per layer (token thread t = token t = TMEM lane t; one MMA-issuing thread; one TMA thread):
//   QKV      [128x32].[32x96]   x (fp16) in T0, in_proj weights in the ring      -> TMEM 128..223
//   tokens   +bias; Q (head-masked K=16 slices) -> T1, K -> T0, V^T -> VT (all fp16, SW128 tiles)
//   per head h, per half j of the keys (same double-buffered pipeline as the feed-forward chunks):
//     S      acc1[j] = Q_h . K_j^T                       (M128 N64 K16)
//     tokens row max / exp2 / row sum in fp32; P (fp16) written IN PLACE over S in tensor memory
//     PV     O[h][j] = P . V_j   (A operand from TMEM, M128 N16 K64)          -> TMEM 128..255
//   tokens   combine the two halves (each was scaled by its own row max), 1/sum, o (fp16) -> T0
//   out-proj [128x32].[32x32]                                                  -> TMEM 128..159
//   tokens   +bias, residual, LayerNorm 1, x (fp16) -> T0
//   FFN      per 64-wide chunk c: acc1[c%3] = [x | 1 1 0..] . [W1_c | b1_hi b1_lo 0..]^T (the bias rides
//            in a third K=16 slice as an fp16 hi+lo pair); tokens: ReLU + fp16 in place (one F2FP.RELU
//            per two units); acc2 += H_c . W2_c^T with H_c read from tensor memory -> TMEM 128..159
//   tokens   +bias, residual, LayerNorm 2
// The hidden activations and the attention probabilities never touch shared memory: the token
// threads read the fp32 accumulator with tcgen05.ld and write the fp16 operand of the next product
// back over it with tcgen05.st, so one chunk costs two mbarrier hand-offs and no proxy fence.
// Shared memory per CTA (two CTAs per SM): T0 16 KB [x or K | o or x'], T1 16 KB Q slices, VT 8 KB,
// weight ring 3 x 20 KB.  TMEM: 256 columns (acc1 2 x 64 | 128 shared by QKV, O, out-proj, acc2).
namespace tc2 {
using namespace ttl_tc;
constexpr int CH = 64;
constexpr int NST = 3;
constexpr int STAGE_BYTES = 20480;
constexpr int ST_W2 = 8192, ST_AUG = 16384;     // inside a feed-forward stage
constexpr int AUG_PAIR_BYTES = 4096;            // bias operand tiles of one chunk pair (2 x [64 units x K16], no swizzle)
constexpr int ST_PRM = 12288;                   // inside an attention stage
constexpr int PRM_FLOATS = 288;                 // in_proj_b 96 | out_proj_b 32 | norm1_w 32 | norm1_b 32 | lin2_b | norm2_w | norm2_b
constexpr int P_INB = 0, P_OUTB = 96, P_N1W = 128, P_N1B = 160, P_L2B = 192, P_N2W = 224, P_N2B = 256;
constexpr uint32_t FFN_TX = 16384 + AUG_PAIR_BYTES;
constexpr uint32_t ATT_TX = 12288 + PRM_FLOATS * 4;
constexpr int OFF_T0 = 0, OFF_T1 = 16384, OFF_VT = 32768, OFF_RING = 40960;
constexpr int OFF_BAR = OFF_RING + NST * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr int COL_B = 128;                      // QKV result / O blocks / out-proj result / FFN output
constexpr int COL_X = 160, COL_ONES = 176;      // FFN phase: x (fp16, 16 columns) and the [1 1 0..] slice (8 columns)
static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");
}  // namespace tc2

// MUFU.EX2 (2 ulp); arguments here are <= 0, results feed fp16 probabilities
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (max(v.x, 0), max(v.y, 0)) -> packed fp16 pair, one F2FP with the .relu modifier
__device__ __forceinline__ uint32_t pack_relu(float2 v) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v.y), "f"(v.x));
  return d;
}
// 8 consecutive fp32 -> one 16-byte chunk of fp16
__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
  return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b),
                    *reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
}
// x[32] (fp16) -> chunks chunk0..chunk0+3 of row `row` of a SW128 tile
__device__ __forceinline__ void store_row32(uint8_t* tile, int row, int chunk0, const float x[D_MODEL]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + ttl_tc::sw128_offset(row, chunk0 + c)) = pack8(x + 8 * c);
}
__device__ __forceinline__ void layer_norm32_s(float x[D_MODEL], const float* g, const float* b) {
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) mean += x[i];
  mean *= (1.f / D_MODEL);
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) { const float d = x[i] - mean; var = fmaf(d, d, var); }
  var *= (1.f / D_MODEL);
  const float inv = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int i = 0; i < D_MODEL; ++i) x[i] = (x[i] - mean) * inv * g[i] + b[i];
}

template <int NH>
__global__ void __launch_bounds__(tc2::THREADS, 2)
oracle_forward_tc2_kernel(ttl_oracle_weights W, const __grid_constant__ CUtensorMap tma_w1,
                          const __grid_constant__ CUtensorMap tma_w2, const __grid_constant__ CUtensorMap tma_wa,
                          const uint8_t* __restrict__ aug_all, const float* __restrict__ prm_all,
                          const float* __restrict__ dirs, int n, float* __restrict__ scores) {
  using namespace tc2;
  constexpr int DH = D_MODEL / NH;
  constexpr int NSUB = 2 * NH;                       // (head, key half) pairs
  constexpr int PV_N = NH == 1 ? 32 : 16;            // channels one PV product writes
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ttl_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sbase = smem_raw + (base - raw);
  const uint32_t bar0 = base + OFF_BAR;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (NST + s); };
  auto acc1_full = [&](uint32_t b) { return bar0 + 8u * (2 * NST + b); };   // MMA -> tokens: S / GEMM1 result
  auto h_full = [&](uint32_t b) { return bar0 + 8u * (2 * NST + 3 + b); };  // tokens -> MMA: fp16 operand in place (4 warps)
  const uint32_t x_ready = bar0 + 8u * (2 * NST + 6);   // tokens -> MMA: a shared-memory operand tile is complete
  const uint32_t done = bar0 + 8u * (2 * NST + 7);      // MMA -> tokens: a result is complete in TMEM
  const uint32_t tmem_slot = bar0 + 8u * (2 * NST + 8);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sbase + OFF_BAR + 8 * (2 * NST + 8));
  // acc1 buffers: columns 0, 64 (attention and feed-forward) and 192 (feed-forward only: the upper half
  // of the shared region is free then), each the fp32 accumulator of a 64-wide product and, after the
  // tokens' pass, its fp16 version in the first 32 columns.
  auto acc1_col = [&](uint32_t b) { return b == 2u ? 192u : b * 64u; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_layers = W.n_layers;
  const int n_chunks = W.d_ff / CH;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w2)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_wa)) : "memory");
    for (int s = 0; s < NST; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int b = 0; b < 3; ++b) { mbar_init(acc1_full(b), 1); mbar_init(h_full(b), 4); }
    mbar_init(x_ready, 1);
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 4) {
    // ===== TMA producer (whole warp, elected lane issues): per layer one attention stage, then d_ff/128
    //       feed-forward stages =====
    uint32_t gp = 0;
    for (int s = blockIdx.x; s < n; s += gridDim.x)
      for (int l = 0; l < n_layers; ++l) {
        {
          const int st = gp % NST;
          mbar_wait_park(empty(st), ((gp / NST) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(full(st), ATT_TX);
            const uint32_t dst = base + OFF_RING + st * STAGE_BYTES;
            tma_load_2d(dst, &tma_wa, full(st), 0, l * 96);
            bulk_load(dst + ST_PRM, prm_all + (size_t)l * PRM_FLOATS, PRM_FLOATS * 4, full(st));
          }
          __syncwarp();
          ++gp;
        }
        for (int j = 0; j < n_chunks / 2; ++j, ++gp) {
          const int st = gp % NST;
          mbar_wait_park(empty(st), ((gp / NST) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(full(st), FFN_TX);
            const uint32_t dst = base + OFF_RING + st * STAGE_BYTES;
            tma_load_2d(dst, &tma_w1, full(st), 0, l * (W.d_ff / 2) + 64 * j);
            tma_load_2d(dst + ST_W2, &tma_w2, full(st), 128 * j, l * D_MODEL);
            tma_load_2d(dst + ST_W2 + 4096, &tma_w2, full(st), 128 * j + 64, l * D_MODEL);
            bulk_load(dst + ST_AUG, aug_all + ((size_t)l * (n_chunks / 2) + j) * AUG_PAIR_BYTES, AUG_PAIR_BYTES, full(st));
          }
          __syncwarp();
        }
      }
  } else if (warp == 5) {
    // ===== MMA issuer (whole warp runs the loop and the waits; the elected lane issues) =====
    constexpr uint32_t id_qkv = umma_idesc_f16(128, 96, 0);
    constexpr uint32_t id_64 = umma_idesc_f16(128, 64, 0);
    constexpr uint32_t id_pv = umma_idesc_f16(128, PV_N, 0);
    constexpr uint32_t id_32 = umma_idesc_f16(128, 32, 0);
    const uint64_t d_t0 = umma_desc_sw128(base + OFF_T0);
    const uint64_t d_q = umma_desc_sw128(base + OFF_T1);
    const uint64_t d_ring = umma_desc_sw128(base + OFF_RING);
    const uint64_t d_vt = umma_desc_sw128(base + OFF_VT);
    uint32_t par_h = 0;  // bit b: parity of the next h_full[b] completion
    uint32_t gp = 0, n_sig = 0;
    auto wait_x = [&]() {
      mbar_wait(x_ready, n_sig & 1u);
      ++n_sig;
      tc_fence_after();
    };
    auto wait_h = [&](uint32_t b) {
      mbar_wait(h_full(b), (par_h >> b) & 1u);
      par_h ^= 1u << b;
      tc_fence_after();
    };
    // S(i) = Q_h . K_j^T into acc1[j]   (i = 2 h + j)
    auto issue_s = [&](int i) {
      const int h = i >> 1, j = i & 1;
      const uint64_t d_k = d_t0 + (uint64_t)(j * 512);       // key rows 64 j ...
      if (NH == 1) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
          tc_mma_bf16(tmem_base + j * CH, d_q + (uint64_t)(2 * t), d_k + (uint64_t)(2 * t), id_64, (uint32_t)(t != 0));
      } else {
        const int p = (h * DH) / 16;
        tc_mma_bf16(tmem_base + j * CH, d_q + (uint64_t)(2 * h), d_k + (uint64_t)(2 * p), id_64, 0u);
      }
      tc_commit(acc1_full(j));
    };
    for (int s = blockIdx.x; s < n; s += gridDim.x)
      for (int l = 0; l < n_layers; ++l) {
        const int stA = gp % NST;
        mbar_wait(full(stA), (gp / NST) & 1u);
        tc_fence_after();
        ++gp;
        const uint64_t d_wa = d_ring + (uint64_t)((stA * STAGE_BYTES) >> 4);
        // ---- QKV ----
        wait_x();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tc_mma_bf16(tmem_base + COL_B, d_t0 + (uint64_t)(2 * k), d_wa + (uint64_t)(2 * k), id_qkv, (uint32_t)(k != 0));
          tc_commit(done);
        }
        __syncwarp();
        // ---- attention: S(0), S(1); then per i: PV(i) from the fp16 P the tokens left in acc1[i&1],
        //      followed by S(i+2) into the same buffer (tcgen05.mma executes in issue order) ----
        wait_x();
        if (elect_one()) {
          issue_s(0);
          issue_s(1);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NSUB; ++i) {
          const int h = i >> 1, j = i & 1;
          wait_h((uint32_t)j);
          if (elect_one()) {
            const int p = NH == 1 ? 0 : (h * DH) / 16;
            const uint64_t d_v = d_vt + (uint64_t)((j * 4096 + p * 2048) >> 4);
            const uint32_t col = COL_B + (uint32_t)((h * 2 + j) * PV_N);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16_ts(tmem_base + col, tmem_base + j * CH + 8u * k, d_v + (uint64_t)(2 * k), id_pv, (uint32_t)(k != 0));
            if (i + 2 < NSUB) issue_s(i + 2);
            if (i == NSUB - 1) tc_commit(done);
          }
          __syncwarp();
        }
        // ---- output projection ----
        wait_x();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tc_mma_bf16(tmem_base + COL_B, d_t0 + (uint64_t)(4 + 2 * k), d_wa + (uint64_t)(4 + 2 * k), id_32, (uint32_t)(k != 0));
          tc_commit(done);
        }
        __syncwarp();
        // ---- feed forward ----
        // The chunk loop is unrolled by 6 = three ring stages (two chunks each) = two rounds of the three
        // accumulator buffers: inside a block every buffer, barrier and half-stage offset is a compile-time
        // constant and the three stages are per-layer values.  The issuing warp's own instruction stream
        // (two mod-3 divisions, descriptor assembly and a dozen R2UR per chunk) was what the token warps
        // waited for 57 % of this phase (profiles/r1_oracle_tc2_v5_ncu_full_raw.csv, source page).
        wait_x();
        uint32_t f_bar[3], e_bar[3], f_par[3];
        uint64_t d_w[3], d_aug[3];
        {
          uint32_t g = gp;
#pragma unroll
          for (int k = 0; k < 3; ++k, ++g) {
            const uint32_t st = g % NST;
            f_bar[k] = full((int)st);
            e_bar[k] = empty((int)st);
            f_par[k] = (g / NST) & 1u;                       // parity of pair k of block 0; flips every block
            d_w[k] = d_ring + (uint64_t)((st * STAGE_BYTES) >> 4);
            d_aug[k] = umma_desc_noswz(base + OFF_RING + st * STAGE_BYTES + ST_AUG, 128, 256);
          }
        }
        // GEMM1 of the chunk at position u (0..5) of a block into acc1[u % 3]: x . W1^T + b1
        auto issue_g1 = [&](int u) {
          const int k = u >> 1, half = u & 1;
          const uint32_t d = tmem_base + acc1_col((uint32_t)(u % 3));
          const uint64_t d_w1 = d_w[k] + (uint64_t)(4 * half);
          // A operand from tensor memory: x (fp16, 16 columns at COL_X) and the [1 1 0..] bias slice
          // (8 columns at COL_ONES); small SS-form MMAs are bound by the shared-memory read of A
          // (measured 48 cycles for N = 64, benchmarks/micro/umma_latency.cu), TS-form runs at 32
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            tc_mma_f16_ts(d, tmem_base + COL_X + 8u * kk, d_w1 + (uint64_t)(2 * kk), id_64, (uint32_t)(kk != 0));
          tc_mma_f16_ts(d, tmem_base + COL_ONES, d_aug[k] + (uint64_t)(half * ((AUG_PAIR_BYTES / 2) >> 4)), id_64, 1u);   // + b1
          tc_commit(acc1_full((uint32_t)(u % 3)));
        };
        auto wait_stage = [&](int k, uint32_t blk) {   // first chunk of a pair: its weights must have landed
          mbar_wait(f_bar[k], f_par[k] ^ (blk & 1u));
          tc_fence_after();
        };
        wait_stage(0, 0u);
        if (n_chunks > 2) wait_stage(1, 0u);
        if (elect_one()) {
          tc_commit(empty(stA));     // the tokens are done with the layer's parameters
          issue_g1(0);
          issue_g1(1);
          if (n_chunks > 2) issue_g1(2);
        }
        __syncwarp();
        for (int c0 = 0, blk = 0; c0 < n_chunks; c0 += 6, ++blk) {
#pragma unroll
          for (int u = 0; u < 6; ++u) {
            const int c = c0 + u;
            if (c >= n_chunks) break;
            const int un = (u + 3) % 6;                            // position of chunk c + 3 ...
            const uint32_t blkn = (uint32_t)blk + (u + 3 >= 6 ? 1u : 0u);   // ... and its block
            const bool more = c + 3 < n_chunks;
            if (more && (un & 1) == 0) wait_stage(un >> 1, blkn);
            wait_h((uint32_t)(u % 3));
            if (elect_one()) {
              const int k = u >> 1, half = u & 1;
              const uint64_t d_w2 = d_w[k] + (uint64_t)((ST_W2 + half * 4096) >> 4);
              const uint32_t a_h = tmem_base + acc1_col((uint32_t)(u % 3));
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                tc_mma_f16_ts(tmem_base + COL_B, a_h + 8u * kk, d_w2 + (uint64_t)(2 * kk), id_32,
                              (uint32_t)(u != 0 || kk != 0 || c0 != 0));
              if (half) tc_commit(e_bar[k]);   // both chunks of the stage have been consumed
              if (more) issue_g1(un);
              if (c == n_chunks - 1) tc_commit(done);
            }
            __syncwarp();
          }
        }
        gp += (uint32_t)(n_chunks / 2);
      }
  } else {  // ===== token threads =====
    const int tid = threadIdx.x;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint8_t* t0 = sbase + OFF_T0;
    const float sm_scale = 1.4426950408889634f / sqrtf((float)DH);   // log2(e) / sqrt(dh)
    uint32_t par_a = 0;  // bit b: parity of the next acc1_full[b] completion
    uint32_t gp = 0, n_done = 0;
    auto signal = [&]() {
      fence_proxy_async();
      token_sync();
      if (tid == 0) mbar_arrive(x_ready);
    };
    auto wait_done = [&]() {
      mbar_wait(done, n_done & 1u);
      ++n_done;
      tc_fence_after();
    };
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
      float x[D_MODEL];
      embed_token(W, dirs, s, tid, x);
      for (int l = 0; l < n_layers; ++l) {
        const int stA = gp % NST;
        const uint32_t parA = (gp / NST) & 1u;
        ++gp;
        const float* prm = reinterpret_cast<const float*>(sbase + OFF_RING + stA * STAGE_BYTES + ST_PRM);
        store_row32(t0, tid, 0, x);
        signal();
        mbar_wait(full(stA), parA);      // the layer's parameter block has landed
        wait_done();
        {  // ---- QKV epilogue: +bias, operand tiles of the attention products ----
          uint32_t r[32];
          float f[32];
          tc_ld32(lane_base + COL_B, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]) + prm[P_INB + i];
          uint8_t* t1 = sbase + OFF_T1;
          constexpr int NS = NH == 1 ? 2 : NH;
#pragma unroll
          for (int t = 0; t < NS; ++t) {
            const int p = NH == 1 ? t : (t * DH) / 16;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ch0 = 16 * p + 8 * e;
              const bool keep = NH == 1 || (ch0 / DH == t);
              *reinterpret_cast<uint4*>(t1 + sw128_offset(tid, 2 * t + e)) = keep ? pack8(f + ch0) : make_uint4(0, 0, 0, 0);
            }
          }
          tc_ld32(lane_base + COL_B + 32, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]) + prm[P_INB + 32 + i];
          store_row32(t0, tid, 0, f);     // K over x: the QKV product has retired
          tc_ld32(lane_base + COL_B + 64, r);
          tc_wait_ld();
          tc_fence_before();
          uint8_t* vt = sbase + OFF_VT + (tid >> 6) * 4096 + (tid & 7) * 2;
          const int kc = (tid & 63) >> 3;
#pragma unroll
          for (int c = 0; c < 32; ++c)
            *reinterpret_cast<__half*>(vt + sw128_offset(c, kc)) = __float2half_rn(__uint_as_float(r[c]) + prm[P_INB + 64 + c]);
        }
        signal();
        // ---- softmax over the key halves, head by head ----
        float mx[NSUB], ls[NSUB];
#pragma unroll
        for (int i = 0; i < NSUB; ++i) {
          const uint32_t b = (uint32_t)(i & 1);
          mbar_wait(acc1_full(b), (par_a >> b) & 1u);
          par_a ^= 1u << b;
          tc_fence_after();
          uint32_t v0[32], v1[32];
          tc_ld32(lane_base + b * CH, v0);
          tc_ld32(lane_base + b * CH + 32, v1);
          tc_wait_ld();
          float m = fmaxf(__uint_as_float(v0[0]), __uint_as_float(v1[0]));
#pragma unroll
          for (int k = 1; k < 32; ++k) m = fmaxf(m, fmaxf(__uint_as_float(v0[k]), __uint_as_float(v1[k])));
          const float2 sc2 = make_float2(sm_scale, sm_scale), mc2 = make_float2(-m * sm_scale, -m * sm_scale);
          float2 sum2 = make_float2(0.f, 0.f);
          uint32_t pk[32];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float2 a = __ffma2_rn(make_float2(__uint_as_float(v0[2 * k]), __uint_as_float(v0[2 * k + 1])), sc2, mc2);
            const float2 pp = make_float2(fast_exp2(a.x), fast_exp2(a.y));
            sum2 = __fadd2_rn(sum2, pp);
            __half2 hh = __float22half2_rn(pp);
            pk[k] = *reinterpret_cast<uint32_t*>(&hh);
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float2 a = __ffma2_rn(make_float2(__uint_as_float(v1[2 * k]), __uint_as_float(v1[2 * k + 1])), sc2, mc2);
            const float2 pp = make_float2(fast_exp2(a.x), fast_exp2(a.y));
            sum2 = __fadd2_rn(sum2, pp);
            __half2 hh = __float22half2_rn(pp);
            pk[16 + k] = *reinterpret_cast<uint32_t*>(&hh);
          }
          const float sum = sum2.x + sum2.y;
          mx[i] = m;
          ls[i] = sum;
          tc_st32(lane_base + b * CH, pk);      // P (fp16, 64 keys = 32 columns) over S
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h_full(b));
        }
        wait_done();
        {  // ---- attention output: merge the halves, normalise, o (fp16) -> T0 ----
          float o[D_MODEL];
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const float m = fmaxf(mx[2 * h], mx[2 * h + 1]);
            const float wa = fast_exp2((mx[2 * h] - m) * sm_scale), wb = fast_exp2((mx[2 * h + 1] - m) * sm_scale);
            const float inv = 1.f / (ls[2 * h] * wa + ls[2 * h + 1] * wb);
            const float ca = wa * inv, cb = wb * inv;
            if (NH >= 4) {        // DH valid columns inside a 16-column block
              constexpr int V = DH >= 8 ? 8 : DH;
              uint32_t a[8], c[8];
              const uint32_t off = (uint32_t)((h * DH) % 16);
              tc_ld8(lane_base + COL_B + (2 * h) * PV_N + off, a);
              tc_ld8(lane_base + COL_B + (2 * h + 1) * PV_N + off, c);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < V; ++e) o[h * DH + e] = __uint_as_float(a[e]) * ca + __uint_as_float(c[e]) * cb;
            } else if (NH == 2) {
              uint32_t a[16], c[16];
              tc_ld16(lane_base + COL_B + (2 * h) * PV_N, a);
              tc_ld16(lane_base + COL_B + (2 * h + 1) * PV_N, c);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[h * 16 + e] = __uint_as_float(a[e]) * ca + __uint_as_float(c[e]) * cb;
            } else {
              uint32_t a[32], c[32];
              tc_ld32(lane_base + COL_B, a);
              tc_ld32(lane_base + COL_B + 32, c);
              tc_wait_ld();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __uint_as_float(a[e]) * ca + __uint_as_float(c[e]) * cb;
            }
          }
          tc_fence_before();
          store_row32(t0, tid, 4, o);
        }
        signal();
        wait_done();
        {  // ---- out-proj epilogue: +bias, residual, LayerNorm 1, x (fp16) -> T0 ----
          uint32_t r[32];
          tc_ld32(lane_base + COL_B, r);
          tc_wait_ld();
          tc_fence_before();
#pragma unroll
          for (int i = 0; i < D_MODEL; ++i) x[i] += __uint_as_float(r[i]) + prm[P_OUTB + i];
          layer_norm32_s(x, prm + P_N1W, prm + P_N1B);
          // x (fp16) and the constant bias slice -> tensor memory, the A operands of GEMM1
          uint32_t xp[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
            xp[i] = *reinterpret_cast<uint32_t*>(&hh);
          }
          tc_st16(lane_base + COL_X, xp);
          const uint32_t ones[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          tc_st8(lane_base + COL_ONES, ones);
          tc_wait_st();
          tc_fence_before();
        }
        signal();
        // ---- feed forward ----
        uint32_t fb = 0;       // buffer of chunk c = c % 3
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t b = fb;
          fb = fb == 2u ? 0u : fb + 1u;
          mbar_wait(acc1_full(b), (par_a >> b) & 1u);
          par_a ^= 1u << b;
          tc_fence_after();
          const uint32_t col = acc1_col(b);
          uint32_t v0[32], v1[32];
          tc_ld32(lane_base + col, v0);
          tc_ld32(lane_base + col + 32, v1);
          tc_wait_ld();
          uint32_t pk[CH / 2];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_relu(make_float2(__uint_as_float(v0[2 * j]), __uint_as_float(v0[2 * j + 1])));
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[16 + j] = pack_relu(make_float2(__uint_as_float(v1[2 * j]), __uint_as_float(v1[2 * j + 1])));
          tc_st32(lane_base + col, pk);         // H chunk (fp16, 64 units = 32 columns) over the accumulator
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h_full(b));
        }
        gp += (uint32_t)(n_chunks / 2);
        wait_done();
        {
          uint32_t o[32];
          tc_ld32(lane_base + COL_B, o);
          tc_wait_ld();
          tc_fence_before();
#pragma unroll
          for (int i = 0; i < D_MODEL; ++i) x[i] += __uint_as_float(o[i]) + __ldg(W.lin2_b[l] + i);
          layer_norm32(x, W.norm2_w[l], W.norm2_b[l]);
        }
      }
      if (tid == 0) score_head(W, x, scores + s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// fp32 in_proj [96][32] and out_proj [32][32] -> fp16 [L][96][64]: row r = [in_proj_w[r][:] | out_proj_w[r][:] (r < 32)]
__global__ void pack_oracle_wa_kernel(ttl_oracle_weights W, __half* __restrict__ out) {
  const int per_layer = 96 * 64;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per_layer * W.n_layers) return;
  const int l = t / per_layer, e = t - l * per_layer;
  const int r = e / 64, c = e - r * 64;
  float v = 0.f;
  if (c < 32) v = __ldg(W.in_proj_w[l] + r * D_MODEL + c);
  else if (r < 32) v = __ldg(W.out_proj_w[l] + r * D_MODEL + (c - 32));
  out[t] = __float2half_rn(v);
}
// per-layer parameter block (tc2::P_*) and the linear1 biases as B-operand tiles of the bias slice:
// per chunk of 64 units a [64 x K16] fp16 tile in the no-swizzle core-matrix layout, unit n =
// [hi(b1), lo(b1), 0 ...] with hi + lo = b1 to ~22 bits
__global__ void pack_oracle_params_kernel(ttl_oracle_weights W, float* __restrict__ prm, uint8_t* __restrict__ aug) {
  const int l = blockIdx.x;
  for (int i = threadIdx.x; i < tc2::PRM_FLOATS; i += blockDim.x) {
    float v;
    if (i < 96) v = W.in_proj_b[l][i];
    else if (i < 128) v = W.out_proj_b[l][i - 96];
    else if (i < 160) v = W.norm1_w[l][i - 128];
    else if (i < 192) v = W.norm1_b[l][i - 160];
    else if (i < 224) v = W.lin2_b[l][i - 192];
    else if (i < 256) v = W.norm2_w[l][i - 224];
    else v = W.norm2_b[l][i - 256];
    prm[l * tc2::PRM_FLOATS + i] = v;
  }
  for (int u = threadIdx.x; u < W.d_ff; u += blockDim.x) {
    const float bv = W.lin1_b[l][u];
    const __half hi = __float2half_rn(bv);
    const __half lo = __float2half_rn(bv - __half2float(hi));
    const int chunk = u / 64, r = u % 64;
    uint8_t* t = aug + ((size_t)l * (W.d_ff / 64) + chunk) * 2048 + (r / 8) * 256 + (r % 8) * 16;
    __half2 hl = __halves2half2(hi, lo);
    *reinterpret_cast<uint4*>(t) = make_uint4(*reinterpret_cast<uint32_t*>(&hl), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(t + 128) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// fp32 [L][d_ff][32] linear1 weights -> fp16 [L][d_ff/2][64]: row 64j+r of a layer holds hidden
// unit 128j+r in its first 32 columns and unit 128j+64+r in the last 32, so that one 64-row
// SWIZZLE_128B TMA box carries two consecutive 64-unit chunks as K-offset 0 / 64 B operands.
__global__ void pack_oracle_w1_kernel(ttl_oracle_weights W, __half* __restrict__ out) {
  const int per_layer = W.d_ff * D_MODEL;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per_layer * W.n_layers) return;
  const int l = t / per_layer, e = t - l * per_layer;
  const int row = e / 64, col = e - row * 64;
  const int j = row / 64, r = row - j * 64;
  const int unit = 128 * j + (col >= 32 ? 64 : 0) + r;
  out[t] = __float2half_rn(__ldg(W.lin1_w[l] + (size_t)unit * D_MODEL + (col & 31)));
}
// fp32 [L][32][d_ff] linear2 weights -> fp16, same shape (K-major B operand of GEMM2 as is)
__global__ void pack_oracle_w2_kernel(ttl_oracle_weights W, __half* __restrict__ out) {
  const int per_layer = W.d_ff * D_MODEL;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per_layer * W.n_layers) return;
  const int l = t / per_layer, e = t - l * per_layer;
  out[t] = __float2half_rn(__ldg(W.lin2_w[l] + e));
}

}  // namespace

struct ttl_oracle_plan {
  ttl_oracle_weights w;
  __half* w1;            // packed linear1 weights (pack_oracle_w1_kernel)
  __half* w2;            // fp16 linear2 weights
  __half* wa;            // packed in_proj | out_proj weights (pack_oracle_wa_kernel)
  float* prm;            // per-layer parameter blocks
  uint8_t* aug;          // linear1 biases as bias-slice operand tiles (pack_oracle_params_kernel)
  CUtensorMap map_w1, map_w2, map_wa;
};

extern "C" {

static int feat_grid(int n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = ttl_div_up(n, FEAT_WARPS);
  return blocks < sms * 8 ? blocks : sms * 8;
}
static int feat_attr() {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(oracle_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(FEAT_WARPS * sizeof(FeatSmem)));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(oracle_features_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(FEAT_WARPS * sizeof(FeatSmem)));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  return 0;
}

int ttl_oracle_features(const float* points, const int64_t* offsets, int32_t n, float* dirs, void* stream) {
  if (!points || !offsets || !dirs) return TTL_ERR_BAD_ARG;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = feat_attr();
  if (rc) return rc;
  TTL_LAUNCH("oracle_features_kernel", s,
             oracle_features_kernel<<<feat_grid(n), 32 * FEAT_WARPS, FEAT_WARPS * sizeof(FeatSmem), s>>>(
                 points, (const long long*)offsets, n, dirs));
  TTL_CHECK_LAST();
  return 0;
}

int ttl_oracle_features_rows(const ttl_batch* b, int32_t cur, int32_t n_upper, float* dirs, void* stream) {
  if (!b || !dirs || (cur != 0 && cur != 1)) return TTL_ERR_BAD_ARG;
  if (n_upper <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = feat_attr();
  if (rc) return rc;
  TTL_LAUNCH("oracle_features_rows_kernel", s,
             oracle_features_rows_kernel<<<feat_grid(n_upper), 32 * FEAT_WARPS, FEAT_WARPS * sizeof(FeatSmem), s>>>(*b, cur, dirs));
  TTL_CHECK_LAST();
  return 0;
}

int ttl_oracle_forward(const ttl_oracle_weights* w, const float* dirs, int32_t n, float* scores, void* stream) {
  if (!w || !dirs || !scores) return TTL_ERR_BAD_ARG;
  if (w->d_model != D_MODEL || w->n_tokens != N_TOK || w->n_layers < 1 || w->n_layers > 8 ||
      (w->n_head != 1 && w->n_head != 2 && w->n_head != 4 && w->n_head != 8) || (w->d_ff % FF_CHUNK))
    return TTL_ERR_UNSUPPORTED;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(oracle_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(OracleSmem));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = n < sms * 3 ? n : sms * 3;
  TTL_LAUNCH("oracle_forward_kernel", s,
             oracle_forward_kernel<<<grid, N_TOK, sizeof(OracleSmem), s>>>(*w, dirs, n, scores));
  TTL_CHECK_LAST();
  return 0;
}


int64_t ttl_oracle_workspace_bytes(const ttl_oracle_weights* w) {
  if (!w || w->n_layers < 1 || w->n_layers > 8 || w->d_model != D_MODEL || w->d_ff <= 0) return -1;
  // fp16 copies of linear1 / linear2, packed attention weights, parameter blocks, linear1 biases
  return (int64_t)2 * w->n_layers * w->d_ff * D_MODEL * 2 + (int64_t)w->n_layers * (96 * 64 * 2 + 2048 + w->d_ff * 32);
}

int ttl_oracle_plan_create(ttl_oracle_plan** out, const ttl_oracle_weights* w, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  if (!out || !w || !workspace) return TTL_ERR_BAD_ARG;
  if (w->d_model != D_MODEL || w->n_tokens != N_TOK || w->n_layers < 1 || w->n_layers > 8 ||
      (w->n_head != 1 && w->n_head != 2 && w->n_head != 4 && w->n_head != 8) || (w->d_ff % 128) ||
      w->d_ff > 4096)
    return TTL_ERR_UNSUPPORTED;
  const int64_t need = ttl_oracle_workspace_bytes(w);
  if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 1023)) return TTL_ERR_BAD_ARG;
  ttl_oracle_plan* p = new (std::nothrow) ttl_oracle_plan();
  if (!p) return TTL_ERR_BAD_ARG;
  p->w = *w;
  p->w1 = static_cast<__half*>(workspace);
  p->w2 = p->w1 + (size_t)w->n_layers * w->d_ff * D_MODEL;
  p->wa = p->w2 + (size_t)w->n_layers * w->d_ff * D_MODEL;
  p->prm = reinterpret_cast<float*>(p->wa + (size_t)w->n_layers * 96 * 64);
  p->aug = reinterpret_cast<uint8_t*>(p->prm + (size_t)w->n_layers * 512);
  cudaStream_t s = (cudaStream_t)stream;
  const int tot = w->n_layers * w->d_ff * D_MODEL;
  TTL_LAUNCH("pack_oracle_w1_kernel", s, pack_oracle_w1_kernel<<<ttl_div_up(tot, 256), 256, 0, s>>>(*w, p->w1));
  TTL_LAUNCH("pack_oracle_w2_kernel", s, pack_oracle_w2_kernel<<<ttl_div_up(tot, 256), 256, 0, s>>>(*w, p->w2));
  TTL_LAUNCH("pack_oracle_wa_kernel", s,
             pack_oracle_wa_kernel<<<ttl_div_up(w->n_layers * 96 * 64, 256), 256, 0, s>>>(*w, p->wa));
  TTL_LAUNCH("pack_oracle_params_kernel", s, pack_oracle_params_kernel<<<w->n_layers, 256, 0, s>>>(*w, p->prm, p->aug));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { delete p; return (int)e; }
  ttl_tc::EncodeTiledFn fn = ttl_tc::get_encode_fn();
  if (!fn) { delete p; return TTL_ERR_DRIVER; }
  {  // linear1, packed [L * d_ff / 2][64] fp16, box 64 rows x 64 columns
    cuuint64_t dims[2] = {64, (cuuint64_t)w->n_layers * (w->d_ff / 2)};
    cuuint64_t strides[1] = {64 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&p->map_w1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->w1, dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { delete p; return TTL_ERR_DRIVER; }
  }
  {  // linear2 [L * 32][d_ff] fp16, box 32 rows x 64 columns
    cuuint64_t dims[2] = {(cuuint64_t)w->d_ff, (cuuint64_t)w->n_layers * D_MODEL};
    cuuint64_t strides[1] = {(cuuint64_t)w->d_ff * 2};
    cuuint32_t box[2] = {64, D_MODEL};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&p->map_w2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->w2, dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { delete p; return TTL_ERR_DRIVER; }
  }
  {  // in_proj | out_proj [L * 96][64] fp16, box 96 rows x 64 columns
    cuuint64_t dims[2] = {64, (cuuint64_t)w->n_layers * 96};
    cuuint64_t strides[1] = {64 * 2};
    cuuint32_t box[2] = {64, 96};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&p->map_wa, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p->wa, dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { delete p; return TTL_ERR_DRIVER; }
  }
  *out = p;
  return 0;
}

void ttl_oracle_plan_destroy(ttl_oracle_plan* plan) { delete plan; }

int ttl_oracle_forward_tc(ttl_oracle_plan* p, const float* dirs, int32_t n, float* scores, void* stream) {
  if (!p || !dirs || !scores) return TTL_ERR_BAD_ARG;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(oracle_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tcgen::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(oracle_forward_tc2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(oracle_forward_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(oracle_forward_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = n < sms * 2 ? n : sms * 2;
  static int ffn_only = -1;       // TTL_ORACLE_FFN_ONLY=1: attention on the FP32 pipes (first tensor-core version)
  if (ffn_only < 0) { const char* e = getenv("TTL_ORACLE_FFN_ONLY"); ffn_only = (e && e[0] == '1') ? 1 : 0; }
  const int nh = p->w.n_head;
  if (ffn_only || nh == 8) {
    // 8 heads of 4 channels do not fit the K = 16 slices of one Q tile: feed-forward blocks on tcgen05,
    // attention in fp32 on the token threads
    TTL_LAUNCH("oracle_forward_tc_kernel", s,
               oracle_forward_tc_kernel<<<grid, tcgen::THREADS, tcgen::SMEM_BYTES, s>>>(p->w, p->map_w1, p->map_w2,
                                                                                       dirs, n, scores));
  } else if (nh == 4) {
    TTL_LAUNCH("oracle_forward_tc2_kernel", s,
               oracle_forward_tc2_kernel<4><<<grid, tc2::THREADS, tc2::SMEM_BYTES, s>>>(
                   p->w, p->map_w1, p->map_w2, p->map_wa, p->aug, p->prm, dirs, n, scores));
  } else if (nh == 2) {
    TTL_LAUNCH("oracle_forward_tc2_kernel", s,
               oracle_forward_tc2_kernel<2><<<grid, tc2::THREADS, tc2::SMEM_BYTES, s>>>(
                   p->w, p->map_w1, p->map_w2, p->map_wa, p->aug, p->prm, dirs, n, scores));
  } else {
    TTL_LAUNCH("oracle_forward_tc2_kernel", s,
               oracle_forward_tc2_kernel<1><<<grid, tc2::THREADS, tc2::SMEM_BYTES, s>>>(
                   p->w, p->map_w1, p->map_w2, p->map_wa, p->aug, p->prm, dirs, n, scores));
  }
  TTL_CHECK_LAST();
  return 0;
}

}  // extern "C"
