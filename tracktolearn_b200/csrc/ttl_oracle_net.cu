// TractOracle-Net scoring on the device.
//
// Reference: oracles/oracle.py:39-89 (OracleSingleton.predict: dipy set_number_of_points(s,128),
// np.diff, batches of 4096 through the model under fp16 autocast) and
// oracles/transformer_oracle.py:40-92 (CLS token + Linear(3,32)+ReLU scaled by sqrt(32), sinusoidal
// positional encoding, n_layers post-norm nn.TransformerEncoderLayer(d=32, n_head, ff=2048, relu),
// sigmoid(Linear(32,1)) of token 0).
//
//   oracle_features_kernel  one thread per streamline, the exact sequential algorithm of dipy's
//                           c_set_number_of_points (segment differences in float, arc lengths and
//                           interpolation in double) fused with the np.diff: no host resampling.
//   oracle_forward_kernel   one CTA per streamline, one thread per token (128).  The residual
//                           stream lives in registers, K/V of the current layer in shared memory,
//                           weights are streamed through shared memory in 16 KB chunks; all layers
//                           in one launch, fp32 throughout (the reference's CPU precision; its CUDA
//                           path autocasts to fp16).  ~147 MFLOP per streamline on the FP32 pipes.
#include "ttl_common.cuh"

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

__global__ void __launch_bounds__(128) oracle_features_kernel(const float* __restrict__ points,
                                                              const long long* __restrict__ offsets, int n,
                                                              float* __restrict__ dirs) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const long long o0 = offsets[s];
  resample_and_diff(points + o0 * 3, (int)(offsets[s + 1] - o0), dirs + (size_t)s * (kOraclePts - 1) * 3);
}

// The alive streamlines of a tracking batch, straight from the streamline buffer (what
// OracleStoppingCriterion / OracleReward score every step, stopping_criteria.py:113-154).
__global__ void __launch_bounds__(128) oracle_features_rows_kernel(ttl_batch b, int cur, float* __restrict__ dirs) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b.ctrl[cur]) return;
  const int row = b.alive[cur][r];
  resample_and_diff(b.points + (size_t)row * b.max_pts * 3, b.npts[row], dirs + (size_t)r * (kOraclePts - 1) * 3);
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

__global__ void __launch_bounds__(N_TOK, 3) oracle_forward_kernel(ttl_oracle_weights W,
                                                                 const float* __restrict__ dirs, int n,
                                                                 float* __restrict__ scores) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  OracleSmem& sm = *reinterpret_cast<OracleSmem*>(smem_raw);
  const int tid = threadIdx.x;   // token index
  const int n_head = W.n_head;
  for (int s = blockIdx.x; s < n; s += gridDim.x) {
    // ---- embedding: relu(Linear(3,32)) * sqrt(32) + positional encoding ----
    float x[D_MODEL];
    {
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
    for (int l = 0; l < W.n_layers; ++l) {
      // ---- self attention: K and V rows to shared memory, Q stays in registers ----
      float q[D_MODEL];
      __syncthreads();
      stage(sm.w, W.in_proj_w[l], 3 * D_MODEL * D_MODEL, tid);
      stage(sm.b, W.in_proj_b[l], 3 * D_MODEL, tid);
      __syncthreads();
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
      __syncthreads();
      float att[D_MODEL];
      switch (n_head) {
        case 1: attention<1>(sm, q, att); break;
        case 2: attention<2>(sm, q, att); break;
        case 4: attention<4>(sm, q, att); break;
        default: attention<8>(sm, q, att); break;
      }
      // ---- output projection, residual, LayerNorm 1 ----
      __syncthreads();
      stage(sm.w, W.out_proj_w[l], D_MODEL * D_MODEL, tid);
      stage(sm.b, W.out_proj_b[l], D_MODEL, tid);
      __syncthreads();
      {
        float y[D_MODEL];
        matvec32<D_MODEL>(sm.w, sm.b, att, y);
#pragma unroll
        for (int i = 0; i < D_MODEL; ++i) x[i] += y[i];
      }
      layer_norm32(x, W.norm1_w[l], W.norm1_b[l]);
      // ---- feed forward 32 -> d_ff -> 32 in chunks of 64 hidden units ----
      float out[D_MODEL];
#pragma unroll
      for (int i = 0; i < D_MODEL; ++i) out[i] = 0.f;
      const int d_ff = W.d_ff;
      for (int c0 = 0; c0 < d_ff; c0 += FF_CHUNK) {
        __syncthreads();
        // W1 rows c0..c0+63 ([64][32], contiguous) and W2 columns c0..c0+63 ([32][64])
        stage(sm.w, W.lin1_w[l] + (size_t)c0 * D_MODEL, FF_CHUNK * D_MODEL, tid);
        stage(sm.b, W.lin1_b[l] + c0, FF_CHUNK, tid);
        for (int t = tid; t < D_MODEL * FF_CHUNK; t += N_TOK) {
          const int i = t / FF_CHUNK, j = t - i * FF_CHUNK;
          sm.w[FF_CHUNK * D_MODEL + t] = __ldg(W.lin2_w[l] + (size_t)i * d_ff + c0 + j);
        }
        __syncthreads();
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
    if (tid == 0) {
      float y = __ldg(W.head_b);
#pragma unroll
      for (int i = 0; i < D_MODEL; ++i) y = fmaf(__ldg(W.head_w + i), x[i], y);
      scores[s] = 1.f / (1.f + expf(-y));
    }
  }
}

}  // namespace

extern "C" {

int ttl_oracle_features(const float* points, const int64_t* offsets, int32_t n, float* dirs, void* stream) {
  if (!points || !offsets || !dirs) return TTL_ERR_BAD_ARG;
  if (n <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  TTL_LAUNCH("oracle_features_kernel", s,
             oracle_features_kernel<<<ttl_div_up(n, 128), 128, 0, s>>>(points, (const long long*)offsets, n, dirs));
  TTL_CHECK_LAST();
  return 0;
}

int ttl_oracle_features_rows(const ttl_batch* b, int32_t cur, int32_t n_upper, float* dirs, void* stream) {
  if (!b || !dirs || (cur != 0 && cur != 1)) return TTL_ERR_BAD_ARG;
  if (n_upper <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  TTL_LAUNCH("oracle_features_rows_kernel", s,
             oracle_features_rows_kernel<<<ttl_div_up(n_upper, 128), 128, 0, s>>>(*b, cur, dirs));
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

}  // extern "C"
