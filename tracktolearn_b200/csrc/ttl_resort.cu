// Periodic tip sort of the alive list (north_star (1): "voxel-local streamlines kept together by a
// per-step tip sort").  In the streaming device mode a streamline's position in the alive list is free --
// rows keep their identity, results are per row -- so between two steps the list can be re-ordered by the
// voxel raster index of every streamline's tip: neighbours in the list then gather from the same or
// adjacent voxels again after their paths have diverged from the seeding order (reset-time slot order,
// ttl_batch.order).  One re-sort = radix sort of (voxel key, rank) pairs (cub) + one pass that moves the
// per-rank data (rank record, alive id, operand row) from the `cur` buffers to the `cur ^ 1` buffers; the
// caller then flips `cur` exactly as after a step.
//
// The reference has no counterpart (it shuffles its seeds, tracking/tracker.py:94, and pays for it in the
// gather of environments/env.py:538-541); DESIGN.md section 4 records what the sort buys on B200.
#include <cub/device/device_radix_sort.cuh>

#include "ttl_common.cuh"

namespace {

__global__ void __launch_bounds__(256) resort_keys_kernel(ttl_volume v, ttl_batch b, int cur, uint32_t* __restrict__ keys,
                                                          int32_t* __restrict__ ranks, int n_max) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_max) return;
  const int n = b.ctrl[cur];
  uint32_t key = 0xffffffffu;       // ranks beyond the alive count sort to the end
  if (r < n) {
    const float4* rec = reinterpret_cast<const float4*>(b.rank_rec[cur]) + 2 * (size_t)r;
    const float4 a = __ldg(rec), c = __ldg(rec + 1);
    // tip = (a.z, a.w, c.x); voxel of the tip (lattice at integer coordinates), clamped; NaN -> 0
    const int x = min(max((int)floorf(a.z), 0), v.X - 1);
    const int y = min(max((int)floorf(a.w), 0), v.Y - 1);
    const int z = min(max((int)floorf(c.x), 0), v.Z - 1);
    key = (uint32_t)((x * v.Y + y) * v.Z + z);
  }
  keys[r] = key;
  ranks[r] = r;
}

// one warp per destination rank: rank record (32 B), alive id, operand row
__global__ void __launch_bounds__(256) resort_move_kernel(ttl_batch b, int cur, const int32_t* __restrict__ perm,
                                                          int row_bytes, int n_max) {
  const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (d == 0 && lane == 0) {        // the list itself is unchanged: same count, same refill cursor
    b.ctrl[cur ^ 1] = b.ctrl[cur];
    b.ctrl[12 + (cur ^ 1)] = b.ctrl[12 + cur];
    // the stop counters of the parity that becomes current again still hold the last step's counts (the
    // state kernel clears the OTHER set): clear them for the step that follows
    const int n_super = (b.max_groups + 63) / 64;
    for (int j = 0; j < n_super; ++j) b.sg_stops[(cur ^ 1) * n_super + j] = 0;
  }
  if (d >= n_max || d >= b.ctrl[cur]) return;
  const int src = perm[d];
  if (lane < 2) {
    const float4* s = reinterpret_cast<const float4*>(b.rank_rec[cur]) + 2 * (size_t)src;
    reinterpret_cast<float4*>(b.rank_rec[cur ^ 1])[2 * (size_t)d + lane] = __ldg(s + lane);
  }
  if (lane == 2) b.alive[cur ^ 1][d] = b.alive[cur][src];
  const uint4* srow = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(b.state_bf16[cur]) + (size_t)src * row_bytes);
  uint4* drow = reinterpret_cast<uint4*>(static_cast<uint8_t*>(b.state_bf16[cur ^ 1]) + (size_t)d * row_bytes);
  for (int q = lane; q < (row_bytes >> 4); q += 32) drow[q] = __ldg(srow + q);
}

}  // namespace

extern "C" {

int64_t ttl_env_resort_workspace_bytes(int32_t n_slots) {
  if (n_slots <= 0) return -1;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, n_slots);
  // keys in / out, ranks in / out, cub scratch
  return (int64_t)(((size_t)n_slots * 16 + 1023) / 1024 * 1024 + tmp + 1024);
}

int ttl_env_resort(const ttl_volume* vol, const ttl_batch* b, int32_t cur, int32_t n_upper, void* workspace,
                   int64_t workspace_bytes, void* stream) {
  if (!vol || !b || (cur != 0 && cur != 1) || !workspace) return TTL_ERR_BAD_ARG;
  if (b->bf16_layout != 1 || b->state[0] || b->state[1]) return TTL_ERR_UNSUPPORTED;   // operand-only device mode
  if (n_upper > b->n_slots) n_upper = b->n_slots;
  if (n_upper <= 0) return 0;
  if (workspace_bytes < ttl_env_resort_workspace_bytes(b->n_slots)) return TTL_ERR_BAD_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)b->n_slots;
  uint32_t* keys_in = static_cast<uint32_t*>(workspace);
  uint32_t* keys_out = keys_in + n;
  int32_t* ranks_in = reinterpret_cast<int32_t*>(keys_out + n);
  int32_t* ranks_out = ranks_in + n;
  uint8_t* tmp = static_cast<uint8_t*>(workspace) + (n * 16 + 1023) / 1024 * 1024;
  size_t tmp_bytes = (size_t)workspace_bytes - (size_t)(tmp - static_cast<uint8_t*>(workspace));
  TTL_LAUNCH("resort_keys_kernel", s,
             resort_keys_kernel<<<ttl_div_up(n_upper, 256), 256, 0, s>>>(*vol, *b, cur, keys_in, ranks_in, n_upper));
  // voxel keys need ceil(log2(X*Y*Z)) bits; ranks beyond the alive count carry the all-ones key
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, ranks_in, ranks_out, n_upper, 0, 32, s);
  if (e != cudaSuccess) return (int)e;
  TTL_LAUNCHED();
  const int row_bytes = b->ld_bf16 * (b->operand_fmt == TTL_OPERAND_TF32 ? 4 : 2);
  TTL_LAUNCH("resort_move_kernel", s,
             resort_move_kernel<<<ttl_div_up((long long)n_upper * 32, 256), 256, 0, s>>>(*b, cur, ranks_out, row_bytes, n_upper));
  TTL_CHECK_LAST();
  return 0;
}

}  // extern "C"
