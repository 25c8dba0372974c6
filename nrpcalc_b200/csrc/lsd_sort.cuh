// Stable LSD radix sort for the SMALL arrays of the pipeline: the candidate records that survive the
// in-bucket duplicate filter (sorted by (k-mer, part)) and the expanded part-pair edge codes.
// The bulk of the records never goes through this sort -- see partition.cuh.
//
// One pass = per-tile digit counts -> exclusive scan over [digit][tile] -> stable scatter.  Inside a tile
// every warp owns a contiguous chunk and walks it 32 records at a time; ranks inside a 32-record row come
// from match.any (stable by lane order), ranks across rows/warps from per-warp running counters.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace nrp {

constexpr int kLsThreads = 256;
constexpr int kLsWarps = kLsThreads / 32;
constexpr int kLsRows = 16;                        // rows of 32 records per warp
constexpr int kLsTile = kLsThreads * kLsRows;      // 4096 records per tile
constexpr int kLsMaxPasses = 16;

struct LsdPass { int from_val; int shift; int bits; };
struct LsdPlan { int n_passes; LsdPass pass[kLsMaxPasses]; };

// passes covering val bits [0, val_bits) then key bits [0, key_bits), 8 bits at a time (val is the minor key)
inline LsdPlan lsd_plan(int key_bits, int val_bits) {
  LsdPlan plan; plan.n_passes = 0;
  for (int s = 0; s < val_bits; s += 8) plan.pass[plan.n_passes++] = {1, s, val_bits - s < 8 ? val_bits - s : 8};
  for (int s = 0; s < key_bits; s += 8) plan.pass[plan.n_passes++] = {0, s, key_bits - s < 8 ? key_bits - s : 8};
  return plan;
}
inline void lsd_plan_add(LsdPlan& plan, int from_val, int lo, int hi) {
  for (int s = lo; s < hi; s += 8) plan.pass[plan.n_passes++] = {from_val, s, hi - s < 8 ? hi - s : 8};
}

template <typename KeyT>
__device__ __forceinline__ uint32_t lsd_digit(KeyT key, uint32_t val, LsdPass p) {
  uint32_t mask = (1u << p.bits) - 1u;
  return p.from_val ? ((val >> p.shift) & mask) : ((uint32_t)(key >> p.shift) & mask);
}

template <typename KeyT, bool HAS_VAL>
__global__ void __launch_bounds__(kLsThreads) k_lsd_count(const KeyT* keys, const uint32_t* vals, int64_t n, LsdPass pass,
                                                          uint32_t* tile_counts, int64_t n_tiles) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kLsTile;
  for (int i = threadIdx.x; i < kLsTile; i += kLsThreads) {
    int64_t idx = base + i;
    if (idx < n) atomicAdd(&hist[lsd_digit<KeyT>(keys[idx], HAS_VAL ? vals[idx] : 0u, pass)], 1u);
  }
  __syncthreads();
  tile_counts[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];
}

template <typename KeyT, bool HAS_VAL>
__global__ void __launch_bounds__(kLsThreads) k_lsd_scatter(const KeyT* keys, const uint32_t* vals, int64_t n, LsdPass pass,
                                                            const uint32_t* tile_offsets, int64_t n_tiles,
                                                            KeyT* out_keys, uint32_t* out_vals) {
  __shared__ uint32_t cnt[kLsWarps][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kLsWarps * 256; i += kLsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t warp_base = (int64_t)blockIdx.x * kLsTile + (int64_t)wid * (32 * kLsRows);
  KeyT key[kLsRows];
  uint32_t val[kLsRows];
  uint32_t dig[kLsRows];
#pragma unroll
  for (int r = 0; r < kLsRows; ++r) {
    int64_t idx = warp_base + r * 32 + lane;
    bool ok = idx < n;
    key[r] = ok ? keys[idx] : KeyT(0);
    val[r] = (ok && HAS_VAL) ? vals[idx] : 0u;
    dig[r] = ok ? lsd_digit<KeyT>(key[r], val[r], pass) : 0xffffffffu;
    if (ok) atomicAdd(&cnt[wid][dig[r]], 1u);
  }
  __syncthreads();
  {  // exclusive prefix over (digit major, warp minor), seeded with the tile's global offset of the digit
    const int d = threadIdx.x;
    uint32_t run = tile_offsets[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kLsWarps; ++w) {
      uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < kLsRows; ++r) {
    bool ok = dig[r] != 0xffffffffu;
    uint32_t probe = ok ? dig[r] : (0x10000u + lane);          // inactive lanes never match anyone
    uint32_t peers = __match_any_sync(0xffffffffu, probe);
    uint32_t base = ok ? cnt[wid][dig[r]] : 0u;
    uint32_t rank = __popc(peers & lt);
    __syncwarp();
    if (ok && rank == 0) cnt[wid][dig[r]] = base + __popc(peers);
    __syncwarp();
    if (ok) {
      uint32_t dst = base + rank;
      out_keys[dst] = key[r];
      if (HAS_VAL) out_vals[dst] = val[r];
    }
  }
}

// Sorts (keys[, vals]) in place semantics: result ends in the buffers returned through *res_keys/*res_vals
// (either the inputs or the alternates).  counts: scratch of 2*(256*n_tiles+1)+scan scratch uint32.
template <typename KeyT, bool HAS_VAL>
void lsd_sort(KeyT* keys, uint32_t* vals, KeyT* alt_keys, uint32_t* alt_vals, int64_t n, const LsdPlan& plan,
              uint32_t* scratch, cudaStream_t stream, KeyT** res_keys, uint32_t** res_vals) {
  KeyT* ck = keys; uint32_t* cv = vals; KeyT* ak = alt_keys; uint32_t* av = alt_vals;
  if (n > 1) {
    const int64_t n_tiles = (n + kLsTile - 1) / kLsTile;
    const int64_t n_counts = 256 * n_tiles;
    uint32_t* counts = scratch;
    uint32_t* offsets = scratch + n_counts + 1;
    uint32_t* scan_scratch = offsets + n_counts + 2;
    for (int p = 0; p < plan.n_passes; ++p) {
      NRP_LAUNCH((k_lsd_count<KeyT, HAS_VAL><<<(unsigned)n_tiles, kLsThreads, 0, stream>>>(ck, cv, n, plan.pass[p], counts, n_tiles)));
      exclusive_scan<uint32_t, uint32_t>(counts, offsets, n_counts, scan_scratch, stream);
      NRP_LAUNCH((k_lsd_scatter<KeyT, HAS_VAL><<<(unsigned)n_tiles, kLsThreads, 0, stream>>>(ck, cv, n, plan.pass[p], offsets, n_tiles, ak, av)));
      KeyT* tk = ck; ck = ak; ak = tk;
      uint32_t* tv = cv; cv = av; av = tv;
    }
  }
  *res_keys = ck;
  if (res_vals) *res_vals = cv;
}

inline int64_t lsd_scratch_elems(int64_t n) {
  const int64_t n_tiles = (n + kLsTile - 1) / kLsTile;
  const int64_t n_counts = 256 * n_tiles;
  return 2 * (n_counts + 2) + scan_scratch_elems(n_counts + 1) + 16;
}

}  // namespace nrp
