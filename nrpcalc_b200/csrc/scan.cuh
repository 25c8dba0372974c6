// Device-wide exclusive prefix sums (hierarchical: block scan -> scan of block sums -> add back).
// out has n + 1 entries; out[n] is the grand total.  Inputs beyond n read as zero.
#pragma once
#include "common.cuh"

namespace nrp {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* warp_sums /* 32 entries of smem */, T& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  T inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    T w = lane < nw ? warp_sums[lane] : T(0);
    T winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    warp_sums[lane] = winc - w;            // exclusive prefix of warp totals
    if (lane == 31) warp_sums[32] = winc;  // block total
  }
  __syncthreads();
  T res = warp_sums[wid] + inc - v;
  total = warp_sums[32];
  __syncthreads();
  return res;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const Tin* in, Tout* out, Tout* tile_sums, int64_t n_in, int64_t n_out) {
  __shared__ Tout warp_sums[33];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  Tout v[kScanItems];
  Tout sum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n_in ? (Tout)in[idx] : Tout(0);
    sum += v[i];
  }
  Tout total;
  Tout run = block_exclusive_scan<Tout>(sum, warp_sums, total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n_out) out[idx] = run;
    run += v[i];
  }
  if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = total;
}

template <typename Tout>
__global__ void __launch_bounds__(kScanThreads) k_scan_add(Tout* out, const Tout* tile_offsets, int64_t n_out) {
  const Tout add = tile_offsets[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
    int64_t idx = base + i;
    if (idx < n_out) out[idx] += add;
  }
}

// scratch needed (in elements of Tout) for a scan of n_out outputs
inline int64_t scan_scratch_elems(int64_t n_out) {
  int64_t total = 0;
  while (n_out > kScanTile) {
    int64_t tiles = (n_out + kScanTile - 1) / kScanTile;
    total += 2 * (tiles + 1);
    n_out = tiles + 1;
  }
  return total + 8;
}

template <typename Tin, typename Tout>
void exclusive_scan(const Tin* in, Tout* out, int64_t n_in, Tout* scratch, cudaStream_t stream) {
  const int64_t n_out = n_in + 1;
  const int64_t tiles = (n_out + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    NRP_LAUNCH((k_scan_tiles<Tin, Tout><<<1, kScanThreads, 0, stream>>>(in, out, nullptr, n_in, n_out)));
    return;
  }
  Tout* sums = scratch;                 // [0, tiles): per-tile totals
  Tout* offsets = scratch + tiles + 1;  // [tiles+1, 2*tiles+2): their exclusive scan (+ grand total)
  NRP_LAUNCH((k_scan_tiles<Tin, Tout><<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, out, sums, n_in, n_out)));
  exclusive_scan<Tout, Tout>(sums, offsets, tiles, scratch + 2 * (tiles + 1), stream);
  NRP_LAUNCH((k_scan_add<Tout><<<(unsigned)tiles, kScanThreads, 0, stream>>>(out, offsets, n_out)));
}

}  // namespace nrp
