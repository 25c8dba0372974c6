// Device-side uniquify of a packed part list.
//
// Replaces reference nrpcalc/base/hgraph.py:11-21 (uniquify_seq_list: upper-case, drop exact duplicates, id = index of the
// FIRST occurrence) together with the order in which get_homology_graph then processes the survivors (hgraph.py:218-223 pops
// the list from its end: descending index of the LAST occurrence).  On the host this is a dict over a million Python strings;
// here it is: one 64-bit hash per upper-cased part (a warp per part), a stable LSD sort of (hash, index), a byte comparison of
// every pair of neighbours with equal hashes (so that the result is EXACT -- two different parts with one hash raise a status bit
// and the caller falls back to the host path), run heads / tails, a compaction in descending last-occurrence order, and a gather of
// the surviving parts -- upper-cased -- into a new contiguous buffer that the build then reads.
#pragma once
#include "common.cuh"

namespace nrp {

constexpr uint32_t kUniqNonACGT = 1u;      // a character outside A/C/G/T (after upper-casing) was seen
constexpr uint32_t kUniqCollision = 2u;    // two DIFFERENT parts share a 64-bit hash: the result must not be used

__device__ __forceinline__ uint8_t upper_ascii(uint8_t b) { return (b >= 'a' && b <= 'z') ? (uint8_t)(b - 32) : b; }

struct PartSpan {
  const uint8_t* ascii; const int64_t* off; int64_t uniform_len;
  __device__ __forceinline__ int64_t begin(int64_t p) const { return uniform_len > 0 ? p * uniform_len : off[p]; }
  __device__ __forceinline__ int64_t length(int64_t p) const { return uniform_len > 0 ? uniform_len : off[p + 1] - off[p]; }
};

// keys[p] = hash of the upper-cased bytes and the length of part p, vals[p] = p.  One warp per part: lane l takes the bytes
// l, l + 32, ... weighted by powers of an odd multiplier (a polynomial mod 2^64), the warp adds up, a finaliser mixes.
__global__ void __launch_bounds__(256) k_part_hash(PartSpan S, int64_t n, uint64_t* keys, uint32_t* vals, uint32_t* status) {
  constexpr uint64_t P = 0x9E3779B97F4A7C15ull;
  const int lane = threadIdx.x & 31;
  uint64_t pw0 = 1, sq = P, p32 = P;
#pragma unroll
  for (int bit = 0; bit < 5; ++bit) { if ((lane >> bit) & 1) pw0 *= sq; sq *= sq; }    // P^lane; sq ends as P^32
  p32 = sq;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  bool other = false;
  for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
    const int64_t lo = S.begin(p), len = S.length(p);
    uint64_t sum = 0, pw = pw0;
    for (int64_t i = lane; i < len; i += 32, pw *= p32) {
      const uint8_t c = upper_ascii(S.ascii[lo + i]);
      other |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
      sum += (uint64_t)(c + 1u) * pw;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) {
      uint64_t h = sum ^ ((uint64_t)len * 0xD6E8FEB86659FD93ull);
      h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
      keys[p] = h;
      vals[p] = (uint32_t)p;
    }
  }
  if (__any_sync(0xffffffffu, other) && lane == 0) atomicOr(status, kUniqNonACGT);
}

// sorted (hash, index): head[i] = 1 where a new hash starts; neighbours with equal hashes must be equal parts
__global__ void k_uniq_heads(PartSpan S, const uint64_t* keys, const uint32_t* vals, int64_t n, uint32_t* head, uint32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool is_head = i == 0 || keys[i] != keys[i - 1];
  head[i] = is_head ? 1u : 0u;
  if (is_head) return;
  const int64_t a = vals[i], b = vals[i - 1];
  const int64_t len = S.length(a);
  bool same = len == S.length(b);
  const int64_t la = S.begin(a), lb = S.begin(b);
  for (int64_t t = 0; same && t < len; ++t) same = upper_ascii(S.ascii[la + t]) == upper_ascii(S.ascii[lb + t]);
  if (!same) atomicOr(status, kUniqCollision);
}

// run r (= head_scan[i] + head[i] - 1): first occurrence = the index at its head (the sort is stable), last occurrence = the
// index at its tail.  rep[last] = r + 1 marks the survivors by their LAST occurrence.
__global__ void k_uniq_runs(const uint64_t* keys, const uint32_t* vals, const uint32_t* head, const uint32_t* head_scan, int64_t n,
                            uint32_t* run_first, uint32_t* rep) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t run = head_scan[i] + head[i] - 1u;
  if (head[i]) run_first[run] = vals[i];
  if (i == n - 1 || keys[i + 1] != keys[i]) rep[vals[i]] = run + 1u;
}

// keep[j] = 1 iff part n - 1 - j is the last occurrence of its string (descending order of last occurrence = processing order)
__global__ void k_uniq_keep(const uint32_t* rep, int64_t n, uint32_t* keep) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) keep[j] = rep[n - 1 - j] != 0u ? 1u : 0u;
}

// processing rank q = keep_scan[j]: source[q] = last occurrence, ids[q] = first occurrence, lens[q] = length
__global__ void k_uniq_emit(PartSpan S, const uint32_t* rep, const uint32_t* keep, const uint32_t* keep_scan, const uint32_t* run_first, int64_t n,
                            uint32_t* source, uint32_t* ids, uint32_t* lens) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !keep[j]) return;
  const int64_t p = n - 1 - j;
  const uint32_t q = keep_scan[j];
  source[q] = (uint32_t)p;
  ids[q] = run_first[rep[p] - 1u];
  lens[q] = (uint32_t)S.length(p);
}

// the surviving parts, upper-cased, in processing order: a warp per part
__global__ void __launch_bounds__(256) k_uniq_gather(PartSpan S, const uint32_t* source, const unsigned long long* new_off, int64_t uniform_len,
                                                     int64_t n_unique, uint8_t* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_unique; q += warps) {
    const int64_t p = source[q];
    const int64_t lo = S.begin(p), len = S.length(p);
    const int64_t dst = uniform_len > 0 ? q * uniform_len : (int64_t)new_off[q];
    for (int64_t i = lane; i < len; i += 32) out[dst + i] = upper_ascii(S.ascii[lo + i]);
  }
}

}  // namespace nrp
