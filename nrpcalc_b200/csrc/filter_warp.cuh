// Warp-private duplicate filter for SMALL leaves of packed records.
//
// Same job as k_filter_bloom (partition.cuh; replaces the dict of reference nrpcalc/base/hgraph.py:82-89): keep the records whose
// k-mer occurs at least twice in their leaf bucket, ordered by (k-mer, part, window), as one contiguous run of candidates.  What is
// different: a leaf is the work of ONE WARP, not of a CTA.  k_filter_bloom spends a leaf's turn in a chain of five to seven CTA
// barriers (5 us per leaf with three CTAs interleaved on an SM; ncu: barrier + wait stalls dominate, issue utilisation 53 %); here
// nothing is shared between warps, the only synchronisation is __syncwarp, and 24 to 48 warps per SM each work on a leaf of their own,
// so one warp's memory and shared-memory latencies are covered by the others.  It needs leaves of a few hundred records (two levels of
// up to 9 bits give 2^17 leaves of ~640 records for BASELINE config 2), which the partition kernels deliver at the same cost.
//
// Per leaf: pass 1 streams the records (coalesced 8-byte loads, four per lane in flight) and sets their bit in map A / map B
// (a second arrival) of the warp's shared-memory slice; pass 2 streams them again -- from L1 / L2, the leaf is ~5 KB -- and collects
// the suspects (bit set in B) with a ballot; the exact open-addressing test, the compaction of the confirmed records, the ordering
// by counting and the emission follow on at most SUSP suspects.  A leaf with more suspects (heavy sharing) is passed on whole and
// unordered: the global sort then runs (always correct, only slower).  No record is held in registers between the passes: ~40
// registers per thread.
#pragma once
#include "partition.cuh"

namespace nrp {

constexpr int kFwWarps = 10;               // warps per CTA: three CTAs of ten warps fit an SM with the 7.2 KB slices (30 leaves in flight)
constexpr int kFwThreads = kFwWarps * 32;

template <int MAPW, int SUSP>
struct alignas(16) FwWarp {
  uint32_t map_a[MAPW], map_b[MAPW];       // contiguous: cleared together
  uint64_t table[2 * SUSP];                // exact test: stored hashes of the suspects (open addressing, EMPTY = all ones)
  uint64_t c_recs[SUSP];                   // suspects in arrival order
  uint64_t o_recs[SUSP];                   // confirmed records, compacted
  uint8_t marked[2 * SUSP];
};

template <typename KeyT, int MAPW, int SUSP>
__global__ void __launch_bounds__(kFwThreads) k_filter_warp(const uint64_t* __restrict__ recs, LeafExtents leaves, uint32_t n_leaf, int leaf_bits,
                                                           PackFmt pf, KeyT* cand_keys, uint32_t* cand_vals, unsigned long long* n_cand,
                                                           uint32_t* n_oversized, uint32_t* unsorted) {
  static_assert((MAPW & (MAPW - 1)) == 0 && MAPW >= 64 && SUSP % 32 == 0, "sizes");
  constexpr int kLogBits = MAPW == 256 ? 13 : MAPW == 512 ? 14 : MAPW == 1024 ? 15 : -1;
  constexpr int kSlots = 2 * SUSP;
  static_assert(kLogBits > 0, "MAPW must be 256, 512 or 1024");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  FwWarp<MAPW, SUSP>& S = reinterpret_cast<FwWarp<MAPW, SUSP>*>(smem_raw)[wib];
  const uint32_t lt = lanemask_lt();
  const int vbits = pf.vbits;
  const uint32_t vmask = (uint32_t)((1ull << vbits) - 1ull);
  const uint64_t rem_mask = (1ull << (pf.n_eff() - leaf_bits)) - 1ull;     // stored hash bits below the leaf bits: what tells k-mers of a leaf apart
  auto fold_rem = [&](uint64_t r) -> uint32_t { const uint64_t rem = (r >> vbits) & rem_mask; return (uint32_t)rem ^ (uint32_t)(rem >> 32); };
  auto emit = [&](unsigned long long dst, uint64_t r, uint32_t lf) {
    cand_keys[dst] = (KeyT)pf.bij.inv(pf.compose(lf >> (leaf_bits - pf.drop), r >> vbits));
    cand_vals[dst] = (uint32_t)r & vmask;
  };
  uint4* const maps16 = reinterpret_cast<uint4*>(S.map_a);
  const uint4 zeros16 = make_uint4(0u, 0u, 0u, 0u);
  auto clear_maps = [&]() {
#pragma unroll
    for (int i = lane; i < 2 * MAPW / 4; i += 32) maps16[i] = zeros16;
  };
  clear_maps();
  for (int i = lane; i < kSlots; i += 32) { S.table[i] = ~0ull; S.marked[i] = 0; }
  __syncwarp();
  const uint32_t n_warps = gridDim.x * kFwWarps;
  // the bits that tell the k-mers of a leaf apart usually fit 32 bits and lie in one word of the record: one shift, one mask
  const int rem_bits = pf.n_eff() - leaf_bits;
  const bool rem32 = rem_bits <= 32 && (vbits >= 32 || vbits + rem_bits <= 32);
  const int rem_shift = vbits >= 32 ? vbits - 32 : vbits;
  const uint32_t rem_mask32 = rem_bits >= 32 ? 0xffffffffu : ((1u << rem_bits) - 1u);
  auto map_bit = [&](uint64_t r) -> uint32_t {
    const uint32_t x = rem32 ? (((vbits >= 32 ? (uint32_t)(r >> 32) : (uint32_t)r) >> rem_shift) & rem_mask32) : fold_rem(r);
    return (x * 0x9E3779B1u) >> (32 - kLogBits);
  };
  uint32_t lo, n, lo_next, n_next;
  leaves.get(blockIdx.x * kFwWarps + wib, n_leaf, lo, n);
  for (uint32_t leaf = blockIdx.x * kFwWarps + wib; leaf < n_leaf; leaf += n_warps, lo = lo_next, n = n_next) {
    leaves.get(leaf + n_warps, n_leaf, lo_next, n_next);         // the next leaf's extent: its load latency hides behind this leaf
    if (n < 2) continue;                                         // uniform across the warp
    const uint64_t* src = recs + lo;
    auto load4 = [&](uint32_t i0, uint64_t (&r)[4]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t i = i0 + q * 32 + lane;
        r[q] = i < n ? __ldg(src + i) : 0ull;
      }
    };
    // pass 1: every record sets its bit in map A; a record that finds it set also sets it in map B.  The loads of the next 128
    // records are issued before the current ones are used (eight 8-byte loads per lane in flight).
    uint64_t cur[4], nxt[4];
    load4(0, cur);
    for (uint32_t i0 = 0; i0 < n; i0 += 128) {
      load4(i0 + 128, nxt);                                      // past the end: no load is issued
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (i0 + q * 32 + lane < n) {
          const uint32_t bit = map_bit(cur[q]);
          const uint32_t w = bit >> 5, m = 1u << (bit & 31u);
          if (atomicOr(&S.map_a[w], m) & m) atomicOr(&S.map_b[w], m);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    }
    load4(0, cur);                                               // pass 2 reads the leaf again (L1 / L2): issued before the warp meets
    __syncwarp();
    // pass 2: suspects = records whose bit is set in map B (both records of every shared k-mer, plus the few that share a bit)
    uint32_t c1 = 0;                                             // uniform
    for (uint32_t i0 = 0; i0 < n; i0 += 128) {
      load4(i0 + 128, nxt);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bool sus = false;
        if (i0 + q * 32 + lane < n) {
          const uint32_t bit = map_bit(cur[q]);
          sus = (S.map_b[bit >> 5] >> (bit & 31u)) & 1u;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, sus);
        if (ballot != 0) {
          const uint32_t pos = c1 + __popc(ballot & lt);
          if (sus && pos < (uint32_t)SUSP) S.c_recs[pos] = cur[q];
          c1 += __popc(ballot);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    }
    __syncwarp();
    clear_maps();                                                // last use of the maps: clean for the next leaf
    __syncwarp();                                                // (no lane may set a bit of the next leaf in a word another lane still clears)
    if (c1 == 0) continue;
    if (c1 > (uint32_t)SUSP) {
      // heavy sharing (or a leaf far above the planned size): hand the whole leaf to the global sort
      unsigned long long base = 0;
      if (lane == 0) { base = atomicAdd(n_cand, (unsigned long long)n); atomicAdd(n_oversized, 1u); *unsorted = 1u; }
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < n; i += 32) emit(base + i, __ldg(src + i), leaf);
      __syncwarp();
      continue;
    }
    // exact test on the suspects: lane l owns the suspects l, l + 32, ...
    uint32_t slot[SUSP / 32];
#pragma unroll
    for (int t = 0; t < SUSP / 32; ++t) {
      const uint32_t s = (uint32_t)t * 32u + lane;
      slot[t] = 0xffffffffu;
      if (s < c1) {
        const uint64_t rec = S.c_recs[s];
        const uint64_t ident = rec >> vbits;
        uint32_t h = ((fold_rem(rec) * 0x85EBCA6Bu) >> 12) % (uint32_t)kSlots;
        for (;;) {
          const uint64_t old = atomic_cas(&S.table[h], ~0ull, ident);
          if (old == ~0ull) break;
          if (old == ident) { S.marked[h] = 1; break; }
          h = h + 1u == (uint32_t)kSlots ? 0u : h + 1u;
        }
        slot[t] = h;
      }
    }
    __syncwarp();
    uint32_t c2 = 0;                                             // uniform
#pragma unroll
    for (int t = 0; t < SUSP / 32; ++t) {
      const bool conf = slot[t] != 0xffffffffu && S.marked[slot[t]] != 0;
      const uint32_t ballot = __ballot_sync(0xffffffffu, conf);
      if (conf) S.o_recs[c2 + __popc(ballot & lt)] = S.c_recs[(uint32_t)t * 32u + lane];
      c2 += __popc(ballot);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < SUSP / 32; ++t)                         // leave the table clean: only the touched slots
      if (slot[t] != 0xffffffffu) { S.table[slot[t]] = ~0ull; S.marked[slot[t]] = 0; }
    if (c2 != 0) {
      // order by the packed word (= k-mer hash, part, window) by counting, emit as one contiguous run
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(n_cand, (unsigned long long)c2);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < c2; i += 32) {
        const uint64_t ck = S.o_recs[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < c2; ++j) {
          const uint64_t ok = S.o_recs[j];
          rank += (ok < ck || (ok == ck && j < i)) ? 1u : 0u;
        }
        emit(base + rank, ck, leaf);
      }
    }
    __syncwarp();
  }
}

}  // namespace nrp
