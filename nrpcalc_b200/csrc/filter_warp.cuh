// Warp-private duplicate filter for SMALL leaves of packed records.
//
// Same job as k_filter_bloom (partition.cuh; replaces the dict of reference nrpcalc/base/hgraph.py:82-89): keep the records whose
// k-mer occurs at least twice in their leaf bucket, ordered by (k-mer, part, window), as one contiguous run of candidates.  What is
// different: a leaf is the work of ONE WARP, not of a CTA.  k_filter_bloom spends a leaf's turn in a chain of five to seven CTA
// barriers (5 us per leaf with three CTAs interleaved on an SM; ncu: barrier + wait stalls dominate, issue utilisation 53 %); here
// nothing is shared between warps, the only synchronisation is __syncwarp, and 20 to 40 warps per SM each work on a leaf of their
// own, so one warp's memory and shared-memory latencies are covered by the others.  It needs leaves of a few hundred records (two
// levels of up to 9 bits give 2^17 leaves of ~640 records for BASELINE config 2), which the partition kernels deliver at the same cost.
//
// Per leaf: every record sets its bit in map A / map B (a second arrival) of the warp's shared-memory slice (pass 1); the records
// whose bit is set in B are the suspects (pass 2); the exact open-addressing test, the compaction of the confirmed records, the
// ordering by counting and the emission follow on at most SUSP suspects.  A leaf with more suspects (heavy sharing) or more than
// 1024 records is passed on whole and unordered: the global sort then runs (always correct, only slower).
//
// How it got here (config 2, 8.4e7 records; profiles/README.md, ncu_full_r02n / o / p / q.txt):
//   first version     two streaming passes over the leaf with 8-byte loads, a ballot per record          0.382 ms  (k_filter_bloom: 0.378)
//   + loads one batch ahead, the next leaf's extent ahead, ten-warp CTAs                                  0.329 ms  92 instr / record, issue 72 %
//   second version    record pairs in 16-byte loads, check-free full batches, one shift for the k-mer bits,
//                     per-lane suspect masks instead of ballots                                           0.303 ms  70 instr / record, long scoreboard on top
//   this version      the leaf's k-mer bits in registers: all loads at once, nothing read twice          0.299 ms  57 instr / record
// Smaller leaves do not help (2^18 leaves of 320 records: 0.39 ms -- the per-leaf latency chain, not the records, sets the pace).
#pragma once
#include "partition.cuh"

namespace nrp {

constexpr int kFwWarps = 10;               // warps per CTA: three CTAs of ten warps fit an SM with the 7.2 KB slices (30 leaves in flight)
constexpr int kFwThreads = kFwWarps * 32;

template <int MAPW, int SUSP>
struct alignas(16) FwWarp {
  uint32_t map_a[MAPW], map_b[MAPW];       // contiguous: cleared together (and map_b is addressed as map_a + MAPW)
  uint64_t table[2 * SUSP];                // exact test: stored hashes of the suspects (open addressing, EMPTY = all ones)
  uint64_t c_recs[SUSP];                   // suspects in arrival order
  uint64_t o_recs[SUSP];                   // confirmed records, compacted
  uint8_t marked[2 * SUSP];
};

// A leaf of <= 1024 records is 32 words per lane once only the 32 bits that tell its k-mers apart are kept -- "the 32 bits of the
// word from bit vbits on": one shift (VHI: of the high word; else a funnel shift), no mask, whatever else those bits hold is the
// same for every record of the leaf.  Records are addressed as PAIRS (16 bytes; leaf regions start at multiples of four records),
// two pairs per lane and batch of 128 records; a batch that lies wholly inside the leaf takes a path without per-record checks.
//   * ALL loads of the leaf are issued at once (16 predicated 16-byte loads per lane; with VHI the compiler fetches just the
//     high words) and reduced to x[32] in registers;
//   * pass 1 and pass 2 run on registers: nothing is read twice, no load sits between the two passes;
//   * only the suspects' full records are fetched again (L1 / L2);
//   * the claim of the output run (one global atomic) is issued before the ranks are computed and consumed after.
// Leaves of more than 1024 records (the fixed regions of a plan for <= 800 end at ~1200: nine sigma) are passed on whole.
template <typename KeyT, int MAPW, int SUSP, bool VHI>
__global__ void __launch_bounds__(kFwThreads) k_filter_warp(const uint64_t* __restrict__ recs, LeafExtents leaves, uint32_t n_leaf, int leaf_bits,
                                                            PackFmt pf, KeyT* cand_keys, uint32_t* cand_vals, unsigned long long* n_cand,
                                                            uint32_t* n_oversized, uint32_t* unsorted) {
  static_assert((MAPW & (MAPW - 1)) == 0 && MAPW >= 64 && SUSP % 32 == 0, "sizes");
  constexpr int kLogBits = MAPW == 256 ? 13 : MAPW == 512 ? 14 : MAPW == 1024 ? 15 : -1;
  constexpr int kSlots = 2 * SUSP;
  constexpr int kB = 8;                                          // batches of 128 records a lane keeps: 4 words each
  static_assert(kLogBits > 0, "MAPW must be 256, 512 or 1024");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  FwWarp<MAPW, SUSP>& S = reinterpret_cast<FwWarp<MAPW, SUSP>*>(smem_raw)[wib];
  const uint32_t lt = lanemask_lt();
  const int vbits = pf.vbits;
  const uint32_t vmask = (uint32_t)((1ull << vbits) - 1ull);
  const int xsh = VHI ? vbits - 32 : vbits;
  auto disc = [&](uint32_t lo, uint32_t hi) -> uint32_t { return VHI ? hi >> xsh : __funnelshift_r(lo, hi, xsh); };
  auto emit = [&](unsigned long long dst, uint64_t r, uint32_t lf) {
    cand_keys[dst] = (KeyT)pf.bij.inv(pf.compose(lf >> (leaf_bits - pf.drop), r >> vbits));
    cand_vals[dst] = (uint32_t)r & vmask;
  };
  uint4* const maps16 = reinterpret_cast<uint4*>(S.map_a);
  const uint4 zeros16 = make_uint4(0u, 0u, 0u, 0u);
  auto clear_maps = [&]() {
#pragma unroll
    for (int i = (int)lane; i < 2 * MAPW / 4; i += 32) maps16[i] = zeros16;
  };
  clear_maps();
  for (int i = (int)lane; i < kSlots; i += 32) { S.table[i] = ~0ull; S.marked[i] = 0; }
  __syncwarp();
  const uint32_t n_warps = gridDim.x * kFwWarps;
  uint32_t lo, n, lo_next, n_next;
  leaves.get(blockIdx.x * kFwWarps + wib, n_leaf, lo, n);
  for (uint32_t leaf = blockIdx.x * kFwWarps + wib; leaf < n_leaf; leaf += n_warps, lo = lo_next, n = n_next) {
    leaves.get(leaf + n_warps, n_leaf, lo_next, n_next);         // the next leaf's extent: its load latency hides behind this leaf
    if (n < 2) continue;                                         // uniform across the warp
    const uint64_t* src = recs + lo;
    const uint32_t skip = lo & 1u;                               // an odd start (exact offsets): the pair grid begins one record earlier
    const uint32_t n_tot = n + skip;                             // record slots of the pair grid, the first `skip` of them not ours
    auto whole_leaf = [&]() {                                    // hand the leaf to the global sort: always correct, only slower
      unsigned long long base = 0;
      if (lane == 0) { base = atomicAdd(n_cand, (unsigned long long)n); atomicAdd(n_oversized, 1u); *unsorted = 1u; }
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < n; i += 32) emit(base + i, __ldg(src + i), leaf);
      __syncwarp();
    };
    if (n_tot > (uint32_t)kB * 128u) { whole_leaf(); continue; }
    const uint4* const pairs = reinterpret_cast<const uint4*>(src - skip) + lane;     // this lane's first pair
    const uint32_t n_pairs = (n_tot + 1u) >> 1;
    const uint32_t n_batches = (n_tot + 127u) >> 7;
    // pair q of batch b of this lane: index b * 64 + q * 32 + lane of the pair grid = record slots 2 * that and + 1.
    // x[b * 4 + q * 2 + h]: the bits that tell record h of that pair from the other k-mers of the leaf
    uint32_t x[kB * 4];
#pragma unroll
    for (int b = 0; b < kB; ++b) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 r = zeros16;
        if ((uint32_t)(b * 64 + q * 32) + lane < n_pairs) r = __ldg(pairs + (b * 64 + q * 32));
        x[b * 4 + q * 2] = disc(r.x, r.y);
        x[b * 4 + q * 2 + 1] = disc(r.z, r.w);
      }
    }
    auto is_full = [&](int b) -> bool { return (uint32_t)(b + 1) * 128u <= n_tot && (b != 0 || skip == 0u); };
    auto slot_ok = [&](int b, int q, int h) -> bool {             // is record h of pair q of batch b one of the leaf's?
      const uint32_t g = ((uint32_t)(b * 64 + q * 32) + lane) * 2u + (uint32_t)h;
      return g - skip < n;                                       // unsigned: also false for g < skip
    };
    // pass 1: every record sets its bit in map A; a record that finds it set also sets it in map B (which follows map A)
    auto mark = [&](uint32_t xv) {
      const uint32_t bit = (xv * 0x9E3779B1u) >> (32 - kLogBits);
      uint32_t* const wa = S.map_a + (bit >> 5);
      const uint32_t m = 1u << (bit & 31u);
      if (atomicOr(wa, m) & m) atomicOr(wa + MAPW, m);
    };
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      if ((uint32_t)b < n_batches) {                             // uniform
        if (is_full(b)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mark(x[b * 4 + j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (slot_ok(b, j >> 1, j & 1)) mark(x[b * 4 + j]);
        }
      }
    }
    __syncwarp();
    // pass 2: suspects = records whose bit is set in map B (both records of every shared k-mer, plus the few that share a bit);
    // bit (b * 4 + q * 2 + h) of this lane's mask
    auto probe = [&](uint32_t xv) -> uint32_t {
      const uint32_t bit = (xv * 0x9E3779B1u) >> (32 - kLogBits);
      return (S.map_b[bit >> 5] >> (bit & 31u)) & 1u;
    };
    uint32_t sus = 0;
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      if ((uint32_t)b < n_batches) {
        if (is_full(b)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) sus |= probe(x[b * 4 + j]) << (b * 4 + j);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (slot_ok(b, j >> 1, j & 1)) sus |= probe(x[b * 4 + j]) << (b * 4 + j);
        }
      }
    }
    __syncwarp();
    clear_maps();                                                // last use of the maps: clean for the next leaf
    // the suspects of all lanes, compacted: one inclusive warp scan of the per-lane counts
    const uint32_t mine = (uint32_t)__popc(sus);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane >= d) incl += o;
    }
    const uint32_t c1 = __shfl_sync(0xffffffffu, incl, 31);      // uniform
    __syncwarp();                                                // (no lane may set a bit of the next leaf in a word another lane still clears)
    if (c1 == 0) continue;
    if (c1 > (uint32_t)SUSP) { whole_leaf(); continue; }         // heavy sharing
    {
      uint32_t pos = incl - mine;
      const uint64_t* const slots = reinterpret_cast<const uint64_t*>(pairs);     // record slot g of this lane's pairs: slots[g]
      while (sus != 0) {
        const uint32_t sbit = (uint32_t)__ffs((int)sus) - 1u;
        sus &= sus - 1u;
        S.c_recs[pos++] = __ldg(slots + (size_t)((sbit >> 2) * 64u + ((sbit >> 1) & 1u) * 32u) * 2u + (sbit & 1u));
      }
    }
    __syncwarp();
    // exact test on the suspects: lane l owns the suspects l, l + 32, ...
    uint32_t slot[SUSP / 32];
#pragma unroll
    for (int t = 0; t < SUSP / 32; ++t) {
      const uint32_t s = (uint32_t)t * 32u + lane;
      slot[t] = 0xffffffffu;
      if (s < c1) {
        const uint64_t rec = S.c_recs[s];
        const uint64_t ident = rec >> vbits;
        uint32_t h = ((disc((uint32_t)rec, (uint32_t)(rec >> 32)) * 0x85EBCA6Bu) >> 12) % (uint32_t)kSlots;
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
      // order by the packed word (= k-mer hash, part, window) by counting, emit as one contiguous run; the run's place is claimed
      // first and looked at last
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(n_cand, (unsigned long long)c2);
      uint64_t ck[SUSP / 32];
      uint32_t rk[SUSP / 32];
#pragma unroll
      for (int t = 0; t < SUSP / 32; ++t) {
        const uint32_t i = (uint32_t)t * 32u + lane;
        ck[t] = 0; rk[t] = 0;
        if (i < c2) {
          ck[t] = S.o_recs[i];
          for (uint32_t j = 0; j < c2; ++j) {
            const uint64_t ok = S.o_recs[j];
            rk[t] += (ok < ck[t] || (ok == ck[t] && j < i)) ? 1u : 0u;
          }
        }
      }
      base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
      for (int t = 0; t < SUSP / 32; ++t)
        if ((uint32_t)t * 32u + lane < c2) emit(base + rk[t], ck[t], leaf);
    }
    __syncwarp();
  }
}

}  // namespace nrp
