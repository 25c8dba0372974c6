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


// ---- second version: fewer instructions per record ------------------------------------------------------------------
// ncu of k_filter_warp on config 2 (profiles/ncu_full_r02o.txt, ncu_lines): 92 thread instructions per record at 72 % issue
// utilisation -- the kernel had become issue bound, and a third of it was bookkeeping around the loads (an 8-byte load with its own
// index, bounds check and zero fill per record), the run-time choice of where the k-mer bits sit in the word, and a ballot with its
// branch per record in pass 2.  Here:
//   * records are loaded as PAIRS (one 16-byte load per two records; leaf regions start at multiples of four records), two pairs
//     per lane and batch of 128 records, the next batch's loads in flight; a batch that lies wholly inside the leaf takes a path
//     without any per-record check;
//   * the bits that tell the k-mers of a leaf apart are "the 32 bits of the word from bit vbits on" -- one shift (VHI: of the
//     high word; else a funnel shift), no mask: whatever else those 32 bits hold is the same for every record of the leaf;
//   * pass 2 gathers a lane's suspects as bits of a 64-bit mask (slot = batch * 4 + record of the batch); the records behind
//     the set bits are fetched again (L1) after one warp scan of the counts.  A leaf of more than 2048 records is passed on whole
//     (the planner aims at <= 800; fixed regions end at ~1200).
template <typename KeyT, int MAPW, int SUSP, bool VHI>
__global__ void __launch_bounds__(kFwThreads) k_filter_warp2(const uint64_t* __restrict__ recs, LeafExtents leaves, uint32_t n_leaf, int leaf_bits,
                                                            PackFmt pf, KeyT* cand_keys, uint32_t* cand_vals, unsigned long long* n_cand,
                                                            uint32_t* n_oversized, uint32_t* unsorted) {
  static_assert((MAPW & (MAPW - 1)) == 0 && MAPW >= 64 && SUSP % 32 == 0, "sizes");
  constexpr int kLogBits = MAPW == 256 ? 13 : MAPW == 512 ? 14 : MAPW == 1024 ? 15 : -1;
  constexpr int kSlots = 2 * SUSP;
  constexpr uint32_t kMaxLeaf = 2048;                            // 16 batches of 128 records: 64 mask bits per lane
  static_assert(kLogBits > 0, "MAPW must be 256, 512 or 1024");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  FwWarp<MAPW, SUSP>& S = reinterpret_cast<FwWarp<MAPW, SUSP>*>(smem_raw)[wib];
  const uint32_t lt = lanemask_lt();
  const int vbits = pf.vbits;
  const uint32_t vmask = (uint32_t)((1ull << vbits) - 1ull);
  const int xsh = VHI ? vbits - 32 : vbits;
  auto disc = [&](uint32_t lo, uint32_t hi) -> uint32_t { return VHI ? hi >> xsh : __funnelshift_r(lo, hi, xsh); };
  auto bit_of = [&](uint32_t x) -> uint32_t { return (x * 0x9E3779B1u) >> (32 - kLogBits); };
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
    if (n_tot > kMaxLeaf) { whole_leaf(); continue; }
    const uint4* const pairs = reinterpret_cast<const uint4*>(src - skip) + lane;     // this lane's first pair
    const uint32_t n_pairs = (n_tot + 1u) >> 1;
    const uint32_t n_batches = (n_tot + 127u) >> 7;
    // pair q of batch b of this lane: index b * 64 + q * 32 + lane of the pair grid = record slots 2 * that and + 1
    auto load2 = [&](uint32_t b, uint4 (&r)[2]) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (b * 64u + q * 32u + lane < n_pairs) r[q] = __ldg(pairs + b * 64u + q * 32u);
    };
    auto is_full = [&](uint32_t b) -> bool { return (b + 1u) * 128u <= n_tot && (b != 0u || skip == 0u); };
    auto slot_ok = [&](uint32_t b, int q, int h) -> bool {        // is record h of pair q of batch b one of the leaf's?
      const uint32_t g = (b * 64u + (uint32_t)q * 32u + lane) * 2u + (uint32_t)h;
      return g - skip < n;                                       // unsigned: also false for g < skip
    };
    // pass 1: every record sets its bit in map A; a record that finds it set also sets it in map B
    auto mark = [&](uint32_t lo32, uint32_t hi32) {
      const uint32_t bit = bit_of(disc(lo32, hi32));
      uint32_t* const wa = S.map_a + (bit >> 5);
      const uint32_t m = 1u << (bit & 31u);
      if (atomicOr(wa, m) & m) atomicOr(wa + MAPW, m);           // map_b follows map_a
    };
    uint4 cur[2] = {zeros16, zeros16}, nxt[2] = {zeros16, zeros16};
    load2(0, cur);
    for (uint32_t b = 0; b < n_batches; ++b) {
      load2(b + 1u, nxt);                                        // past the end: no load is issued
      if (is_full(b)) {
#pragma unroll
        for (int q = 0; q < 2; ++q) { mark(cur[q].x, cur[q].y); mark(cur[q].z, cur[q].w); }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (slot_ok(b, q, 0)) mark(cur[q].x, cur[q].y);
          if (slot_ok(b, q, 1)) mark(cur[q].z, cur[q].w);
        }
      }
      cur[0] = nxt[0]; cur[1] = nxt[1];
    }
    load2(0, cur);                                               // pass 2 reads the leaf again (L1 / L2): issued before the warp meets
    __syncwarp();
    // pass 2: suspects = records whose bit is set in map B (both records of every shared k-mer, plus the few that share a bit)
    auto probe = [&](uint32_t lo32, uint32_t hi32) -> uint32_t {
      const uint32_t bit = bit_of(disc(lo32, hi32));
      return (S.map_b[bit >> 5] >> (bit & 31u)) & 1u;
    };
    uint32_t sus_lo = 0, sus_hi = 0;                             // bit (b * 4 + q * 2 + h) of the 64-bit mask (sus_hi, sus_lo)
    for (uint32_t b = 0; b < n_batches; ++b) {
      load2(b + 1u, nxt);
      uint32_t four;
      if (is_full(b)) {
        four = probe(cur[0].x, cur[0].y) | (probe(cur[0].z, cur[0].w) << 1) | (probe(cur[1].x, cur[1].y) << 2) | (probe(cur[1].z, cur[1].w) << 3);
      } else {
        four = 0;
        if (slot_ok(b, 0, 0)) four |= probe(cur[0].x, cur[0].y);
        if (slot_ok(b, 0, 1)) four |= probe(cur[0].z, cur[0].w) << 1;
        if (slot_ok(b, 1, 0)) four |= probe(cur[1].x, cur[1].y) << 2;
        if (slot_ok(b, 1, 1)) four |= probe(cur[1].z, cur[1].w) << 3;
      }
      if (b < 8u) sus_lo |= four << (4u * b); else sus_hi |= four << (4u * (b - 8u));
      cur[0] = nxt[0]; cur[1] = nxt[1];
    }
    __syncwarp();
    clear_maps();                                                // last use of the maps: clean for the next leaf
    // the suspects of all lanes, compacted: one inclusive warp scan of the per-lane counts
    const uint32_t mine = (uint32_t)__popc(sus_lo) + (uint32_t)__popc(sus_hi);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane >= d) incl += o;
    }
    const uint32_t c1 = __shfl_sync(0xffffffffu, incl, 31);      // uniform
    __syncwarp();                                                // (no lane may set a bit of the next leaf in a word another lane still clears)
    if (c1 == 0) continue;
    if (c1 > (uint32_t)SUSP) { whole_leaf(); continue; }         // heavy sharing (or a leaf far above the planned size)
    {
      uint32_t pos = incl - mine;
      const uint64_t* const slots = reinterpret_cast<const uint64_t*>(pairs);     // record slot g of this lane's pairs: slots[g]
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t m = half ? sus_hi : sus_lo;
        while (m != 0) {
          const uint32_t sbit = (uint32_t)__ffs((int)m) - 1u;
          m &= m - 1u;
          const uint32_t b = (uint32_t)half * 8u + (sbit >> 2), q = (sbit >> 1) & 1u, h = sbit & 1u;
          S.c_recs[pos++] = __ldg(slots + (size_t)(b * 64u + q * 32u) * 2u + h);
        }
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


// ---- third version: the leaf's k-mer bits stay in registers ----------------------------------------------------------------
// ncu of k_filter_warp2 on config 2 (profiles/ncu_full_r02p.txt): 70 thread instructions per record, 60 % issue utilisation, and the
// long-scoreboard stall back on top (5.3 per issue): per leaf a warp walks a chain of dependent round trips -- the first batch from
// HBM, every further batch one step behind, the whole leaf AGAIN for pass 2, the suspects, the global counter.  A leaf of <= 1024
// records is 32 words per lane once only the 32 discriminating bits of each record are kept, so here
//   * ALL loads of the leaf are issued at once (16 predicated 16-byte loads per lane; with VHI the compiler fetches just the
//     high words) and reduced to x[32] in registers;
//   * pass 1 and pass 2 run on registers: nothing is read twice, no load sits between the two passes;
//   * only the suspects' full records are fetched again (L1 / L2);
//   * the claim of the output run (one global atomic) is issued before the ranks are computed and consumed after.
// Leaves of more than 1024 records (the fixed regions of a plan for <= 800 end at ~1200: nine sigma) are passed on whole.
template <typename KeyT, int MAPW, int SUSP, bool VHI>
__global__ void __launch_bounds__(kFwThreads) k_filter_warp3(const uint64_t* __restrict__ recs, LeafExtents leaves, uint32_t n_leaf, int leaf_bits,
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
