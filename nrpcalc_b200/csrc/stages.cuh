// Kernels around the bulk path: per-part exclusion flags, background table, grouping of the sorted candidate
// records, group ranks / last single windows for the order-exact host replay, and part-pair edge expansion.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "scan.cuh"

namespace nrp {

// ---- background hash set (replaces the kmerSetDB lookups of reference kmerSetDB.py:119-127,177-207) ------
// Open addressing, linear probing, EMPTY = all ones (never canonical).  Sized >= 2n and a power of two.
template <typename KeyT>
__global__ void k_bg_insert(const uint64_t* keys, int64_t n, KeyT* table, int log_slots) {
  const KeyT kEmpty = ~KeyT(0);
  const uint64_t mask = (1ull << log_slots) - 1ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    KeyT key = (KeyT)keys[i];
    uint64_t h = key_hash((uint64_t)key) >> (64 - log_slots);
    for (;;) {
      KeyT old = atomic_cas(&table[h], kEmpty, key);
      if (old == kEmpty || old == key) break;
      h = (h + 1) & mask;
    }
  }
}

// ---- background build on the device (replaces the per-k-mer puts of reference kmerSetDB.py:129-175) --------------
// canonical K-mers of every K-window of the staged sequences, appended in tile order (the sort that follows orders them)
__global__ void __launch_bounds__(kTileThreads) k_extract_canon(Parts P, int K, uint64_t* out, unsigned long long* n_out) {
  __shared__ TileStage stage;
  __shared__ uint64_t s_keys[kTileBytes];
  __shared__ uint32_t s_count;
  __shared__ unsigned long long s_base;
  const int64_t tile = P.tile_lo + blockIdx.x;
  if (threadIdx.x == 0) s_count = 0;
  tile_load(P, tile, stage, threadIdx.x);
  __syncthreads();
  tile_windows_blocked<false>(P, stage, tile, K, nullptr, threadIdx.x, [&](int, uint64_t canon, uint32_t, uint32_t, bool) {
    s_keys[atomicAdd(&s_count, 1u)] = canon;
  });
  __syncthreads();
  const uint32_t total = s_count;
  if (threadIdx.x == 0 && total) s_base = atomicAdd(n_out, (unsigned long long)total);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) out[s_base + i] = s_keys[i];
}

// distinct keys of a sorted array: head marks + compaction
__global__ void k_unique_heads(const uint64_t* keys, int64_t n, uint32_t* head) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void k_unique_compact(const uint64_t* keys, const uint32_t* head, const uint32_t* head_scan, int64_t n, uint64_t* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && head[i]) out[head_scan[i]] = keys[i];
}

template <typename KeyT>
__device__ __forceinline__ bool bg_contains(const KeyT* table, int log_slots, uint64_t key64) {
  const KeyT kEmpty = ~KeyT(0);
  const KeyT key = (KeyT)key64;
  const uint64_t mask = (1ull << log_slots) - 1ull;
  uint64_t h = key_hash(key64) >> (64 - log_slots);
  for (;;) {
    KeyT cur = __ldg(&table[h]);
    if (cur == key) return true;
    if (cur == kEmpty) return false;
    h = (h + 1) & mask;
  }
}

// ---- per-part exclusion flags that do not need grouping (reference hgraph.py:59-79) ----------------------
// mode 0: palindromic k-windows (a window equal to its reverse complement, even k only)
// mode 1: background hit of a canonical K-window (window length = the background's K)
template <typename BgKeyT>
__global__ void __launch_bounds__(kTileThreads) k_flag_windows(Parts P, int klen, int mode, const BgKeyT* bg_table, int bg_log_slots,
                                                              uint8_t* flags) {
  __shared__ TileStage stage;
  const int64_t tile = P.tile_lo + blockIdx.x;
  tile_load(P, tile, stage, threadIdx.x);
  __syncthreads();
  tile_windows_blocked<false>(P, stage, tile, klen, flags, threadIdx.x, [&](int, uint64_t canon, uint32_t part, uint32_t, bool palindrome) {
    if (mode == 0) {
      if (palindrome) flags[part] |= kFlagPalindrome;             // benign race: all writers OR the same bit
    } else {
      if (!(flags[part] & kFlagBackground) && bg_contains<BgKeyT>(bg_table, bg_log_slots, canon)) flags[part] |= kFlagBackground;
    }
  });
}

// ---- grouping of the sorted candidate records -------------------------------------------------------------
// records sorted by (key, part).  head[i] = 1 where a new key starts.  Two equal adjacent records (same key,
// same part) mean that the part has an internal repeat (reference hgraph.py:53-69).
template <typename KeyT>
__global__ void k_group_heads(const KeyT* keys, const uint32_t* vals, int64_t n, int j_bits, uint32_t* head, uint8_t* flags,
                              unsigned long long* n_repeat_records) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool is_head = i == 0 || keys[i] != keys[i - 1];
  head[i] = is_head ? 1u : 0u;
  if (!is_head && (vals[i] >> j_bits) == (vals[i - 1] >> j_bits)) {
    if (flags) flags[vals[i] >> j_bits] |= kFlagRepeat;
    atomicAdd(n_repeat_records, 1ull);
  }
}

// records of parts flagged for an internal repeat must vanish from every group (the reference drops such a part
// before it touches the table, hgraph.py:50-69): keep[i] = 1 unless the record's part carries the repeat flag.
// flags_global is indexed by the record value itself (global part index).
__global__ void k_keep_unflagged(const uint32_t* vals, int64_t n, int j_bits, const uint8_t* flags_by_val, uint32_t* keep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = (flags_by_val[vals[i] >> j_bits] & kFlagRepeat) ? 0u : 1u;
}

template <typename KeyT>
__global__ void k_compact_records(const KeyT* keys, const uint32_t* vals, const uint32_t* keep, const uint32_t* keep_scan, int64_t n,
                                  KeyT* out_keys, uint32_t* out_vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  uint32_t o = keep_scan[i];
  out_keys[o] = keys[i];
  out_vals[o] = vals[i];
}

// windows of the parts that carry the repeat flag (they were counted as records before the flag was known)
__global__ void k_flagged_windows(const int64_t* off, const uint8_t* flags, int64_t lo, int64_t hi, int k, unsigned long long* out) {
  int64_t p = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hi || !(flags[p] & kFlagRepeat)) return;
  int64_t w = off[p + 1] - off[p] - k + 1;
  if (w > 0) atomicAdd(out, (unsigned long long)w);
}

// ---- key runs -> groups, with the record / run counts read from device memory ---------------------------------
// The grouping stage is enqueued without host round trips: grids are sized by the upper bound n_max (the number of
// candidates) and every kernel reads the actual count from a device word; slots beyond it are written as zero so that
// the prefix sums over n_max entries give the right totals.
template <typename KeyT>
__global__ void k_group_heads_n(const KeyT* keys, const uint32_t* vals, const uint32_t* n_ptr, int64_t n_max, uint32_t* head) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_max) return;
  head[i] = (i < (int64_t)*n_ptr && (i == 0 || keys[i] != keys[i - 1])) ? 1u : 0u;
}

__global__ void k_group_starts_n(const uint32_t* head, const uint32_t* head_scan, const uint32_t* n_ptr, int64_t n_max, uint32_t* group_start) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)*n_ptr;
  if (i > n) return;
  if (i == n) { group_start[head_scan[n_max]] = (uint32_t)n; return; }
  if (head[i]) group_start[head_scan[i]] = (uint32_t)i;
}

// Fused: per key run its validity (>= 2 records), member count and pair count, and the exclusive prefix sums of all three
// over n_max run slots (zeros beyond the actual number of runs) -- three launches instead of one kernel + three scans.
//   pass 1  per tile of kScanTile runs: local exclusive scans, tile totals
//   pass 2  one CTA: exclusive scan of the tile totals (+ grand totals at index n_max of the outputs)
//   pass 3  add the tile offsets
struct GroupTotals { uint32_t valid; unsigned long long members, pairs; };

__device__ __forceinline__ void run_size_values(const uint32_t* group_start, int64_t g, int64_t n_runs, uint32_t& v, unsigned long long& m,
                                                unsigned long long& p) {
  const uint32_t sz = g < n_runs ? group_start[g + 1] - group_start[g] : 0u;
  const bool ok = sz >= 2;
  v = ok ? 1u : 0u;
  m = ok ? sz : 0ull;
  p = ok ? (unsigned long long)sz * (sz - 1) / 2 : 0ull;
}

__global__ void __launch_bounds__(kScanThreads) k_sizes_scan_tiles(const uint32_t* group_start, const uint32_t* n_runs_ptr, int64_t n_max,
                                                                  uint32_t* valid, uint32_t* valid_scan, unsigned long long* mem_scan,
                                                                  unsigned long long* pair_scan, GroupTotals* tile_totals) {
  __shared__ uint32_t ws32[33];
  __shared__ unsigned long long ws64[33];
  const int64_t n_runs = (int64_t)*n_runs_ptr;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems]; unsigned long long m[kScanItems], p[kScanItems];
  uint32_t sv = 0; unsigned long long sm = 0, sp = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    run_size_values(group_start, base + i, n_runs, v[i], m[i], p[i]);
    if (base + i >= n_max) { v[i] = 0; m[i] = 0; p[i] = 0; }
    sv += v[i]; sm += m[i]; sp += p[i];
  }
  uint32_t tv; unsigned long long tm, tp;
  uint32_t rv = block_exclusive_scan<uint32_t>(sv, ws32, tv);
  unsigned long long rm = block_exclusive_scan<unsigned long long>(sm, ws64, tm);
  unsigned long long rp = block_exclusive_scan<unsigned long long>(sp, ws64, tp);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t g = base + i;
    if (g < n_max) { valid[g] = v[i]; valid_scan[g] = rv; mem_scan[g] = rm; pair_scan[g] = rp; }
    rv += v[i]; rm += m[i]; rp += p[i];
  }
  if (threadIdx.x == 0) tile_totals[blockIdx.x] = GroupTotals{tv, tm, tp};
}

__global__ void __launch_bounds__(kScanThreads) k_sizes_scan_top(GroupTotals* tile_totals, int64_t n_tiles, int64_t n_max, uint32_t* valid_scan,
                                                                unsigned long long* mem_scan, unsigned long long* pair_scan) {
  __shared__ uint32_t ws32[33];
  __shared__ unsigned long long ws64[33];
  const int64_t per = (n_tiles + kScanThreads - 1) / kScanThreads;
  const int64_t t0 = (int64_t)threadIdx.x * per, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
  uint32_t sv = 0; unsigned long long sm = 0, sp = 0;
  for (int64_t t = t0; t < t1; ++t) { sv += tile_totals[t].valid; sm += tile_totals[t].members; sp += tile_totals[t].pairs; }
  uint32_t tv; unsigned long long tm, tp;
  uint32_t rv = block_exclusive_scan<uint32_t>(sv, ws32, tv);
  unsigned long long rm = block_exclusive_scan<unsigned long long>(sm, ws64, tm);
  unsigned long long rp = block_exclusive_scan<unsigned long long>(sp, ws64, tp);
  for (int64_t t = t0; t < t1; ++t) {
    const GroupTotals cur = tile_totals[t];
    tile_totals[t] = GroupTotals{rv, rm, rp};          // becomes the tile's offset
    rv += cur.valid; rm += cur.members; rp += cur.pairs;
  }
  if (threadIdx.x == 0) { valid_scan[n_max] = tv; mem_scan[n_max] = tm; pair_scan[n_max] = tp; }
}

__global__ void __launch_bounds__(kScanThreads) k_sizes_scan_add(const GroupTotals* tile_offsets, int64_t n_max, uint32_t* valid_scan,
                                                                unsigned long long* mem_scan, unsigned long long* pair_scan) {
  const GroupTotals add = tile_offsets[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
    const int64_t g = base + i;
    if (g < n_max) { valid_scan[g] += add.valid; mem_scan[g] += add.members; pair_scan[g] += add.pairs; }
  }
}

template <typename KeyT>
__global__ void k_group_emit_n(const KeyT* keys, const uint32_t* vals, int j_bits, const uint32_t* group_start, const uint32_t* valid,
                               const uint32_t* valid_scan, const unsigned long long* member_scan, const unsigned long long* pair_scan,
                               const uint32_t* n_runs_ptr, int64_t n_max, uint64_t* out_keys, int64_t* out_indptr, unsigned long long* out_pair_off,
                               uint32_t* out_first_window) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_runs = (int64_t)*n_runs_ptr;
  if (g > n_runs) return;
  if (g == n_runs) {                       // totals: the scans ran over n_max entries, zeros beyond n_runs
    uint32_t ng = valid_scan[n_max];
    out_indptr[ng] = (int64_t)member_scan[n_max];
    out_pair_off[ng] = pair_scan[n_max];
    return;
  }
  if (!valid[g]) return;
  uint32_t o = valid_scan[g];
  out_keys[o] = (uint64_t)keys[group_start[g]];
  if (j_bits) out_first_window[o] = vals[group_start[g]] & ((1u << j_bits) - 1u);     // the first record is (first part, first window)
  out_indptr[o] = (int64_t)member_scan[g];
  out_pair_off[o] = pair_scan[g];
}

__global__ void k_group_members_n(const uint32_t* vals, const uint32_t* head_scan, const uint32_t* head, const uint32_t* group_start,
                                  const uint32_t* valid, const unsigned long long* member_scan, const uint32_t* n_ptr, int j_bits,
                                  uint32_t* out_members, uint32_t* out_member_end) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)*n_ptr) return;
  uint32_t g = head_scan[i] + head[i] - 1u;           // run index of record i (inclusive scan - 1)
  if (!valid[g]) return;
  const unsigned long long pos = member_scan[g] + (i - group_start[g]);
  out_members[pos] = vals[i] >> j_bits;
  // where this member's group ends in the member list (the edge expansion needs no search for it)
  out_member_end[pos] = (uint32_t)(member_scan[g] + (group_start[g + 1] - group_start[g]));
}

// ---- the whole grouping stage as ONE cooperative kernel ---------------------------------------------------------
// Same results as the kernel sequence above (k_keep_unflagged .. k_group_members_n): the phases are separated by grid-wide
// barriers instead of kernel boundaries, prefix sums are tile-local + a scan of the tile totals by CTA 0, and consumers add
// the tile offset on the fly.  ~9 grid barriers instead of ~20 launches: the stage is latency bound, not work bound.
struct GroupCoopArgs {
  const void* keys; const uint32_t* vals;          // ordered candidates
  void* keys2; uint32_t* vals2;                     // compaction target (used when drop_flagged)
  int64_t n; int j_bits; int drop_flagged; const uint8_t* flags;
  uint32_t* head; uint32_t* head_scan; uint32_t* group_start;
  uint32_t* valid; uint32_t* valid_scan; unsigned long long* mem_scan; unsigned long long* pair_scan;
  uint32_t* tile_u32; GroupTotals* tile_tot;        // per-tile totals -> offsets (n_tiles + 1 entries each)
  uint64_t* out_keys; int64_t* out_indptr; unsigned long long* out_pair_off; uint32_t* out_first_window;
  uint32_t* out_members; uint32_t* out_member_end;
  unsigned long long* counts;                       // [0] records kept [1] key runs [2] groups [3] members [4] pairs
};

// exclusive scan of n_tiles per-tile totals in place by one CTA; returns the grand total to every thread
__device__ __forceinline__ uint32_t cta_scan_tiles_u32(uint32_t* t, int64_t n_tiles, uint32_t* ws) {
  const int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
  const int64_t t0 = (int64_t)threadIdx.x * per, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
  uint32_t sum = 0;
  for (int64_t i = t0; i < t1; ++i) sum += t[i];
  uint32_t total;
  uint32_t run = block_exclusive_scan<uint32_t>(sum, ws, total);
  for (int64_t i = t0; i < t1; ++i) { const uint32_t v = t[i]; t[i] = run; run += v; }
  return total;
}

template <typename KeyT>
__global__ void __launch_bounds__(kScanThreads) k_group_coop(GroupCoopArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t ws32[33];
  __shared__ unsigned long long ws64[33];
  const int64_t n = a.n;
  const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
  const KeyT* K = reinterpret_cast<const KeyT*>(a.keys);
  const uint32_t* V = a.vals;
  int64_t n_kept = n;

  if (a.drop_flagged) {
    // A1: keep[i] = the record's part does not carry the repeat flag; tile-local scan, tile totals
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int64_t base = t * kScanTile + (int64_t)threadIdx.x * kScanItems;
      uint32_t keep[kScanItems], sum = 0;
#pragma unroll
      for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + i;
        keep[i] = (idx < n && !(a.flags[V[idx] >> a.j_bits] & kFlagRepeat)) ? 1u : 0u;
        sum += keep[i];
      }
      uint32_t total;
      uint32_t run = block_exclusive_scan<uint32_t>(sum, ws32, total);
#pragma unroll
      for (int i = 0; i < kScanItems; ++i) {
        const int64_t idx = base + i;
        if (idx < n) { a.valid[idx] = keep[i]; a.valid_scan[idx] = run; }
        run += keep[i];
      }
      if (threadIdx.x == 0) a.tile_u32[t] = total;
    }
    grid.sync();
    if (blockIdx.x == 0) {
      const uint32_t total = cta_scan_tiles_u32(a.tile_u32, n_tiles, ws32);
      if (threadIdx.x == 0) { a.tile_u32[n_tiles] = total; a.counts[0] = total; }
    }
    grid.sync();
    n_kept = (int64_t)a.tile_u32[n_tiles];
    // A2: compaction (order preserved)
    KeyT* K2 = reinterpret_cast<KeyT*>(a.keys2);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
      if (a.valid[idx]) {
        const uint32_t dst = a.tile_u32[idx / kScanTile] + a.valid_scan[idx];
        K2[dst] = K[idx];
        a.vals2[dst] = V[idx];
      }
    }
    grid.sync();
    K = K2; V = a.vals2;
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.counts[0] = (unsigned long long)n;
  }

  // B: key-run heads of the kept records, tile-local scan, tile totals
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t base = t * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t h[kScanItems], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = base + i;
      h[i] = (idx < n_kept && (idx == 0 || K[idx] != K[idx - 1])) ? 1u : 0u;
      sum += h[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan<uint32_t>(sum, ws32, total);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = base + i;
      if (idx < n) { a.head[idx] = h[i]; a.head_scan[idx] = run; }
      run += h[i];
    }
    if (threadIdx.x == 0) a.tile_u32[t] = total;
  }
  grid.sync();
  if (blockIdx.x == 0) {
    const uint32_t total = cta_scan_tiles_u32(a.tile_u32, n_tiles, ws32);
    if (threadIdx.x == 0) { a.tile_u32[n_tiles] = total; a.counts[1] = total; }
  }
  grid.sync();
  const int64_t n_runs = (int64_t)a.tile_u32[n_tiles];
  // C: start of every key run (+ the sentinel)
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx <= n_kept; idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx == n_kept) a.group_start[n_runs] = (uint32_t)n_kept;
    else if (a.head[idx]) a.group_start[a.tile_u32[idx / kScanTile] + a.head_scan[idx]] = (uint32_t)idx;
  }
  grid.sync();
  // D: per run validity / members / pairs, tile-local scans, tile totals
  const int64_t r_tiles = (n_runs + kScanTile - 1) / kScanTile;
  for (int64_t t = blockIdx.x; t < r_tiles; t += gridDim.x) {
    const int64_t base = t * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems]; unsigned long long m[kScanItems], p[kScanItems];
    uint32_t sv = 0; unsigned long long sm = 0, sp = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      run_size_values(a.group_start, base + i, n_runs, v[i], m[i], p[i]);
      sv += v[i]; sm += m[i]; sp += p[i];
    }
    uint32_t tv; unsigned long long tm, tp;
    uint32_t rv = block_exclusive_scan<uint32_t>(sv, ws32, tv);
    unsigned long long rm = block_exclusive_scan<unsigned long long>(sm, ws64, tm);
    unsigned long long rp = block_exclusive_scan<unsigned long long>(sp, ws64, tp);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t g = base + i;
      if (g < n_runs) { a.valid[g] = v[i]; a.valid_scan[g] = rv; a.mem_scan[g] = rm; a.pair_scan[g] = rp; }
      rv += v[i]; rm += m[i]; rp += p[i];
    }
    if (threadIdx.x == 0) a.tile_tot[t] = GroupTotals{tv, tm, tp};
  }
  grid.sync();
  if (blockIdx.x == 0) {
    const int64_t per = (r_tiles + blockDim.x - 1) / blockDim.x;
    const int64_t t0 = (int64_t)threadIdx.x * per, t1 = t0 + per < r_tiles ? t0 + per : r_tiles;
    uint32_t sv = 0; unsigned long long sm = 0, sp = 0;
    for (int64_t t = t0; t < t1; ++t) { sv += a.tile_tot[t].valid; sm += a.tile_tot[t].members; sp += a.tile_tot[t].pairs; }
    uint32_t tv; unsigned long long tm, tp;
    uint32_t rv = block_exclusive_scan<uint32_t>(sv, ws32, tv);
    unsigned long long rm = block_exclusive_scan<unsigned long long>(sm, ws64, tm);
    unsigned long long rp = block_exclusive_scan<unsigned long long>(sp, ws64, tp);
    for (int64_t t = t0; t < t1; ++t) {
      const GroupTotals cur = a.tile_tot[t];
      a.tile_tot[t] = GroupTotals{rv, rm, rp};
      rv += cur.valid; rm += cur.members; rp += cur.pairs;
    }
    if (threadIdx.x == 0) {
      a.counts[2] = tv; a.counts[3] = tm; a.counts[4] = tp;
      a.out_indptr[tv] = (int64_t)tm;
      a.out_pair_off[tv] = tp;
    }
  }
  grid.sync();
  // E: groups (keys, CSR offsets, first windows) and their members
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_runs; g += (int64_t)gridDim.x * blockDim.x) {
    if (!a.valid[g]) continue;
    const GroupTotals off = a.tile_tot[g / kScanTile];
    const uint32_t o = off.valid + a.valid_scan[g];
    const uint32_t first = a.group_start[g];
    a.out_keys[o] = (uint64_t)K[first];
    if (a.j_bits) a.out_first_window[o] = V[first] & ((1u << a.j_bits) - 1u);
    a.out_indptr[o] = (int64_t)(off.members + a.mem_scan[g]);
    a.out_pair_off[o] = off.pairs + a.pair_scan[g];
  }
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_kept; idx += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t g = a.tile_u32[idx / kScanTile] + a.head_scan[idx] + a.head[idx] - 1u;      // run of record idx
    if (!a.valid[g]) continue;
    const uint32_t first = a.group_start[g], size = a.group_start[g + 1] - first;
    const unsigned long long gbase = a.tile_tot[g / kScanTile].members + a.mem_scan[g];
    const unsigned long long pos = gbase + (idx - first);
    a.out_members[pos] = V[idx] >> a.j_bits;
    a.out_member_end[pos] = (uint32_t)(gbase + size);
  }
}

// ---- the grouping stage with TWO grid barriers ("walk" variant) ---------------------------------------------------
// Same results as k_group_coop.  The records of flagged parts are not compacted away: a key run of the ordered candidate list is
// contiguous with or without them, so the thread that holds a run's first record WALKS the run (two records on average),
// counts the records that stay, and contributes (group?, members, pairs) to ONE set of prefix sums over the record slots; after
// the scan of the tile totals the same threads walk their runs again and write groups and members.  A walk longer than
// kGroupWalkMax gives up and raises a flag: the host then runs k_group_coop (pathological inputs: one k-mer shared by
// thousands of records).  counts: [0] records kept [1] key runs with a kept record [2] groups [3] members [4] pairs [5] gave up.
constexpr int kGroupWalkMax = 256;

template <typename KeyT>
__global__ void __launch_bounds__(kScanThreads) k_group_walk(GroupCoopArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ uint32_t ws32[33];
  __shared__ unsigned long long ws64[33];
  const int64_t n = a.n;
  const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
  const KeyT* K = reinterpret_cast<const KeyT*>(a.keys);
  const uint32_t* V = a.vals;
  const bool drop = a.drop_flagged != 0;
  auto kept = [&](int64_t idx) -> bool { return !drop || !(a.flags[V[idx] >> a.j_bits] & kFlagRepeat); };
  unsigned long long my_kept = 0, my_runs = 0;
  // pass 1: per run head its kept size; tile-local exclusive scans of (group?, members, pairs) over the record slots
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t base = t * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems]; unsigned long long m[kScanItems], p[kScanItems];
    uint32_t sv = 0; unsigned long long sm = 0, sp = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = base + i;
      v[i] = 0; m[i] = 0; p[i] = 0;
      if (idx < n && (idx == 0 || K[idx] != K[idx - 1])) {
        const KeyT key = K[idx];
        uint32_t size = 0;
        int64_t j = idx;
        for (; j < n && K[j] == key; ++j) {
          if (j - idx >= kGroupWalkMax) { a.counts[5] = 1ull; break; }
          size += kept(j) ? 1u : 0u;
        }
        my_kept += size;
        my_runs += size ? 1u : 0u;
        if (size >= 2) { v[i] = 1u; m[i] = size; p[i] = (unsigned long long)size * (size - 1) / 2; }
      }
      sv += v[i]; sm += m[i]; sp += p[i];
    }
    uint32_t tv; unsigned long long tm, tp;
    uint32_t rv = block_exclusive_scan<uint32_t>(sv, ws32, tv);
    unsigned long long rm = block_exclusive_scan<unsigned long long>(sm, ws64, tm);
    unsigned long long rp = block_exclusive_scan<unsigned long long>(sp, ws64, tp);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = base + i;
      if (idx < n && v[i]) { a.valid_scan[idx] = rv; a.mem_scan[idx] = rm; a.pair_scan[idx] = rp; }
      if (idx < n) a.valid[idx] = v[i];
      rv += v[i]; rm += m[i]; rp += p[i];
    }
    if (threadIdx.x == 0) a.tile_tot[t] = GroupTotals{tv, tm, tp};
  }
  // the two plain totals: one atomic per warp
  for (int d = 16; d > 0; d >>= 1) { my_kept += __shfl_down_sync(0xffffffffu, my_kept, d); my_runs += __shfl_down_sync(0xffffffffu, my_runs, d); }
  if ((threadIdx.x & 31) == 0) { if (my_kept) atomicAdd(&a.counts[0], my_kept); if (my_runs) atomicAdd(&a.counts[1], my_runs); }
  grid.sync();
  if (blockIdx.x == 0) {
    const int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    const int64_t t0 = (int64_t)threadIdx.x * per, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
    uint32_t sv = 0; unsigned long long sm = 0, sp = 0;
    for (int64_t t = t0; t < t1; ++t) { sv += a.tile_tot[t].valid; sm += a.tile_tot[t].members; sp += a.tile_tot[t].pairs; }
    uint32_t tv; unsigned long long tm, tp;
    uint32_t rv = block_exclusive_scan<uint32_t>(sv, ws32, tv);
    unsigned long long rm = block_exclusive_scan<unsigned long long>(sm, ws64, tm);
    unsigned long long rp = block_exclusive_scan<unsigned long long>(sp, ws64, tp);
    for (int64_t t = t0; t < t1; ++t) {
      const GroupTotals cur = a.tile_tot[t];
      a.tile_tot[t] = GroupTotals{rv, rm, rp};
      rv += cur.valid; rm += cur.members; rp += cur.pairs;
    }
    if (threadIdx.x == 0) {
      a.counts[2] = tv; a.counts[3] = tm; a.counts[4] = tp;
      a.out_indptr[tv] = (int64_t)tm;
      a.out_pair_off[tv] = tp;
    }
  }
  grid.sync();
  // pass 2: the run heads write their groups and members
  const uint32_t j_mask = a.j_bits ? ((1u << a.j_bits) - 1u) : 0u;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const GroupTotals off = a.tile_tot[t];
    const int64_t base = t * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      const int64_t idx = base + i;
      if (idx >= n || !a.valid[idx]) continue;
      const uint32_t o = off.valid + a.valid_scan[idx];
      const unsigned long long gbase = off.members + a.mem_scan[idx];
      const KeyT key = K[idx];
      unsigned long long pos = gbase;
      bool first = true;
      int64_t j = idx;
      for (; j < n && K[j] == key; ++j) {
        if (!kept(j)) continue;
        if (first) {                                   // the first kept record is (first part, first window)
          a.out_keys[o] = (uint64_t)key;
          if (a.j_bits) a.out_first_window[o] = V[j] & j_mask;
          first = false;
        }
        a.out_members[pos++] = V[j] >> a.j_bits;
      }
      const uint32_t end = (uint32_t)pos;
      for (unsigned long long q = gbase; q < pos; ++q) a.out_member_end[q] = end;
      a.out_indptr[o] = (int64_t)gbase;
      a.out_pair_off[o] = off.pairs + a.pair_scan[idx];
    }
  }
}

// ---- order information for the host replay ----------------------------------------------------
// canonical k-mer of window j of a part, straight from the ASCII buffer (only used on the few shared keys)
__device__ __forceinline__ uint64_t window_from_ascii(const uint8_t* seq, int k) {
  uint64_t fwd = 0;
  for (int t = 0; t < k; ++t) {
    uint32_t c = seq[t];
    uint32_t x = (c >> 1) & 3u;
    fwd = (fwd << 2) | (x ^ (x >> 1));
  }
  return fwd;
}

// first window of part `first_part[g]` whose canonical k-mer equals the group's key: the rank of the group
// (reference: position of the key's first insertion into repeat_dict, hgraph.py:82-89)
// (shard-only staging: the bytes of part p live on the rank whose shard holds p -- every rank keeps them at the same global
// offset of its own part buffer, mapped here through CUDA IPC, so a foreign part is read over NVLink)
struct PartHomes {
  const uint8_t* base[kMaxPeers];    // part buffer of rank r (the own rank: local memory)
  int64_t first[kMaxPeers + 1];      // rank r holds the parts [first[r], first[r + 1])
  int n;                             // 0: everything is staged locally
};
__global__ void k_group_first_window(const uint8_t* ascii, const int64_t* off, const uint64_t* group_keys, const uint32_t* first_part,
                                     int64_t n_parts, int64_t n_groups, int k, uint32_t* first_window, PartHomes homes) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  int64_t p = (int64_t)first_part[g];
  if (p >= n_parts) return;
  for (int r = 0; r < homes.n; ++r)
    if (p >= homes.first[r] && p < homes.first[r + 1]) ascii = homes.base[r];
  const uint8_t* seq = ascii + off[p];
  const int64_t len = off[p + 1] - off[p];
  const uint64_t key = group_keys[g];
  const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
  uint64_t fwd = window_from_ascii(seq, k);
  uint32_t found = 0xffffffffu;
  for (int64_t j = 0; j + k <= len; ++j) {
    if (j > 0) {
      uint32_t c = seq[j + k - 1];
      uint32_t x = (c >> 1) & 3u;
      fwd = ((fwd << 2) | (x ^ (x >> 1))) & mask;
    }
    uint64_t rc = revcomp2(fwd, k);
    if ((fwd < rc ? fwd : rc) == key) { found = (uint32_t)j; break; }
  }
  first_window[g] = found;
}

// last window of each surviving part whose canonical k-mer is not a shared key: the rank of the part's
// singleton tuple (id,) in the reference's popitem order
__global__ void k_last_single(const uint8_t* ascii, const int64_t* off, const uint8_t* flags, int64_t lo, int64_t hi, int k,
                              const uint64_t* shared_table, int log_slots, int32_t* last_single) {
  int64_t p = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hi) return;
  const int64_t len = off[p + 1] - off[p];
  int32_t result = -1;
  if (len >= k && !(flags && flags[p])) {
    const uint8_t* seq = ascii + off[p];
    for (int64_t j = len - k; j >= 0; --j) {
      uint64_t fwd = window_from_ascii(seq + j, k);
      uint64_t rc = revcomp2(fwd, k);
      uint64_t canon = fwd < rc ? fwd : rc;
      const bool shared = log_slots > 0 && bg_contains<uint64_t>(shared_table, log_slots, canon);   // hash set of the shared k-mers
      if (!shared) { result = (int32_t)j; break; }
    }
  }
  last_single[p] = result;
}

// ---- part-pair edges (reference hgraph.py:157-200, order-free) ---------------------------------------------------
// Distinct (id_u < id_v) pairs through an open-addressing hash set of 64-bit codes (EMPTY = all ones): the thread that
// wins the slot appends the edge to the output list.  Pairs of a part with itself (a k-mer occurring twice in one
// part, internal_repeats=True) are skipped.  The output order is unspecified.
__device__ __forceinline__ void edge_insert(uint32_t ida, uint32_t idb, uint64_t* table, int log_slots, uint32_t* out_u, uint32_t* out_v,
                                            unsigned long long* n_out) {
  if (ida == idb) return;
  const uint32_t lo = ida < idb ? ida : idb, hi = ida < idb ? idb : ida;
  const uint64_t code = ((uint64_t)lo << 32) | hi;
  const uint64_t mask = (1ull << log_slots) - 1ull;
  uint64_t h = key_hash(code) >> (64 - log_slots);
  for (;;) {
    uint64_t old = atomic_cas(&table[h], ~0ull, code);
    if (old == code) return;
    if (old == ~0ull) {
      unsigned long long o = atomicAdd(n_out, 1ull);
      out_u[o] = lo; out_v[o] = hi;
      return;
    }
    h = (h + 1) & mask;
  }
}

// one thread per member a of a group: pairs (a, b > a)
__global__ void k_edge_pairs(const uint32_t* member_end, const uint32_t* members, int64_t n_members, const uint32_t* part_id,
                             uint64_t* table, int log_slots, uint32_t* out_u, uint32_t* out_v, unsigned long long* n_out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_members) return;
  const int64_t end = (int64_t)member_end[t];
  const uint32_t ida = part_id[members[t]];
  for (int64_t b = t + 1; b < end; ++b) edge_insert(ida, part_id[members[b]], table, log_slots, out_u, out_v, n_out);
}

// union of edge codes ((u << 32) | v) gathered from several ranks
__global__ void k_edge_merge_codes(const uint64_t* codes, int64_t n, uint64_t* table, int log_slots, uint32_t* out_u, uint32_t* out_v,
                                   unsigned long long* n_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) edge_insert((uint32_t)(codes[i] >> 32), (uint32_t)codes[i], table, log_slots, out_u, out_v, n_out);
}

// union of edge lists gathered from several ranks
__global__ void k_edge_merge(const uint32_t* u, const uint32_t* v, int64_t n, uint64_t* table, int log_slots, uint32_t* out_u, uint32_t* out_v,
                             unsigned long long* n_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) edge_insert(u[i], v[i], table, log_slots, out_u, out_v, n_out);
}

// sharded path: the parts this rank flagged for an internal repeat (it owns the repeated k-mer) as a compact list
// list[0] = count (may exceed cap: the list then holds the first cap entries), list[1..] = part indices.  16 flag bytes per thread.
__global__ void k_repeat_list(const uint8_t* flags, int64_t n, uint32_t* list, uint32_t cap) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i0 >= n) return;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (i0 + 16 <= n) v = *reinterpret_cast<const uint4*>(flags + i0);                  // the flag buffer is 16-byte aligned and padded
  else for (int b = 0; b < 16 && i0 + b < n; ++b) reinterpret_cast<uint8_t*>(&v)[b] = flags[i0 + b];
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  if (((w[0] | w[1] | w[2] | w[3]) & 0x01010101u) == 0) return;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    for (int b = 0; b < 4; ++b)
      if ((w[q] >> (8 * b)) & kFlagRepeat) {
        const uint32_t pos = atomicAdd(list, 1u);
        if (pos < cap) list[1 + pos] = (uint32_t)(i0 + 4 * q + b);
      }
}
// ... and the lists of all ranks (world rows of 1 + cap words) applied to this rank's flags; *overflow is set when some rank's
// list was truncated (the caller then falls back to the dense all-reduce)
__global__ void k_repeat_apply(const uint32_t* lists, int world, uint32_t cap, uint8_t* flags, int64_t n, uint32_t* overflow) {
  for (int r = 0; r < world; ++r) {
    const uint32_t* row = lists + (size_t)r * (cap + 1);
    const uint32_t count = row[0];
    if (count > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) *overflow = 1u; continue; }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      const uint32_t p = row[1 + i];
      if ((int64_t)p < n) flags[p] |= kFlagRepeat;          // several ranks may flag one part: same value, benign
    }
  }
}

__global__ void k_count_flagged(const uint8_t* flags, int64_t n, unsigned long long* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool f = i < n && flags[i] != 0;
  uint32_t m = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

}  // namespace nrp
