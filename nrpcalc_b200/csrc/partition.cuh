// The bulk path: canonical k-mer extraction fused with a hash-radix partition into shared-memory sized
// buckets, and the in-bucket duplicate filter.
//
// Replaces the construction of repeat_dict in reference nrpcalc/base/hgraph.py:82-89.  The reference needs
// the table only to learn which canonical k-mers occur in more than one (window, part) record; every key
// that occurs once contributes nothing but the singleton tuple (id,).  So instead of sorting all M records
// we (1) scatter them by the top bits of a multiplicative hash into buckets of a few thousand records
// (two levels of <= 9 bits, ranks taken with shared-memory atomics -- order inside a bucket is irrelevant),
// (2) load every bucket into shared memory once and keep only the records whose key was seen twice.
// What survives is ordered inside the filter (or by the small fallback sort) and grouped.
//
// Two ways to place the buckets:
//   direct (default)  every bucket owns a FIXED region of mean + 12.5 % + 8 sigma records; no histogram pass.  The hash
//                     spreads random k-mers evenly, so a region overflows only on heavily duplicated input; the overflow
//                     is detected (nothing is written past a region) and the batch is repeated on the exact path.
//   exact             leaf sizes from a histogram pass over the parts (k_hist_from_ascii), tight offsets.
//
// HBM traffic per record on the direct path (key K_b bytes + 4-byte value): ASCII once, one write by split1, one read
// and one write by split2, one read by the filter  ->  L/W + 4*(K_b+4) bytes, against
// L/W + (2*ceil(2k/8)+4)*(K_b+4) + K_b for the LSD formulation in SURVEY.md 8d.
#pragma once
#include <type_traits>
#include "common.cuh"

// A/B switches of the split kernels (tools/build_variants.sh builds one library per combination; measured on config 2,
// profiles/README.md: LATE -0.018 ms split1 / -0.013 ms split2, DUMP + EARLY another -0.011 ms split2)
#ifndef NRP_V_DUMP
#define NRP_V_DUMP 1         // split2: slots past the end of a tile count in per-lane dump counters instead of branching
#endif
#ifndef NRP_V_EARLY
#define NRP_V_EARLY 1        // the cursor claims are issued before the scan of the bucket counts (they only need the counts)
#endif
#ifndef NRP_V_FPIPE
#define NRP_V_FPIPE 1        // filter: a leaf's candidates leave the kernel during the NEXT leaf's turn (the atomic that claims their
#endif                       // place in the output has a whole filter stage to come back)
#ifndef NRP_V_LATE
#define NRP_V_LATE 1         // ... and their results are consumed after the reorder stores: the round trip to L2 overlaps both
#endif

#ifndef NRP_V_KEYDIG
#define NRP_V_KEYDIG 1       // k_split_from_records on packed records: the flush takes the digit from the record, no digit array
#endif

namespace nrp {

constexpr int kSplitThreads = kTileThreads;     // 512
constexpr int kSplitItems = 8;
constexpr int kSplitTile = kSplitThreads * kSplitItems;   // 4096 records staged per CTA
constexpr int kMaxLevelBits = 9;
constexpr int kLevelBuckets = 1 << kMaxLevelBits;
constexpr int kLeafSortCap = 1024;              // candidates of one leaf that are ordered inside the filter kernel
static_assert(kLevelBuckets <= kSplitThreads, "one thread per bucket in split_scatter");

// PACKED: a record is ONE 64-bit word ((bijective hash of the k-mer without its top `drop` bits) << vbits) | value -- possible
// whenever 2k - drop + vbits <= 64 (vbits = bits of the part index plus bits of the window index, drop = bits of the first
// split level): one load / one store per record in every pass, 8 instead of 12 bytes, and no hashing after the extraction.
template <typename KeyT, bool PACKED = false>
struct alignas(16) SplitSmem {
  KeyT keys[kSplitTile];
  uint32_t vals[PACKED ? 4 : kSplitTile];
  uint16_t dig[kSplitTile];          // bucket of the record at this position of the reordered tile
  uint32_t cnt[kLevelBuckets + 32];  // + one dump counter per lane: register slots without a record count there (no branch)
  uint32_t base[kLevelBuckets];
  uint32_t delta[kLevelBuckets];     // global position of the bucket's run minus its position in the tile (kNoSpace: region full)
  uint32_t warp_sums[40];
  uint32_t total;
};

// Destination buffers of the owner-mode scatter when records are stored straight into the owners' receive buffers
// over NVLink (peer pointers opened with CUDA IPC); entry o = owner o's buffers (the own rank maps to local memory).
constexpr int kMaxPeers = 16;
struct PeerBuffers {
  void* keys[kMaxPeers];
  uint32_t* vals[kMaxPeers];
};

// Batches too large for two split levels are processed in several passes over disjoint slices of the hash space:
// a record belongs to pass `id` iff the hash bits selected by (shift, mask) equal id.  mask = 0 accepts everything.
struct PassSel {
  int shift; uint32_t mask; uint32_t id;
  __device__ __forceinline__ bool accept(uint64_t hash) const { return (((uint32_t)(hash >> shift)) & mask) == id; }
};

// Direct mode (no histogram pass): bucket i of a level owns the FIXED region [i * cap, (i + 1) * cap) of the output buffer
// and cursor[i] starts at i * cap.  A claim that would cross the region's end sets *overflow and writes nothing; the
// host then repeats the batch on the exact (histogram) path.  cap == 0: exact offsets, nothing to check.
// With peers (direct sharded path) cursor[d] indexes the OWNER's buffer and the region of digit d = (owner, bucket) is
// region_of_digit = (bucket * n_senders + my_rank): the limit is passed as explicit per-digit ends.
struct RegionLimit {
  uint32_t cap;          // records per region, 0 = unchecked
  uint32_t first;        // index of the region that cursor[0] belongs to
  uint32_t* overflow;
  const uint32_t* ends;  // optional: explicit end of digit d's region (overrides first/cap arithmetic)
};
constexpr uint32_t kNoSpace = 0xffffffffu;

// Steps shared by both split kernels once every thread holds its records, digits and in-bucket ranks:
// scan the bucket counts, claim global space for each non-empty bucket with one atomic per (tile, bucket),
// reorder the tile through shared memory and write each bucket's run contiguously.
// peer_shift >= 0 with peers: the buffers of digit d are those of owner d >> peer_shift (peer stores over NVLink).
template <typename KeyT, bool PACKED = false, bool KEYDIG = false>
__device__ __forceinline__ void split_reorder(SplitSmem<KeyT, PACKED>& s, const KeyT (&key)[kSplitItems], const uint32_t (&val)[kSplitItems],
                                              const uint32_t (&dr)[kSplitItems], int nb, uint32_t* cursor, RegionLimit lim) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int kWarps = kSplitThreads / 32;
  // exclusive scan of cnt[0..nb) by the first nb threads (nb <= kLevelBuckets <= blockDim)
  uint32_t c = (int)threadIdx.x < nb ? s.cnt[threadIdx.x] : 0u;
  const bool claims = c != 0;
  uint32_t g = 0;
  if (NRP_V_EARLY && claims) g = atomicAdd(&cursor[threadIdx.x], c);
  uint32_t inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s.warp_sums[wid] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t t = lane < kWarps ? s.warp_sums[lane] : 0u;
    uint32_t tin = t;
#pragma unroll
    for (int d = 1; d < kWarps; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, tin, d);
      if (lane >= d) tin += o;
    }
    if (lane < kWarps) s.warp_sums[lane] = tin - t;
    if (lane == kWarps - 1) s.total = tin;
  }
  __syncthreads();
  const uint32_t base = s.warp_sums[wid] + inc - c;
  // one atomic per (tile, bucket) claims the run's place; its result is needed by the flush only, so (NRP_V_LATE) it is
  // consumed after the reorder stores and the round trip to L2 overlaps them
  auto publish = [&]() {
    uint32_t dl = g - base + (uint32_t)kSplitTile;     // biased by the tile size: never wraps, never equals kNoSpace
    if (lim.cap) {
      const uint64_t end = lim.ends ? (uint64_t)lim.ends[threadIdx.x] : (uint64_t)(lim.first + threadIdx.x + 1u) * lim.cap;
      if ((uint64_t)g + c > end) { dl = kNoSpace; *lim.overflow = 1u; }
    }
    s.delta[threadIdx.x] = dl;
  };
  if ((int)threadIdx.x < nb) s.base[threadIdx.x] = base;
  if (claims) {
    if (!NRP_V_EARLY) g = atomicAdd(&cursor[threadIdx.x], c);
    if (!NRP_V_LATE) publish();
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < kSplitItems; ++e) {
    if (dr[e] != 0xffffffffu) {
      const uint32_t d = dr[e] >> 16;
      const uint32_t pos = s.base[d] + (dr[e] & 0xffffu);
      s.keys[pos] = key[e];
      if (!PACKED) s.vals[pos] = val[e];
      if (!KEYDIG) s.dig[pos] = (uint16_t)d;      // (KEYDIG: the flush reads the digit off the record itself)
    }
  }
  if (NRP_V_LATE && claims) publish();
  __syncthreads();
}

// second half: every bucket's run of the reordered tile goes to its place in global (or peer) memory
template <typename KeyT, bool PACKED = false>
__device__ __forceinline__ void split_flush(SplitSmem<KeyT, PACKED>& s, KeyT* out_keys, uint32_t* out_vals, const PeerBuffers* peers = nullptr,
                                            int peer_shift = 0) {
  const uint32_t total = s.total;
  for (uint32_t i = threadIdx.x; i < total; i += kSplitThreads) {
    const uint32_t d = s.dig[i];
    const uint32_t dl = s.delta[d];
    if (dl == kNoSpace) continue;
    const uint32_t dst = i + dl - (uint32_t)kSplitTile;
    if (peers) {                                   // digit d lives in its owner's receive buffer (peer store over NVLink)
      const uint32_t o = d >> peer_shift;
      reinterpret_cast<KeyT*>(peers->keys[o])[dst] = s.keys[i];
      if (!PACKED) peers->vals[o][dst] = s.vals[i];
    } else {
      out_keys[dst] = s.keys[i];
      if (!PACKED) out_vals[dst] = s.vals[i];
    }
  }
}

// The flush for packed records whose word still holds the digit of this level (every level but the first): the digit is the bit
// field (key >> key_shift) & key_mask of the record the flush loads anyway -- no digit array, one scattered 16-bit store and one
// load less per record in shared memory, whose pipe these kernels keep busy (~20 wavefronts per 32 records).
template <typename KeyT, bool PACKED>
__device__ __forceinline__ void split_flush_keydigit(SplitSmem<KeyT, PACKED>& s, KeyT* out_keys, int key_shift, uint32_t key_mask) {
  const uint32_t total = s.total;
  for (uint32_t i = threadIdx.x; i < total; i += kSplitThreads) {
    const KeyT kv = s.keys[i];
    const uint32_t dl = s.delta[(uint32_t)((uint64_t)kv >> key_shift) & key_mask];
    if (dl == kNoSpace) continue;
    out_keys[i + dl - (uint32_t)kSplitTile] = kv;
  }
}

// ---- histogram of leaf buckets straight from the part buffer (exact path) ---------------------------------
// hist has 2^bits entries (bits <= 15): leaf bucket = top `bits` bits of key_hash(canonical k-mer).
// Persistent CTAs accumulate in shared memory and flush once.
// HALVES groups of kTileThreads threads work on different tiles and share the histogram (HALVES = 2 keeps 32 warps
// on an SM when the 128 KB histogram of 2^15 buckets allows only one CTA).
template <int HALVES>
__global__ void __launch_bounds__(kTileThreads * HALVES) k_hist_from_ascii(Parts P, const uint8_t* flags, int k, int bits, int n_owners,
                                                                          int64_t n_tiles, PassSel sel, uint32_t* hist) {
  extern __shared__ uint32_t s_hist[];
  __shared__ TileStage stages[HALVES];
  const int half = threadIdx.x / kTileThreads, tid = threadIdx.x % kTileThreads;
  TileStage& stage = stages[half];
  const int nb = n_owners > 0 ? n_owners : (1 << bits);
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s_hist[i] = 0;
  const int64_t rounds = (n_tiles + (int64_t)gridDim.x * HALVES - 1) / ((int64_t)gridDim.x * HALVES);
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t t = (r * gridDim.x + blockIdx.x) * HALVES + half;
    const int64_t tile = P.tile_lo + t;
    __syncthreads();
    if (t < n_tiles) tile_load(P, tile, stage, tid);
    __syncthreads();
    if (t < n_tiles)
      tile_windows_blocked<true>(P, stage, tile, k, flags, tid, [&](int, uint64_t canon, uint32_t, uint32_t, bool) {
        const uint64_t h = key_hash(canon);
        if (!sel.accept(h)) return;
        uint32_t bin = n_owners > 0 ? owner_hash(canon) % (uint32_t)n_owners : (uint32_t)(h >> (64 - bits));
        atomicAdd(&s_hist[bin], 1u);
      });
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    uint32_t v = s_hist[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// same histogram over a segmented record list given by tile descriptors {first record, end, -, -}
template <typename KeyT>
__global__ void __launch_bounds__(512) k_hist_from_tiles(const KeyT* keys, const uint4* tile_desc, const uint32_t* n_tiles_ptr, int bits,
                                                        PassSel sel, uint32_t* hist) {
  extern __shared__ uint32_t s_hist[];
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint32_t n_tiles = *n_tiles_ptr;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint4 d4 = __ldg(tile_desc + tile);
    for (uint32_t i = d4.x + threadIdx.x; i < d4.y; i += blockDim.x) {
      const uint64_t h = key_hash((uint64_t)keys[i]);
      if (sel.accept(h)) atomicAdd(&s_hist[h >> (64 - bits)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    uint32_t v = s_hist[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// packed words of a segmented record list -> (key, value) arrays at the same indices (fallback to the exact path only)
// (tile descriptor .z = index of the level-1 bucket the tile's segment belongs to = the hash bits the words do not store)
template <typename KeyT>
__global__ void __launch_bounds__(512) k_unpack_tiles(const uint64_t* words, const uint4* tile_desc, const uint32_t* n_tiles_ptr, PackFmt pf,
                                                     KeyT* out_keys, uint32_t* out_vals) {
  const uint32_t n_tiles = *n_tiles_ptr;
  const uint32_t vmask = (uint32_t)((1ull << pf.vbits) - 1ull);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint4 d4 = __ldg(tile_desc + tile);
    for (uint32_t i = d4.x + threadIdx.x; i < d4.y; i += blockDim.x) {
      const uint64_t w = words[i];
      out_keys[i] = (KeyT)pf.bij.inv(pf.compose(d4.z, w >> pf.vbits));
      out_vals[i] = (uint32_t)w & vmask;
    }
  }
}

// cursors of the direct path: cursor[i] = i * cap (region starts)
__global__ void k_init_region_cursors(uint32_t* cursor, uint32_t n, uint32_t cap) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cursor[i] = i * cap;
}

// records held by n fixed regions: sum of min(cursor[i], (i+1)*cap) - i*cap  (one CTA)
__global__ void __launch_bounds__(1024) k_region_total(const uint32_t* cursor, uint32_t n, uint32_t cap, unsigned long long* out) {
  __shared__ unsigned long long warp_sums[32];
  unsigned long long sum = 0;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t lo = i * cap;
    uint32_t e = cursor[i];
    if (e > lo + cap) e = lo + cap;
    sum += e - lo;
  }
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_sums[w];
    *out = t;
  }
}

// ---- level 1: extraction fused with the first split ------------------------------------------------------
// mode 0  digit = top r1 bits of the hash
// mode 1  digit = owner_hash % n_owners (sharded path, exact exchange); with use_peers the runs go to the owners' buffers
// mode 2  digit = (owner_of, top r1 bits) (direct sharded path): stored into the owner's region of this (bucket, sender)
// cursor[d] must hold the start offset of bucket d and is advanced atomically.
// Record value = part index, or (part << j_bits) | window index when both fit 32 bits (j_bits > 0): sorting by value
// is then sorting by (part, window), and a group's first record carries the group's rank for free.
template <typename KeyT, int MODE, bool PASSES, bool PACKED = false>
__global__ void __launch_bounds__(kTileThreads, 2) k_split_from_ascii(Parts P, const uint8_t* flags, int k, int r1, int n_owners, int64_t n_tiles,
                                                                  uint32_t part_base, int j_bits, PackFmt pf, PassSel sel, uint32_t* cursor, KeyT* out_keys,
                                                                  uint32_t* out_vals, RegionLimit lim, PeerBuffers peers, int use_peers) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SplitSmem<KeyT, PACKED>& s = *reinterpret_cast<SplitSmem<KeyT, PACKED>*>(smem_raw);
  TileStage& stage = *reinterpret_cast<TileStage*>(smem_raw + sizeof(SplitSmem<KeyT, PACKED>));
  const int nb = MODE == 0 ? (1 << r1) : MODE == 1 ? n_owners : (n_owners << r1);
  static_assert(kLevelBuckets == kTileThreads, "one counter per thread");
  const uint32_t j_mask = j_bits ? 0xffffffffu : 0u;
  const int top_shift = PACKED ? pf.n_eff() - r1 : 64 - r1;
  const uint64_t low_mask = PACKED ? ((1ull << (pf.n_eff() - pf.drop)) - 1ull) : 0ull;    // hash bits a packed record stores
  const bool hash_owner = PACKED && MODE == 2 && pf.owner_bits > 0;                       // owner = top bits of the hash itself
  // persistent CTAs: the global loads of the next tile are issued before the current tile is written out
  TileRegs pre;
  if ((int64_t)blockIdx.x < n_tiles) tile_fetch(P, P.tile_lo + blockIdx.x, threadIdx.x, flags, pre);
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t tile = P.tile_lo + t;
    s.cnt[threadIdx.x] = 0;
    tile_commit(P, tile, stage, threadIdx.x, flags, pre);
    __syncthreads();
    KeyT key[kSplitItems];
    uint32_t val[kSplitItems], dr[kSplitItems];
#pragma unroll
    for (int e = 0; e < kSplitItems; ++e) { dr[e] = 0xffffffffu; key[e] = 0; val[e] = 0; }
    tile_windows_blocked<true>(P, stage, tile, k, flags, threadIdx.x, [&](int e, uint64_t canon, uint32_t part, uint32_t j, bool) {
      uint32_t d;
      uint64_t h = 0;
      if (MODE == 1) {
        d = owner_hash(canon) % (uint32_t)n_owners;
        if (PASSES && !sel.accept(key_hash(canon))) return;
      } else {
        h = PACKED ? pf.bij.fwd(canon) : key_hash(canon);
        if (PASSES && !sel.accept(h)) return;
        if (MODE == 0) d = (uint32_t)(h >> top_shift);
        else if (hash_owner) d = (uint32_t)(h >> top_shift);            // (owner, bucket) = the top owner_bits + r1 bits
        else d = (owner_of(canon, (uint32_t)n_owners) << r1) | (r1 ? (uint32_t)(h >> top_shift) : 0u);
      }
      const uint32_t rank = atomicAdd(&s.cnt[d], 1u);
#pragma unroll
      for (int q = 0; q < kSplitItems; ++q)
        if (q == e) {
          const uint32_t v = ((part_base + part) << j_bits) | (j & j_mask);
          if (PACKED) key[q] = (KeyT)(((h & low_mask) << pf.vbits) | v);
          else { key[q] = (KeyT)canon; val[q] = v; }
          dr[q] = (d << 16) | rank;
        }
    });
    __syncthreads();
    split_reorder<KeyT, PACKED>(s, key, val, dr, nb, cursor, lim);
    if (t + gridDim.x < n_tiles) tile_fetch(P, tile + gridDim.x, threadIdx.x, flags, pre);
    if (MODE == 0) split_flush<KeyT, PACKED>(s, out_keys, out_vals);
    else if (MODE == 1) split_flush<KeyT, PACKED>(s, out_keys, out_vals, use_peers ? &peers : nullptr, 0);
    else split_flush<KeyT, PACKED>(s, out_keys, out_vals, &peers, r1);
    __syncthreads();
  }
}

// ---- level 2 (or level 1 when the records are already in HBM): segmented split of records ---------------
// The input is a list of segments (level-1 buckets, or the per-sender parts of them on the sharded path); every
// segment is cut into tiles of kSplitTile records and each tile has a descriptor {first record, end, index of the
// first child cursor}.  Persistent CTAs walk the tile list; the next tile's descriptor is fetched one tile ahead.
// Child bucket = hash bits [64 - r_done - r2, 64 - r_done).
template <typename KeyT, bool PASSES, bool PACKED = false>
__global__ void __launch_bounds__(kSplitThreads, 2) k_split_from_records(const KeyT* __restrict__ in_keys, const uint32_t* __restrict__ in_vals,
                                                                     const uint4* __restrict__ tile_desc, const uint32_t* n_tiles_ptr,
                                                                     int r_done, int r2, PackFmt pf, PassSel sel, uint32_t* cursor, KeyT* out_keys,
                                                                     uint32_t* out_vals, RegionLimit lim) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SplitSmem<KeyT, PACKED>& s = *reinterpret_cast<SplitSmem<KeyT, PACKED>*>(smem_raw);
  const uint32_t n_tiles = *n_tiles_ptr;
  const uint32_t mask = (1u << r2) - 1u;
  const int packed_shift = PACKED ? pf.vbits + pf.n_eff() - r_done - r2 : 0;
  KeyT key[kSplitItems];
  uint32_t val[kSplitItems], dr[kSplitItems];
  // the loads of a tile are issued while the previous tile is still being written out (its registers are free by then)
  auto load_tile = [&](const uint4& t4) {
#pragma unroll
    for (int e = 0; e < kSplitItems; ++e) {                  // all loads first: 16 independent requests per thread
      const uint32_t idx = t4.x + threadIdx.x + e * kSplitThreads;
      const bool ok = idx < t4.y;
      key[e] = ok ? __ldg(in_keys + idx) : KeyT(0);
      if (!PACKED) val[e] = ok ? __ldg(in_vals + idx) : 0u;
    }
  };
  uint4 d4 = blockIdx.x < n_tiles ? __ldg(tile_desc + blockIdx.x) : make_uint4(0, 0, 0, 0);
  uint4 next = blockIdx.x + gridDim.x < n_tiles ? __ldg(tile_desc + blockIdx.x + gridDim.x) : make_uint4(0, 0, 0, 0);
  if (blockIdx.x < n_tiles) load_tile(d4);
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t start = d4.x, end = d4.y, child0 = d4.z;
    s.cnt[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kSplitItems; ++e) {
      uint32_t idx = start + threadIdx.x + e * kSplitThreads;
      dr[e] = 0xffffffffu;
      if (PACKED && NRP_V_DUMP) {          // the digit is a bit field of the record: hash bits [n - r_done - r2, n - r_done)
        // branch-free: slots past the end of the tile count in the lane's dump counter
        const bool ok = idx < end;
        const uint32_t d = (uint32_t)((uint64_t)key[e] >> packed_shift) & mask;
        const uint32_t rank = atomicAdd(&s.cnt[ok ? d : (uint32_t)kLevelBuckets + (threadIdx.x & 31u)], 1u);
        dr[e] = ok ? ((d << 16) | rank) : 0xffffffffu;
      } else if (PACKED) {
        if (idx < end) {
          const uint32_t d = (uint32_t)((uint64_t)key[e] >> packed_shift) & mask;
          dr[e] = (d << 16) | atomicAdd(&s.cnt[d], 1u);
        }
      } else if (idx < end) {
        {
          const uint64_t h = key_hash((uint64_t)key[e]);
          if (!PASSES || sel.accept(h)) {
            uint32_t d = (uint32_t)(h >> (64 - r_done - r2)) & mask;
            dr[e] = (d << 16) | atomicAdd(&s.cnt[d], 1u);
          }
        }
      }
    }
    __syncthreads();
    RegionLimit tile_lim = lim;
    tile_lim.first = child0;
    split_reorder<KeyT, PACKED, PACKED && NRP_V_KEYDIG>(s, key, val, dr, 1 << r2, cursor + child0, tile_lim);
    d4 = next;
    if (tile + gridDim.x < n_tiles) {
      load_tile(d4);
      if (tile + 2 * gridDim.x < n_tiles) next = __ldg(tile_desc + tile + 2 * gridDim.x);
    }
    if (PACKED && NRP_V_KEYDIG) split_flush_keydigit<KeyT, PACKED>(s, out_keys, packed_shift, mask);
    else split_flush<KeyT, PACKED>(s, out_keys, out_vals);
    __syncthreads();
  }
}

// Extents of the segments a segmented split reads.
//   exact   segment g = records [leaf_off[g << shift], leaf_off[(g + 1) << shift])
//   direct  segment g = records [g * cap, min(ends[g], (g + 1) * cap))   (ends = final cursors of the previous level, or
//           the end positions the senders delivered on the sharded path)
struct SegExtents {
  const uint32_t* leaf_off; int shift;
  const uint32_t* ends; uint32_t cap;
  __device__ __forceinline__ void get(uint32_t g, uint32_t& a, uint32_t& b) const {
    if (cap) {
      a = g * cap;
      b = ends[g];
      if (b > a + cap) b = a + cap;
      if (b < a) b = a;
    } else {
      a = leaf_off[(size_t)g << shift];
      b = leaf_off[(size_t)(g + 1) << shift];
    }
  }
};

// tiles per segment -> tile_prefix (n_seg + 1 entries) and the total number of records in *n_records.  One CTA.
__global__ void __launch_bounds__(1024) k_segment_scan(SegExtents ext, uint32_t n_seg, uint32_t* tile_prefix, unsigned long long* n_records) {
  __shared__ uint32_t warp_sums[33];
  __shared__ unsigned long long rec_sums[32];
  const uint32_t per = (n_seg + blockDim.x - 1) / blockDim.x;
  const uint32_t g0 = threadIdx.x * per, g1 = min(g0 + per, n_seg);
  uint32_t tiles = 0;
  unsigned long long recs = 0;
  for (uint32_t g = g0; g < g1; ++g) {
    uint32_t a, b;
    ext.get(g, a, b);
    tiles += (b - a + kSplitTile - 1) / kSplitTile;
    recs += b - a;
  }
  // block exclusive scan of `tiles`
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = tiles;
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  for (int d = 16; d > 0; d >>= 1) recs += __shfl_down_sync(0xffffffffu, recs, d);
  if (lane == 31) warp_sums[wid] = inc;
  if (lane == 0) rec_sums[wid] = recs;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = warp_sums[lane], winc = w;
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    warp_sums[lane] = winc - w;
    if (lane == 31) warp_sums[32] = winc;
  }
  __syncthreads();
  uint32_t run = warp_sums[wid] + inc - tiles;
  for (uint32_t g = g0; g < g1; ++g) {
    uint32_t a, b;
    ext.get(g, a, b);
    tile_prefix[g] = run;
    run += (b - a + kSplitTile - 1) / kSplitTile;
  }
  if (threadIdx.x == 0) {
    tile_prefix[n_seg] = warp_sums[32];
    if (n_records) {
      unsigned long long t = 0;
      for (int w = 0; w < 32; ++w) t += rec_sums[w];
      *n_records = t;
    }
  }
}

// tile descriptors: one warp per segment.  The children of segment g are the cursors (g / segs_per_bucket) << r_child ...
__global__ void __launch_bounds__(256) k_segment_fill(SegExtents ext, uint32_t n_seg, const uint32_t* tile_prefix, uint32_t segs_per_bucket,
                                                     int r_child, uint4* tile_desc) {
  const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= n_seg) return;
  uint32_t a, b;
  ext.get(g, a, b);
  const uint32_t t0 = tile_prefix[g], nt = tile_prefix[g + 1] - t0;
  const uint32_t child0 = (g / segs_per_bucket) << r_child;
  for (uint32_t t = threadIdx.x & 31; t < nt; t += 32) {
    const uint32_t start = a + t * kSplitTile;
    tile_desc[t0 + t] = make_uint4(start, min(start + (uint32_t)kSplitTile, b), child0, g);
  }
}

// Extents of the leaf buckets the filter reads (same two forms as SegExtents with shift = 0).
struct LeafExtents {
  const uint32_t* off;      // exact: n_leaf + 1 offsets
  const uint32_t* ends;     // direct: final cursor of every leaf region
  uint32_t cap;             // direct: records per region (0 = exact)
  __device__ __forceinline__ void get(uint32_t lf, uint32_t n_leaf, uint32_t& lo, uint32_t& n) const {
    lo = 0; n = 0;
    if (lf >= n_leaf) return;
    if (cap) {
      lo = lf * cap;
      uint32_t e = __ldg(ends + lf);
      if (e > lo + cap) e = lo + cap;
      n = e > lo ? e - lo : 0u;
    } else {
      lo = __ldg(off + lf);
      n = __ldg(off + lf + 1) - lo;
    }
  }
};

// ---- in-bucket duplicate filter -------------------------------------------------------------------------
// One CTA per leaf bucket (work-stealing).  Keys go into an open-addressing table in shared memory; a key met
// again marks its slot.  Records whose slot is marked -- i.e. whose canonical k-mer occurs at least twice --
// are appended to the candidate list.  Buckets larger than `cap` records are appended whole (the small sort
// then sees a superset, which is still correct).  EMPTY = all ones is never a canonical k-mer: the reverse
// complement of T..T is A..A = 0.
__device__ __forceinline__ void cp_async_u32(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_u64(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_key(uint32_t* d, const uint32_t* s) { cp_async_u32(d, s); }
__device__ __forceinline__ void cp_async_key(uint64_t* d, const uint64_t* s) { cp_async_u64(d, s); }

// STAGED: the records of the NEXT leaf are copied global -> shared with cp.async (LDGSTS) while the current leaf is
// processed out of registers, which takes the DRAM latency of a leaf off the critical path without costing
// registers (two 1024-thread CTAs per SM need <= 32 registers per thread).  CAP = records a leaf may hold to be
// filtered in shared memory (table of 2*T*E slots, staging of CAP records).
template <typename KeyT, int T, int E, int CAP, bool STAGED>
__global__ void __launch_bounds__(T, (E <= 4) ? 2 : 1) k_filter(const KeyT* __restrict__ keys, const uint32_t* __restrict__ vals, LeafExtents leaves,
                                              uint32_t n_leaf, int leaf_bits, uint32_t cap, KeyT* cand_keys, uint32_t* cand_vals,
                                              unsigned long long* n_cand, uint32_t* n_oversized, uint32_t* unsorted) {
  // candidates of one leaf are ordered by (key, part) in shared memory before they are written as one contiguous run:
  // equal keys end up adjacent and in part order WITHOUT a global sort (keys of different leaves differ).
  __shared__ KeyT c_keys[kLeafSortCap];
  __shared__ uint32_t c_vals[kLeafSortCap];
  __shared__ uint32_t c_rank[kLeafSortCap];
  __shared__ uint32_t s_ccount;
  __shared__ unsigned long long s_cbase;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KeyT* table = reinterpret_cast<KeyT*>(smem_raw);
  unsigned char* marked = smem_raw + sizeof(KeyT) * (size_t)(2 * T * E);
  KeyT* st_keys = reinterpret_cast<KeyT*>(marked + 2 * T * E);
  uint32_t* st_vals = reinterpret_cast<uint32_t*>(st_keys + (STAGED ? CAP : 0));
  const KeyT kEmpty = ~KeyT(0);
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kLeafSortCap; i += T) c_rank[i] = 0;
  uint32_t next_lo = 0, next_n = 0;
  auto stage_next = [&](uint32_t lf) {            // start the asynchronous copy of leaf lf into the staging area
    leaves.get(lf, n_leaf, next_lo, next_n);
    if (STAGED && next_n >= 2 && next_n <= cap) {
      for (uint32_t i = threadIdx.x; i < next_n; i += T) {
        cp_async_key(st_keys + i, keys + next_lo + i);
        cp_async_u32(st_vals + i, vals + next_lo + i);
      }
    }
    if (STAGED) asm volatile("cp.async.commit_group;" ::);
  };
  stage_next(blockIdx.x);
  for (uint32_t leaf = blockIdx.x; leaf < n_leaf; leaf += gridDim.x) {
    const uint32_t lo = next_lo, n = next_n;
    const bool fits = n >= 2 && n <= cap;
    KeyT key[E];
    uint32_t val[E], slot[E];
    if (STAGED) {
      asm volatile("cp.async.wait_group 0;" ::);
      __syncthreads();                              // staging holds this leaf; previous leaf's table readers are done
#pragma unroll
      for (int e = 0; e < E; ++e) {
        uint32_t i = threadIdx.x + e * T;
        bool ok = fits && i < n;
        key[e] = ok ? st_keys[i] : kEmpty;
        val[e] = ok ? st_vals[i] : 0u;
      }
      __syncthreads();                              // everyone copied out: the staging area is free again
      stage_next(leaf + gridDim.x);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {                 // issue the loads first; they are in flight while the table is cleared
        uint32_t i = threadIdx.x + e * T;
        bool ok = fits && i < n;
        key[e] = ok ? __ldg(keys + lo + i) : kEmpty;
        val[e] = ok ? __ldg(vals + lo + i) : 0u;
      }
      stage_next(leaf + gridDim.x);
      __syncthreads();                              // previous leaf's readers are done with the table
    }
    if (n < 2) continue;
    if (n > cap) {
      if (threadIdx.x == 0) { atomicAdd(n_oversized, 1u); *unsorted = 1u; }
      for (uint32_t i = threadIdx.x; i < n + 31u - ((n - 1u) & 31u) ; i += T) {   // whole warps iterate together
        bool ok = i < n;
        uint32_t m = __ballot_sync(0xffffffffu, ok);
        unsigned long long basepos = 0;
        if (lane == 0) basepos = atomicAdd(n_cand, (unsigned long long)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, 0);
        if (ok) {
          unsigned long long dst = basepos + __popc(m & lanemask_lt());
          cand_keys[dst] = keys[lo + i];
          cand_vals[dst] = vals[lo + i];
        }
      }
      continue;
    }
    uint32_t slots = 64;
    while (slots < 2u * n) slots <<= 1;
    const uint32_t smask = slots - 1u;
    const int log_slots = 31 - __clz(slots);
    for (uint32_t i = threadIdx.x; i < slots; i += T) { table[i] = kEmpty; marked[i] = 0; }
    if (threadIdx.x == 0) s_ccount = 0;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < E; ++e) {
      slot[e] = 0xffffffffu;
      if (key[e] != kEmpty) {
        uint32_t h = (uint32_t)(key_hash((uint64_t)key[e]) >> (64 - leaf_bits - log_slots)) & smask;
        for (;;) {
          KeyT old = atomic_cas(&table[h], kEmpty, key[e]);
          if (old == kEmpty) break;
          if (old == key[e]) { marked[h] = 1; break; }
          h = (h + 1u) & smask;
        }
        slot[e] = h;
      }
    }
    __syncthreads();
    bool dup[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      dup[e] = slot[e] != 0xffffffffu && marked[slot[e]];
      if (dup[e]) {
        uint32_t pos = atomicAdd(&s_ccount, 1u);
        if (pos < (uint32_t)kLeafSortCap) { c_keys[pos] = key[e]; c_vals[pos] = val[e]; }
      }
    }
    __syncthreads();
    const uint32_t c = s_ccount;
    if (c == 0) continue;
    if (c <= (uint32_t)kLeafSortCap) {
      // rank by counting: position of candidate i in (key, part, arrival) order.  T / c threads share one candidate
      // (each scans a strided part of the list), so the loop stays short when many records of a leaf are shared.
      const uint32_t share = T / c;                           // >= 1 because c <= kLeafSortCap <= T
      const uint32_t cand = threadIdx.x % c, seg = threadIdx.x / c;
      if (seg < share) {
        const KeyT mk = c_keys[cand];
        const uint32_t mv = c_vals[cand];
        uint32_t partial = 0;
        for (uint32_t j = seg; j < c; j += share) {
          KeyT ok = c_keys[j]; uint32_t ov = c_vals[j];
          partial += (ok < mk || (ok == mk && (ov < mv || (ov == mv && j < cand)))) ? 1u : 0u;
        }
        if (partial) atomicAdd(&c_rank[cand], partial);
      }
      if (threadIdx.x == 0) s_cbase = atomicAdd(n_cand, (unsigned long long)c);
      __syncthreads();
      if (threadIdx.x < c) {
        const uint32_t rank = c_rank[threadIdx.x];
        cand_keys[s_cbase + rank] = c_keys[threadIdx.x];
        cand_vals[s_cbase + rank] = c_vals[threadIdx.x];
        c_rank[threadIdx.x] = 0;                              // ready for the next leaf (ordered by the loop-top barrier)
      }
    } else {
      // too many candidates for the in-leaf ordering: emit unordered and ask for the global sort
      if (threadIdx.x == 0) *unsorted = 1u;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        uint32_t m = __ballot_sync(0xffffffffu, dup[e]);
        if (m) {
          unsigned long long basepos = 0;
          if (lane == 0) basepos = atomicAdd(n_cand, (unsigned long long)__popc(m));
          basepos = __shfl_sync(0xffffffffu, basepos, 0);
          if (dup[e]) {
            unsigned long long dst = basepos + __popc(m & lanemask_lt());
            cand_keys[dst] = key[e];
            cand_vals[dst] = val[e];
          }
        }
      }
    }
  }
  if (STAGED) asm volatile("cp.async.wait_group 0;" ::);
}

// ---- in-bucket duplicate filter, two stages ---------------------------------------------------------------------
// ~99 % of the records of a leaf are singletons.  Stage 1 throws most of them out with two 65536-bit maps in shared
// memory: every record sets its bit in map A, a record that finds its bit already set also sets it in map B; afterwards
// a record is a SUSPECT iff its bit is set in B (every key that occurs twice makes both of its records suspects; a
// singleton is a suspect only if another key shares its bit, ~n/65536 of them).  Stage 2 runs the exact open-addressing
// test on the few suspects only and orders the confirmed ones by (key, part, window) before they are written as one run.
// Per record this is a hash, one or two shared-memory atomics and a bit test instead of a CAS loop in a table that has to
// be cleared for every leaf.  Leaves above `cap` records, or with more than kLeafSortCap suspects, are passed on
// unfiltered / unordered and the host runs the global sort (always correct, only slower).
// --- TMA (cp.async.bulk) + mbarrier helpers: one thread streams a leaf bucket global -> shared, the CTA waits on the barrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();          // never hang the device: a lost copy is a bug, fail loudly
  }
}

// capacity of a leaf = T * E records; WORDS = 32-bit words per bit map (16 bits per record or more).
// Shared memory: exact table + suspect list + two bit maps + key staging (one leaf) + value staging (two leaves, ping-pong;
// none for packed records).  SORTCAP = suspects of one leaf that the exact stage and the in-leaf ordering can take (more:
// the global sort runs).  REC = bytes of a record word (8 for packed records, else the key size).
template <int REC, bool PACKED, int CAP, int WORDS, int SORTCAP> constexpr size_t filter_bloom_smem() {
  return (size_t)REC * (2 * SORTCAP + SORTCAP) + 4 * (2 * WORDS + 2 * SORTCAP) + 2 * SORTCAP +
         (size_t)(REC + 4) * SORTCAP +                    // the ordered candidates of the previous leaf (NRP_V_FPIPE)
         (size_t)REC * (CAP + 8) + (PACKED ? 0 : 2 * 4 * (CAP + 8)) + 128;
}

template <typename KeyT, int T, int E, int WORDS, int MINB, int SORTCAP, bool PACKED>
__global__ void __launch_bounds__(T, MINB) k_filter_bloom(const void* __restrict__ keys_in, const uint32_t* __restrict__ vals, LeafExtents leaves,
                                                      uint32_t n_leaf, int leaf_bits, PackFmt pf, uint32_t cap, KeyT* cand_keys, uint32_t* cand_vals,
                                                      unsigned long long* n_cand, uint32_t* n_oversized, uint32_t* unsorted) {
  // RecT: what a leaf holds -- the key (values in a second array) or the packed word (key << vbits) | value
  using RecT = typename std::conditional<PACKED, uint64_t, KeyT>::type;
  const RecT* __restrict__ keys = reinterpret_cast<const RecT*>(keys_in);
  constexpr int kBloomWords = WORDS, kBloomCap = T * E;
  constexpr int kSuspectSlots = 2 * SORTCAP, kLeafSortCap = SORTCAP;
  static_assert(SORTCAP <= T, "one thread per suspect");
  constexpr uint32_t kBitMask = (uint32_t)WORDS * 32u - 1u;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // 16-byte aligned blocks first (they are cleared with 128-bit stores)
  RecT* table = reinterpret_cast<RecT*>(smem_raw);                                  // kSuspectSlots (canonical k-mers)
  uint32_t* map_a = reinterpret_cast<uint32_t*>(table + kSuspectSlots);             // kBloomWords
  uint32_t* map_b = map_a + kBloomWords;                                             // kBloomWords
  unsigned char* marked = reinterpret_cast<unsigned char*>(map_b + kBloomWords);    // kSuspectSlots
  RecT* c_keys = reinterpret_cast<RecT*>(marked + kSuspectSlots);                   // kLeafSortCap
  RecT* p_keys = c_keys + kLeafSortCap;                                              // kLeafSortCap: the previous leaf's candidates, ordered
  uint32_t* c_vals = reinterpret_cast<uint32_t*>(p_keys + kLeafSortCap);            // kLeafSortCap
  uint32_t* c_rank = c_vals + kLeafSortCap;                                          // kLeafSortCap
  uint32_t* p_vals = c_rank + kLeafSortCap;                                          // kLeafSortCap
  // staging, filled by TMA bulk copies (16-byte aligned source windows around the leaf): the records of the NEXT leaf (they are
  // copied to registers at the top of a leaf's turn) and the values of this and the next leaf (only suspects read theirs)
  RecT* st_keys = reinterpret_cast<RecT*>((reinterpret_cast<uintptr_t>(p_vals + kLeafSortCap) + 15) & ~uintptr_t(15));
  uint32_t* st_vals0 = reinterpret_cast<uint32_t*>(st_keys + kBloomCap + 8);
  uint32_t* st_vals1 = st_vals0 + kBloomCap + 8;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_c1, s_c2;
  __shared__ unsigned long long s_cbase;
  const RecT kEmpty = ~RecT(0);
  const int vbits = pf.vbits;
  const uint32_t vmask = PACKED ? (uint32_t)((1ull << vbits) - 1ull) : 0u;
  // packed: what identifies the k-mer inside a leaf is the stored hash (record >> vbits); its bits below the leaf bits feed the
  // bit maps and the exact table through two cheap 32-bit mixes
  const uint64_t rem_mask = PACKED ? ((1ull << (pf.n_eff() - leaf_bits)) - 1ull) : 0ull;
  constexpr int kLogBits = WORDS == 1024 ? 15 : WORDS == 2048 ? 16 : 17, kLogSlots = SORTCAP == 256 ? 9 : SORTCAP == 512 ? 10 : 11;
  static_assert((WORDS == 1024 || WORDS == 2048 || WORDS == 4096) && (SORTCAP == 256 || SORTCAP == 512 || SORTCAP == 1024), "index widths");
  auto ident_of = [&](RecT r) -> uint64_t { return PACKED ? ((uint64_t)r >> vbits) : (uint64_t)r; };
  auto fold_rem = [&](RecT r) -> uint32_t { const uint64_t rem = ((uint64_t)r >> vbits) & rem_mask; return (uint32_t)rem ^ (uint32_t)(rem >> 32); };
  const int lane = threadIdx.x & 31;
  const int bit_shift = 44 - leaf_bits, slot_shift = 33 - leaf_bits;     // 20 hash bits for the maps, the next 11 for the table
  for (int i = threadIdx.x; i < kLeafSortCap; i += T) c_rank[i] = 0;
  constexpr uint32_t kKeysPer16 = 16 / sizeof(RecT);
  // thread 0 starts the bulk copies of a leaf: the source windows are widened to 16-byte boundaries
  auto stage = [&](uint32_t slo, uint32_t sn, uint32_t* st_vals) {
    if (sn < 2 || sn > cap || sn > (uint32_t)kBloomCap) return;
    const uint32_t k0 = slo & ~(kKeysPer16 - 1), k1 = (slo + sn + kKeysPer16 - 1) & ~(kKeysPer16 - 1);
    const uint32_t v0 = slo & ~3u, v1 = (slo + sn + 3u) & ~3u;
    const uint32_t kb = (k1 - k0) * (uint32_t)sizeof(RecT), vb = PACKED ? 0u : (v1 - v0) * 4u;
    mbar_expect_tx(&s_bar, kb + vb);
    tma_bulk_load(st_keys, keys + k0, kb, &s_bar);
    if (!PACKED) tma_bulk_load(st_vals, vals + v0, vb, &s_bar);
  };
  // one candidate record leaves the kernel as (key, value)
  // (packed: the hash bits the record does not store are the top `drop` bits of the leaf index; the k-mer is inv(hash))
  auto emit = [&](unsigned long long dst, RecT r, uint32_t v, uint32_t lf) {
    if (PACKED) {
      cand_keys[dst] = (KeyT)pf.bij.inv(pf.compose(lf >> (leaf_bits - pf.drop), (uint64_t)r >> vbits));
      cand_vals[dst] = (uint32_t)r & vmask;
    } else { cand_keys[dst] = (KeyT)r; cand_vals[dst] = v; }
  };
  // leaf extents are read two leaves ahead so that no global-load latency sits on the per-leaf critical path
  auto extent = [&](uint32_t lf, uint32_t& elo, uint32_t& en) { leaves.get(lf, n_leaf, elo, en); };
  // The exact table, both maps and the marks are contiguous and 16-byte aligned: cleared with 128-bit stores.  They are CLEAN at
  // the top of every leaf: each structure is cleared for the next leaf right after its last use in the current one, while the
  // threads that still work on the leaf carry on (one barrier less per leaf, no dedicated clearing phase).
  constexpr int kTable16 = (int)(sizeof(RecT) * kSuspectSlots / 16), kMaps16 = (int)(8 * kBloomWords / 16), kMarks16 = kSuspectSlots / 16;
  uint4* const z16 = reinterpret_cast<uint4*>(smem_raw);
  const uint4 ones16 = make_uint4(~0u, ~0u, ~0u, ~0u), zeros16 = make_uint4(0u, 0u, 0u, 0u);
  auto clear_maps = [&]() {
    for (int i = threadIdx.x; i < kMaps16; i += T) z16[kTable16 + i] = zeros16;
  };
  auto clear_table = [&]() {
    for (int i = threadIdx.x; i < kTable16; i += T) z16[i] = ones16;
    for (int i = threadIdx.x; i < kMarks16; i += T) z16[kTable16 + kMaps16 + i] = zeros16;
  };
  clear_maps();
  clear_table();
  // NRP_V_FPIPE: the ordered candidates of a leaf wait in p_keys / p_vals; thread 0 holds the (not yet consumed) result of the
  // atomic that claimed their place and publishes it one filter stage into the next leaf's turn
  uint32_t pend_n = 0, pend_leaf = 0;               // uniform across the CTA
  unsigned long long pend_base = 0;                 // thread 0 only
  auto pend_emit = [&]() {                          // after a barrier that follows `s_cbase = pend_base`
    if (threadIdx.x < pend_n) emit(s_cbase + threadIdx.x, p_keys[threadIdx.x], PACKED ? 0u : p_vals[threadIdx.x], pend_leaf);
    pend_n = 0;
  };
  auto pend_flush = [&]() {                         // on the paths that skip the filter stages (uniform)
    if (pend_n == 0) return;
    if (threadIdx.x == 0) s_cbase = pend_base;
    __syncthreads();
    pend_emit();
  };
  uint32_t lo, n, lo1, n1, lo2, n2;
  extent(blockIdx.x, lo, n);
  extent(blockIdx.x + gridDim.x, lo1, n1);
  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) stage(lo, n, st_vals0);
  uint32_t parity = 0, turn = 0;
  for (uint32_t leaf = blockIdx.x; leaf < n_leaf; leaf += gridDim.x, lo = lo1, n = n1, lo1 = lo2, n1 = n2, turn ^= 1u) {
    extent(leaf + 2 * gridDim.x, lo2, n2);
    const bool fits = n >= 2 && n <= cap && n <= (uint32_t)kBloomCap;
    const uint32_t* my_vals = (turn ? st_vals1 : st_vals0) + (lo & 3u);          // this leaf's values (valid for the whole turn)
    // records held by this thread: i = threadIdx.x + e * T < n
    const int ne = (fits && n > threadIdx.x) ? (int)((n - threadIdx.x + (uint32_t)T - 1u) / (uint32_t)T) : 0;
    RecT key[E];
    if (fits) {                                     // uniform: the copy of this leaf was issued one iteration ago
      mbar_wait(&s_bar, parity);
      parity ^= 1u;
      const RecT* src = st_keys + (lo & (kKeysPer16 - 1)) + threadIdx.x;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (e >= ne) break;
        key[e] = src[e * T];
      }
    }
    __syncthreads();                                // everyone copied out; the previous leaf is done with the shared arrays
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before the async-proxy overwrite
      stage(lo1, n1, turn ? st_vals0 : st_vals1);
      s_c1 = 0; s_c2 = 0;                           // every reader of the previous leaf's counts is past the barrier above
    }
    if (n < 2) { pend_flush(); continue; }
    if (!fits) {
      pend_flush();
      if (threadIdx.x == 0) { atomicAdd(n_oversized, 1u); *unsorted = 1u; }
      for (uint32_t i = threadIdx.x; i < n + 31u - ((n - 1u) & 31u) ; i += T) {   // whole warps iterate together
        bool ok = i < n;
        uint32_t m = __ballot_sync(0xffffffffu, ok);
        unsigned long long basepos = 0;
        if (lane == 0) basepos = atomicAdd(n_cand, (unsigned long long)__popc(m));
        basepos = __shfl_sync(0xffffffffu, basepos, 0);
        if (ok) emit(basepos + __popc(m & lanemask_lt()), keys[lo + i], PACKED ? 0u : vals[lo + i], leaf);
      }
      continue;
    }
    // stage 1a: set bits (the zeroing of s_c1 / s_c2 above is ordered before their first use by the barrier after this stage)
    uint32_t bits[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (e >= ne) break;
      bits[e] = PACKED ? (fold_rem(key[e]) * 0x9E3779B1u) >> (32 - kLogBits) : (uint32_t)(key_hash((uint64_t)key[e]) >> bit_shift) & kBitMask;
      const uint32_t w = bits[e] >> 5, m = 1u << (bits[e] & 31u);
      if (atomicOr(&map_a[w], m) & m) atomicOr(&map_b[w], m);
    }
    if (pend_n != 0 && threadIdx.x == 0) s_cbase = pend_base;      // the claim issued at the end of the previous leaf has landed by now
    __syncthreads();
    if (pend_n != 0) pend_emit();
    // stage 1b: collect the suspects -- one shared-memory atomic per warp claims the slots of its suspects
    uint32_t smask = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (e >= ne) break;
      smask |= ((map_b[bits[e] >> 5] >> (bits[e] & 31u)) & 1u) << e;
    }
    const uint32_t mine = __popc(smask);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    uint32_t wbase = 0;
    if (lane == 31 && incl) wbase = atomicAdd(&s_c1, incl);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    if (smask) {
      uint32_t pos = wbase + incl - mine;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((smask >> e) & 1u) {
          if (pos < (uint32_t)kLeafSortCap) { c_keys[pos] = key[e]; if (!PACKED) c_vals[pos] = my_vals[threadIdx.x + e * T]; }
          ++pos;
        }
      }
    }
    __syncthreads();
    const uint32_t c1 = s_c1;
    clear_maps();                                   // last use of the maps was the suspect test: clean for the next leaf
    if (c1 == 0) continue;
    if (c1 > (uint32_t)kLeafSortCap) {
      // heavy sharing: hand all suspects (a superset of the shared records) to the global sort
      if (threadIdx.x == 0) *unsorted = 1u;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const bool sus = (smask >> e) & 1u;
        uint32_t m = __ballot_sync(0xffffffffu, sus);
        if (m) {
          unsigned long long basepos = 0;
          if (lane == 0) basepos = atomicAdd(n_cand, (unsigned long long)__popc(m));
          basepos = __shfl_sync(0xffffffffu, basepos, 0);
          if (sus) emit(basepos + __popc(m & lanemask_lt()), key[e], PACKED ? 0u : my_vals[threadIdx.x + e * T], leaf);
        }
      }
      continue;
    }
    // stage 2: exact test on the suspects (thread t owns suspect t)
    uint32_t slot = 0xffffffffu;
    RecT mk = kEmpty;
    uint32_t mv = 0;
    if (threadIdx.x < c1) {
      mk = c_keys[threadIdx.x];
      if (!PACKED) mv = c_vals[threadIdx.x];
      const RecT canon = (RecT)ident_of(mk);
      uint32_t h = PACKED ? (fold_rem(mk) * 0x85EBCA6Bu) >> (32 - kLogSlots) : (uint32_t)(key_hash((uint64_t)canon) >> slot_shift) & (kSuspectSlots - 1);
      for (;;) {
        RecT old = atomic_cas(&table[h], kEmpty, canon);
        if (old == kEmpty) break;
        if (old == canon) { marked[h] = 1; break; }
        h = (h + 1u) & (kSuspectSlots - 1);
      }
      slot = h;
    }
    __syncthreads();
    // confirmed records are compacted to the front of the suspect list (every owner holds its record in registers)
    if (threadIdx.x < c1 && marked[slot] != 0) {
      const uint32_t pos = atomicAdd(&s_c2, 1u);
      c_keys[pos] = mk;
      if (!PACKED) c_vals[pos] = mv;
    }
    __syncthreads();
    const uint32_t c2 = s_c2;
    clear_table();                                  // table and marks are done: clean for the next leaf
    if (c2 == 0) continue;
    {  // order them by (key, part, window): rank by counting, a power-of-two group of threads per record (2^lg * c2 <= T)
      // largest power of two 2^lg with 2^lg * c2 <= T (no integer division: T is a power of two, c2 <= T)
      static_assert((T & (T - 1)) == 0, "T must be a power of two");
      const int lg = max(0, (31 - __clz((uint32_t)T)) - (32 - __clz(c2 - 1u)));       // log2(T) - ceil(log2(c2))
      const uint32_t cand = threadIdx.x >> lg, seg = threadIdx.x & ((1u << lg) - 1u);
      if (cand < c2) {
        const RecT ck = c_keys[cand];
        const uint32_t cv = PACKED ? 0u : c_vals[cand];
        uint32_t partial = 0;
        for (uint32_t j = seg; j < c2; j += (1u << lg)) {
          const RecT ok = c_keys[j];
          if constexpr (PACKED) {                     // the packed word orders by (key, value) as it is
            partial += (ok < ck || (ok == ck && j < cand)) ? 1u : 0u;
          } else {
            const uint32_t ov = c_vals[j];
            partial += (ok < ck || (ok == ck && (ov < cv || (ov == cv && j < cand)))) ? 1u : 0u;
          }
        }
        if (partial) atomicAdd(&c_rank[cand], partial);
      }
      if (threadIdx.x == 0) {
        if (NRP_V_FPIPE) pend_base = atomicAdd(n_cand, (unsigned long long)c2);
        else s_cbase = atomicAdd(n_cand, (unsigned long long)c2);
      }
    }
    __syncthreads();
    if (threadIdx.x < c2) {
      const uint32_t rank = c_rank[threadIdx.x];
      if (NRP_V_FPIPE) { p_keys[rank] = c_keys[threadIdx.x]; if (!PACKED) p_vals[rank] = c_vals[threadIdx.x]; }
      else emit(s_cbase + rank, c_keys[threadIdx.x], PACKED ? 0u : c_vals[threadIdx.x], leaf);
      c_rank[threadIdx.x] = 0;
    }
    if (NRP_V_FPIPE) { pend_n = c2; pend_leaf = leaf; }
  }
  pend_flush();
}

}  // namespace nrp
