// Shared device helpers: 2-bit encoding, canonical k-mers, hashing, byte-tile staging of the part buffer.
//
// Replaces reference nrpcalc/base/utils.py:10,37-50 (stream_kmers, get_comp, get_revcomp, stream_min_kmers).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nrp {

constexpr uint8_t kFlagRepeat = 1, kFlagPalindrome = 2, kFlagBackground = 4;

// number of kernels launched by this thread (reported through NRP_C_LAUNCHES)
inline int64_t& launch_counter() { static thread_local int64_t n = 0; return n; }
// ... and by every library call of this thread since the library was loaded (never reset; nrp_launches_total)
inline int64_t& launch_total() { static thread_local int64_t n = 0; return n; }
#define NRP_LAUNCH(kernel_call) do { ++::nrp::launch_counter(); ++::nrp::launch_total(); kernel_call; } while (0)

// ---------------------------------------------------------------------------------------------------
// hashing: one multiplicative (Fibonacci) hash per key; consecutive bit fields of the product, from the top,
// select the level-1 bucket, the level-2 bucket and the slot of the in-bucket table.
__host__ __device__ __forceinline__ uint64_t key_hash(uint64_t key) {
  key ^= key >> 32;                         // fold the high word into the low one (one LOP3 on the device)
  return key * 0x9E3779B97F4A7C15ull;
}
// independent hash for choosing the owner GPU of a key (sharded path)
__host__ __device__ __forceinline__ uint32_t owner_hash(uint64_t key) {
  key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 33; key *= 0xc4ceb9fe1a85ec53ull; key ^= key >> 33;
  return (uint32_t)(key >> 32);
}

// Bijective hash on the n = 2k bits of a canonical k-mer (multiply, fold, multiply -- every step invertible mod 2^n; the
// multipliers are 32-bit odd constants, so that a 64-bit product costs two IMADs, see extract.cuh).  Packed
// records carry this hash INSTEAD of the k-mer: every later bucket digit and filter index is a bit field of the record (no
// re-hashing after the extraction), the bits implied by a record's level-1 region need not be stored, and the filter
// recovers the k-mer of the few records it emits with inv().
struct BijHash {
  uint64_t c1inv, c2inv, mask;
  int n, s;
  static constexpr uint64_t kC1 = 0x9E3779B1ull, kC2 = 0x85EBCA6Bull;
  __host__ __device__ __forceinline__ uint64_t fwd(uint64_t x) const {
    uint64_t h = (x * kC1) & mask;
    h ^= h >> s;                                   // s >= n/2: an involution
    return (h * kC2) & mask;
  }
  __host__ __device__ __forceinline__ uint64_t inv(uint64_t h) const {
    h = (h * c2inv) & mask;
    h ^= h >> s;
    return (h * c1inv) & mask;
  }
};
inline uint64_t odd_inverse64(uint64_t x) {       // Newton iteration: the inverse of an odd number mod 2^64
  uint64_t inv = x;
  for (int i = 0; i < 6; ++i) inv *= 2 - x * inv;
  return inv;
}
inline BijHash make_bijhash(int k) {
  BijHash b;
  b.n = 2 * k;
  b.mask = b.n >= 64 ? ~0ull : ((1ull << b.n) - 1ull);
  b.s = (b.n + 1) / 2;
  b.c1inv = odd_inverse64(BijHash::kC1);
  b.c2inv = odd_inverse64(BijHash::kC2);
  return b;
}
// layout of a packed record: ((hash without its top owner_bits + drop bits) << vbits) | value.  The stored hash bits sit at the
// TOP of the 64-bit word (vbits = 64 - stored bits), the value at the bottom, zero bits in between: the word is a funnel
// shift of the left-aligned hash, it orders by (hash, value), and `word & ((1 << vbits) - 1)` is the value.
struct PackFmt {
  BijHash bij;
  int vbits;        // position of the stored hash bits = 64 - (n_eff() - drop); the value occupies the low bits below it
  int drop;         // hash bits implied by the record's level-1 region (not stored)
  int owner_bits;   // sharded path with a power-of-two number of owners: the top hash bits ARE the owner (not stored either)
  uint32_t owner;   // ... and this is the owner whose records these are (needed to recover a k-mer)
  __host__ __device__ __forceinline__ int n_eff() const { return bij.n - owner_bits; }      // hash bits below the owner bits
  // the full hash of a stored one: owner bits, the bits of level-1 bucket `region`, the stored bits
  __host__ __device__ __forceinline__ uint64_t compose(uint32_t region, uint64_t stored) const {
    uint64_t h = stored;
    if (drop > 0) h |= (uint64_t)region << (n_eff() - drop);
    if (owner_bits > 0) h |= (uint64_t)owner << n_eff();
    return h;
  }
};

// owner GPU of a key on the direct sharded path: a cheap 32-bit mix (fold, two multiplies), independent of the bit
// fields of key_hash that select the leaf buckets; the top bits of the product are mapped onto [0, n_owners)
__host__ __device__ __forceinline__ uint32_t owner_of(uint64_t key, uint32_t n_owners) {
  uint32_t x = ((uint32_t)(key >> 32) * 0x85EBCA6Bu) ^ (uint32_t)key;
  x *= 0xC2B2AE35u;
  x ^= x >> 15;
  x *= 0x9E3779B1u;
  return (uint32_t)(((uint64_t)x * n_owners) >> 32);
}

// ---------------------------------------------------------------------------------------------------
// 2-bit alphabet, MSB-first.  ASCII -> code: (c>>1)&3 gives A0 C1 T2 G3 for both cases; t ^ (t>>1) swaps the
// last two so that A<C<G<T numerically as in the reference's string comparison, and complement is x^3.
__device__ __forceinline__ uint32_t pack4_top(uint32_t x) {   // 4 ASCII bytes (first char = lowest byte) -> 8 bits in the TOP byte
  uint32_t t = (x >> 1) & 0x03030303u;
  uint32_t c = t ^ ((t >> 1) & 0x01010101u);
  // one multiply gathers the four 2-bit codes: byte i (bits 8i, 8i+1) moves to bits 30-2i.. of the product; the partial
  // products occupy distinct bit positions (no carries) and nothing else lands in the top byte
  return c * 0x40100401u;
}
__device__ __forceinline__ uint32_t pack4(uint32_t x) { return pack4_top(x) >> 24; }
__device__ __forceinline__ uint32_t pack16(uint4 v) {     // 16 ASCII bytes -> 32 bits, first char in the top bits
  const uint32_t hi = __byte_perm(pack4_top(v.y), pack4_top(v.x), 0x0073);   // byte 1 = code byte of v.x, byte 0 = of v.y
  const uint32_t lo = __byte_perm(pack4_top(v.w), pack4_top(v.z), 0x0073);
  return __byte_perm(lo, hi, 0x5410);
}

// reverse complement of a right-aligned 2k-bit value: complement all bits, reverse the 64 bits (BREV), swap
// the two bits inside each base, shift the k bases that moved to the top back down.
__device__ __forceinline__ uint64_t revcomp2(uint64_t fwd, int k) {
  uint64_t b = __brevll(~fwd);
  b = ((b >> 1) & 0x5555555555555555ull) | ((b & 0x5555555555555555ull) << 1);
  return b >> (64 - 2 * k);
}

// ---------------------------------------------------------------------------------------------------
// The part buffer as the kernels see it.  The device copy of `ascii` is padded with >= 64 zero bytes.
struct Parts {
  const uint8_t* ascii;
  const int64_t* off;          // n_parts + 1
  const uint32_t* tile_first;  // per byte tile: index of the part containing its first byte (n_tiles + 1 entries)
  int64_t n_parts;
  int64_t total_len;
  int uniform_len;             // > 0: every part has exactly this length (fast path, no offsets lookup)
  int64_t part_lo, part_hi;    // only windows of parts in [part_lo, part_hi) are produced (this rank's shard)
  int64_t tile_lo;             // first byte tile of the shard (kernels add it to blockIdx.x)
};

constexpr int kTileBytes = 4096;      // window start positions handled by one CTA
constexpr int kTileThreads = 512;
constexpr int kTileItems = kTileBytes / kTileThreads;
constexpr int kPackedWords = kTileBytes / 16 + 3;   // tile + 32-base halo, rounded up
constexpr int kSliceCap = 1024;                     // offsets cached per tile (parts overlapping the tile)

constexpr int kTileFlags = 64;                      // per-part flag bytes cached per tile (uniform parts of >= 70 bases)
constexpr int kTileFlagMinLen = 70;
struct TileStage {
  uint32_t packed[kPackedWords + 1];
  int64_t slice[kSliceCap + 2];
  int64_t first_part;
  int n_slice;                 // number of parts cached in `slice` (0 = uniform or overflow -> global search)
  int n_flag;                  // > 0: part_flag[i] is the flag byte of part first_part + i
  uint8_t part_flag[kTileFlags];
};

// What one thread fetches from global memory for a tile; kept in registers so that the loads of the NEXT tile can be in
// flight while the current tile is written out (tile_fetch / tile_commit), or consumed at once (tile_load).
struct TileRegs { uint4 v; uint32_t flag; };

// Stage one byte tile: 2-bit pack the bytes [g_lo, g_lo + 4096 + 48) and cache the offsets of its parts.
// `tid` is the thread's index inside the group of kTileThreads threads that works on this tile.
static_assert(kPackedWords <= kTileThreads, "one 16-byte word per thread");
__device__ __forceinline__ void tile_fetch(const Parts& P, int64_t tile, int tid, const uint8_t* flags, TileRegs& r) {
  const int64_t g_lo = tile * kTileBytes;
  const int64_t g = g_lo + 16 * (int64_t)tid;
  r.v = make_uint4(0, 0, 0, 0);
  r.flag = 0;
  if (tid < kPackedWords && g < P.total_len + 48) r.v = __ldg(reinterpret_cast<const uint4*>(P.ascii + g));   // buffer is padded and 16B aligned
  if (flags != nullptr && P.uniform_len >= kTileFlagMinLen && tid < kTileFlags) {
    const int64_t p = g_lo / P.uniform_len + tid;
    if (p < P.n_parts) r.flag = flags[p];
  }
}

__device__ __forceinline__ void tile_commit(const Parts& P, int64_t tile, TileStage& s, int tid, const uint8_t* flags, const TileRegs& r) {
  const int64_t g_lo = tile * kTileBytes;
  if (tid < kPackedWords) s.packed[tid] = pack16(r.v);
  const bool cached_flags = flags != nullptr && P.uniform_len >= kTileFlagMinLen;
  if (cached_flags && tid < kTileFlags) s.part_flag[tid] = (uint8_t)r.flag;
  if (tid == 0) s.n_flag = cached_flags ? kTileFlags : 0;
  if (P.uniform_len == 0) {
    int64_t p_lo = P.tile_first[tile];
    int64_t p_hi = P.tile_first[tile + 1];            // first part of the next tile: covers every part of this one
    int64_t cnt = p_hi - p_lo + 1;
    if (cnt > kSliceCap) cnt = 0;
    if (tid == 0) { s.first_part = p_lo; s.n_slice = (int)cnt; }
    for (int i = tid; i <= cnt && cnt > 0; i += kTileThreads) s.slice[i] = P.off[p_lo + i];
  } else if (tid == 0) {
    s.first_part = g_lo / P.uniform_len;
    s.n_slice = 0;
  }
}

__device__ __forceinline__ void tile_load(const Parts& P, int64_t tile, TileStage& s, int tid, const uint8_t* flags = nullptr) {
  TileRegs r;
  tile_fetch(P, tile, tid, flags, r);
  tile_commit(P, tile, s, tid, flags, r);
}

// Part containing byte g and the end offset of that part.  Requires g < total_len.
__device__ __forceinline__ void tile_locate(const Parts& P, const TileStage& s, int64_t g, int64_t& part, int64_t& end) {
  if (P.uniform_len > 0) {
    int64_t p0 = s.first_part;
    uint32_t rel = (uint32_t)(g - p0 * P.uniform_len);
    part = p0 + rel / (uint32_t)P.uniform_len;
    end = (part + 1) * P.uniform_len;
    return;
  }
  if (s.n_slice > 0) {
    int lo = 0, hi = s.n_slice;                // find last i in [0, n_slice] with slice[i] <= g
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (s.slice[mid] <= g) lo = mid; else hi = mid;
    }
    part = s.first_part + lo;
    end = s.slice[lo + 1];
    return;
  }
  int64_t lo = s.first_part, hi = P.n_parts;   // overflow path: search the global offsets
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (P.off[mid] <= g) lo = mid; else hi = mid;
  }
  part = lo;
  end = P.off[lo + 1];
}

// Forward value of the k-window that starts `c` bases into the staged tile (right-aligned 2k bits).
__device__ __forceinline__ uint64_t tile_window(const TileStage& s, int c, int k) {
  int w = c >> 4;
  int sh = (c & 15) * 2;
  uint32_t x0 = s.packed[w], x1 = s.packed[w + 1], x2 = s.packed[w + 2];
  uint32_t hi = __funnelshift_l(x1, x0, sh);
  uint32_t lo = __funnelshift_l(x2, x1, sh);
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return v >> (64 - 2 * k);
}

// Canonical k-mer starting at byte g of part-relative validity: returns false when no window starts at g.
struct Window {
  uint64_t canon;
  uint32_t part;
  bool palindrome;
};
__device__ __forceinline__ bool tile_canonical(const Parts& P, const TileStage& s, int64_t tile, int idx, int k, Window& w) {
  int64_t g = tile * kTileBytes + idx;
  if (g >= P.total_len) return false;
  int64_t part, end;
  tile_locate(P, s, g, part, end);
  if (g + k > end) return false;
  uint64_t fwd = tile_window(s, idx, k);
  uint64_t rc = revcomp2(fwd, k);
  w.canon = fwd < rc ? fwd : rc;
  w.palindrome = fwd == rc;
  w.part = (uint32_t)part;
  return true;
}

// Blocked iteration: thread t owns the kTileItems consecutive window start positions 8t..8t+7 of the tile.
// The forward value and its reverse complement ROLL from one window to the next (one shift-or each).  Part
// bookkeeping is 32-bit and tile relative: for uniform-length parts a (quotient, remainder) pair is advanced
// per position, for variable-length parts the cached offsets slice is walked.  A part is skipped when it is outside
// the shard [part_lo, part_hi) or -- with SKIP_FLAGGED -- when its flag byte is non-zero; both tests are made
// once per part, not per window.  f(e, canon, part, j, is_palindrome) is called for every valid window (j = index of
// the window inside its part), e ascending.
template <bool SKIP_FLAGGED, typename F>
__device__ __forceinline__ void tile_windows_blocked(const Parts& P, const TileStage& s, int64_t tile, int k, const uint8_t* flags, int tid, F f) {
  const int c0 = tid * kTileItems;
  const int64_t g_lo = tile * kTileBytes;
  if (g_lo + c0 >= P.total_len) return;
  const int w = c0 >> 4, sh = (c0 & 15) * 2;                  // sh is 0 or 16
  const uint32_t x0 = s.packed[w], x1 = s.packed[w + 1], x2 = s.packed[w + 2];
  uint64_t v = ((uint64_t)__funnelshift_l(x1, x0, sh) << 32) | __funnelshift_l(x2, x1, sh);   // bases c0 .. c0+31
  const uint32_t next16 = (x2 >> (16 - sh)) & 0xffffu;       // bases c0+32 .. c0+39
  const int down = 64 - 2 * k, top = 2 * k - 2;
  uint64_t fwd = v >> down;
  uint64_t rc = revcomp2(fwd, k);
  auto part_usable = [&](int64_t part) -> int {
    if (part < P.part_lo || part >= P.part_hi) return 0;
    if (!SKIP_FLAGGED) return 1;
    if (s.n_flag > 0) return s.part_flag[(int)(part - s.first_part)] == 0 ? 1 : 0;    // flag bytes cached with the tile
    return flags[part] == 0 ? 1 : 0;
  };
  if (P.uniform_len > 0) {
    const uint32_t L = (uint32_t)P.uniform_len;
    const int64_t p0 = s.first_part;
    const uint32_t rel = (uint32_t)(g_lo - p0 * (int64_t)L) + (uint32_t)c0;
    uint32_t q = rel / L;
    uint32_t r = rel - q * L;                                 // offset of the position inside its part
    int64_t part = p0 + q;
    if (L < (uint32_t)k) return;
    const uint32_t last = L - (uint32_t)k;                   // last window start inside a part
    int ok = part_usable(part);
#pragma unroll
    for (int e = 0; e < kTileItems; ++e) {
      if (e > 0) {
        v = (v << 2) | ((next16 >> (16 - 2 * e)) & 3u);
        fwd = v >> down;
        rc = (rc >> 2) | ((uint64_t)(3u - ((uint32_t)fwd & 3u)) << top);
        if (++r == L) { r = 0; ++part; ok = part_usable(part); }
      }
      if (ok != 0 && r <= last) f(e, fwd < rc ? fwd : rc, (uint32_t)part, r, fwd == rc);
    }
    return;
  }
  // variable-length parts: walk the offsets (tile-relative, clamped to int32)
  const int64_t remaining = P.total_len - g_lo;
  const int limit = remaining > kTileBytes ? kTileBytes : (int)remaining;      // positions >= limit hold no data
  int64_t part, end64;
  tile_locate(P, s, g_lo + c0, part, end64);
  auto rel_end = [&](int64_t e64) -> int { int64_t d = e64 - g_lo; return d > (1 << 30) ? (1 << 30) : (int)d; };
  int end = rel_end(end64);
  int start = (int)((s.n_slice > 0 ? s.slice[part - s.first_part] : P.off[part]) - g_lo);   // may be negative
  int ok = part_usable(part);
  int c = c0;
#pragma unroll
  for (int e = 0; e < kTileItems; ++e) {
    if (e > 0) {
      v = (v << 2) | ((next16 >> (16 - 2 * e)) & 3u);
      fwd = v >> down;
      rc = (rc >> 2) | ((uint64_t)(3u - ((uint32_t)fwd & 3u)) << top);
      ++c;
      if (c >= limit) break;
      if (c >= end) {
        do {                                                   // next non-empty part
          ++part;
          start = end;
          end = rel_end(s.n_slice > 0 ? s.slice[part - s.first_part + 1] : P.off[part + 1]);
        } while (c >= end);
        ok = part_usable(part);
      }
    }
    if (ok != 0 && c + k <= end) f(e, fwd < rc ? fwd : rc, (uint32_t)part, (uint32_t)(c - start), fwd == rc);
  }
}

// first part of every byte tile (variable-length parts only): one thread per tile
__global__ void k_tile_first_part(const int64_t* off, int64_t n_parts, int64_t total_len, int64_t n_tiles, uint32_t* tile_first) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  if (t == n_tiles) { tile_first[t] = (uint32_t)(n_parts > 0 ? n_parts - 1 : 0); return; }
  int64_t g = t * kTileBytes;
  int64_t lo = 0, hi = n_parts;              // last p with off[p] <= g
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (off[mid] <= g) lo = mid; else hi = mid;
  }
  tile_first[t] = (uint32_t)lo;
}

__device__ __forceinline__ uint32_t atomic_cas(uint32_t* a, uint32_t cmp, uint32_t v) { return atomicCAS(a, cmp, v); }
__device__ __forceinline__ uint64_t atomic_cas(uint64_t* a, uint64_t cmp, uint64_t v) {
  return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long*>(a), (unsigned long long)cmp, (unsigned long long)v);
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace nrp
