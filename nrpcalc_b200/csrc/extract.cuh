// Extraction fused with the first split for parts of ONE length (every BASELINE shape but the variable-length one):
// the part-aligned kernel.
//
// Replaces reference nrpcalc/base/utils.py:37-50 (stream_kmers, get_revcomp, min) + hgraph.py:82-89 (table insert) for
// the bulk of the records, like k_split_from_ascii (partition.cuh), which stays the general kernel (variable lengths,
// unpacked records, the exact path).  What is different here:
//
//   * threads are mapped to WINDOWS OF ONE PART, not to byte positions: a part with W windows is cut into
//     Tp = ceil(W / 8) chunks of 8 consecutive windows, thread t of a tile owns chunk t % Tp of part t / Tp.  No thread
//     ever crosses a part boundary, so the window loop has no part bookkeeping, no divergence (the byte-position kernel
//     executes its part-change branch in almost every warp iteration: 32 lanes x 8 % crossing probability) and no
//     positions without a window (16 % of the positions of a 100-bp part at k = 17);
//   * forward and reverse-complement values roll LEFT-ALIGNED (the k-mer's first base in the top bits of the word): bits
//     that leave the window fall off the top, "mod 2^2k" of the bijective hash is free, the bucket digit is one shift of
//     the high word and the packed record is one funnel shift -- no masks;
//   * k <= 16 runs in 32-bit arithmetic (KEY64 = false);
//   * windows without a record (the tail of a part's last chunk, flagged parts, idle threads) take the same instructions
//     and count in a per-lane dump counter: no branches around the shared-memory atomic.
//
// About 60 thread instructions per window against 163 of the byte-position kernel (profiles/ncu_full_r02*.txt).
#pragma once
#include "partition.cuh"
#include "stages.cuh"

namespace nrp {

constexpr int kUniMaxWords = 1264;     // packed 16-base words of one tile: parts_per_tile * L <= 512 * 39 bytes, + alignment + halo

// host-computed constants of the mapping
struct UniPlan {
  uint32_t L;        // part length
  uint32_t W;        // windows per part = L - k + 1 (>= 1)
  uint32_t Tp;       // threads per part = ceil(W / 8)
  uint32_t G;        // parts per tile = 512 / Tp
  uint32_t inv_Tp;   // ceil(2^20 / Tp): t / Tp == (t * inv_Tp) >> 20 for t < 512
};
inline bool uni_plan(int64_t L, int k, UniPlan* up) {
  if (L < k || L >= (1 << 20)) return false;
  const int64_t W = L - k + 1;
  if (W > 4096) return false;                      // more than 512 chunks: the byte-position kernel
  up->L = (uint32_t)L; up->W = (uint32_t)W;
  up->Tp = (uint32_t)((W + 7) / 8);
  up->G = 512u / up->Tp;
  up->inv_Tp = ((1u << 20) + up->Tp - 1u) / up->Tp;
  return (int64_t)up->G * L + 64 <= (int64_t)(kUniMaxWords - 8) * 16;
}

// Bit map in front of the background table for the folded probe: a blocked Bloom filter with ONE 32-bit word per probe.  The word
// is selected by the top (log_bits - 5) bits of the bijective hash of the k-mer, three bit positions inside it by a 32-bit mix of
// the top 32 hash bits: at >= 16 bits per key a window that is not in the set passes all three with probability ~1e-3.
__host__ __device__ __forceinline__ uint32_t bg_bit_mask(uint32_t top_hash) {
  const uint32_t m = top_hash * 0x9E3779B1u;
  return (1u << (m >> 27)) | (1u << ((m >> 22) & 31u)) | (1u << ((m >> 17) & 31u));
}
__global__ void k_bg_bitmap(const uint64_t* keys, int64_t n, BijHash bij, int log_bits, uint32_t* bits) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t h = bij.fwd(keys[i]);
    const uint32_t top = bij.n >= 32 ? (uint32_t)(h >> (bij.n - 32)) : (uint32_t)(h << (32 - bij.n));     // the hash left-aligned in 32 bits
    const uint32_t word = log_bits > 5 ? top >> (37 - log_bits) : 0u;
    atomicOr(bits + word, bg_bit_mask(top));
  }
}

struct alignas(16) UniSmem {
  SplitSmem<uint64_t, true> split;
  uint32_t packed[kUniMaxWords];
  uint32_t pflag[kSplitThreads];       // FOLD: flag bits found by the threads of one part (a tile holds at most 512 parts)
};

// FOLD: the exclusion flags that need no grouping are computed HERE instead of in a pass of their own over the parts
// (k_flag_windows; reference hgraph.py:59-79): a window equal to its own reverse complement (even k, internal_repeats off) and --
// when the background's K equals k -- background membership of the canonical k-mer (kmerSetDB.py:177-207).  The probe is one
// word of a blocked Bloom filter indexed by the SAME bijective hash the split uses (k_bg_bitmap), loaded three windows ahead of
// its test; the few windows that pass (~0.1 % + the true hits) are confirmed in the exact open-addressing table after the window
// loop.  The threads of a part OR their findings in shared memory; after one more CTA barrier a flagged part emits no record and
// its first thread writes the flag byte.
struct FoldArgs {
  uint8_t* flags;              // per-part flag bytes (read-modify-write by the part's first thread only)
  const uint32_t* bg_bits;     // nullptr: no background probe
  const void* bg_table;        // canonical k-mers, open addressing (k_bg_insert)
  int bg_log_bits, bg_log_slots, bg_key_bytes;
  int check_pal;
};

// MODE 0: digit = top r1 bits of the hash (single GPU, fixed regions).  MODE 2: digit = (owner, top r1 bits), stored into the
// owner's region of this (bucket, sender) over NVLink (direct sharded path); the owner is the top owner_bits of the hash when
// pf.owner_bits > 0, else owner_of(k-mer).
template <bool KEY64, int MODE, bool FOLD = false>
__global__ void __launch_bounds__(kSplitThreads, 2) k_split_uniform(const uint8_t* __restrict__ ascii, const uint8_t* flags, UniPlan up,
                                                                    int64_t part_lo, int64_t part_hi, int k, int r1, int n_owners, uint32_t part_base,
                                                                    int j_bits, PackFmt pf, uint32_t* cursor, uint64_t* out_keys, RegionLimit lim,
                                                                    PeerBuffers peers, FoldArgs fold = FoldArgs{}) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UniSmem& sm = *reinterpret_cast<UniSmem*>(smem_raw);
  SplitSmem<uint64_t, true>& s = sm.split;
  const int t = threadIdx.x;
  const int64_t P0 = part_lo + (int64_t)blockIdx.x * up.G;
  const int64_t P1 = min(P0 + (int64_t)up.G, part_hi);
  const int64_t B0 = P0 * up.L, B1 = P1 * up.L;
  const int64_t A0 = B0 & ~15ll;
  const int n_chunks = (int)((B1 - A0 + 15) >> 4) + 3;          // + the words the last threads read past their windows (buffer is padded)
  // global loads first (at most three 16-byte words per thread), the per-thread constants while they are in flight
  uint4 raw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = t + i * kSplitThreads;
    raw[i] = make_uint4(0, 0, 0, 0);
    if (c < n_chunks) raw[i] = __ldg(reinterpret_cast<const uint4*>(ascii + A0) + c);
  }
  const uint32_t pl = ((uint32_t)t * up.inv_Tp) >> 20;          // part of this thread inside the tile
  const uint32_t j0 = ((uint32_t)t - pl * up.Tp) * (uint32_t)kSplitItems;   // its first window inside the part
  const int64_t part = P0 + pl;
  int n_e = 0;                                                   // windows of this thread (0: idle)
  if (pl < up.G && part < P1) {
    n_e = (int)min((uint32_t)kSplitItems, up.W - j0);
    if (flags != nullptr && flags[part] != 0) n_e = 0;          // part excluded by the flag pass (palindrome / background)
  }
  s.cnt[t] = 0;
  if (FOLD) sm.pflag[t] = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = t + i * kSplitThreads;
    if (c < n_chunks) sm.packed[c] = pack16(raw[i]);
  }
  __syncthreads();

  const int j = KEY64 ? 64 - 2 * k : 32 - 2 * k;                 // free low bits of a left-aligned k-mer
  const uint32_t rel = n_e > 0 ? (uint32_t)(B0 - A0) + pl * up.L + j0 : 0u;     // base index of the first window in the packed tile
  const int w = (int)(rel >> 4), sh = (int)(rel & 15u) * 2;
  const uint32_t relT = rel + (uint32_t)k - 1u;                  // T: the 16 bases from the LAST base of window 0 on
  const int wT = (int)(relT >> 4), shT = (int)(relT & 15u) * 2;
  const uint32_t T = __funnelshift_l(sm.packed[wT + 1], sm.packed[wT], shT);
  const uint32_t TF = __funnelshift_l(T, T, (j + 2) & 31);       // rotated: base e of T at bits (j+1, j) after e more 2-bit rotations
  const uint32_t TR = ~T;                                        // complements; base e at bits (31, 30) after e 2-bit rotations
  const uint32_t in_mask = 3u << j;
  const uint32_t low_clear = ~((1u << j) - 1u);                  // clears the free low bits (of the low word when KEY64)
  const int s_fold = k;                                          // BijHash::s = (2k + 1) / 2
  const int dbits = r1 + ((MODE == 2 && pf.owner_bits > 0) ? pf.owner_bits : 0);
  const int sdrop = pf.owner_bits + pf.drop;                     // top hash bits a record does not store
  const bool mix_owner = MODE == 2 && pf.owner_bits == 0;        // owner from owner_of(k-mer), not from the hash
  const uint32_t step = j_bits ? 1u : 0u;
  const uint32_t v0 = ((part_base + (uint32_t)part) << j_bits) | (j_bits ? j0 : 0u);
  const uint32_t lane = (uint32_t)t & 31u;
  constexpr uint32_t C1 = (uint32_t)BijHash::kC1, C2 = (uint32_t)BijHash::kC2;
  static_assert(BijHash::kC1 < (1ull << 32) && BijHash::kC2 < (1ull << 32), "32-bit multipliers");

  uint64_t key[kSplitItems];
  uint32_t val[kSplitItems], dr[kSplitItems];
  // FOLD: flag bits this thread found.  The bit-map word of window e is tested while window e + 3 is computed (three loads in
  // flight per thread); a window whose bits are all set becomes a SUSPECT and is confirmed in the exact table after the loop --
  // inside the loop the confirmation would run for whole warps although a fraction of a percent of the windows need it.
  uint32_t fbits = 0, suspects = 0;
  uint32_t bw0 = 0, bw1 = 0, bw2 = 0, bm0 = 1, bm1 = 1, bm2 = 1;     // (word, mask) of the three windows in flight
  const bool probe = FOLD && fold.bg_bits != nullptr;
  auto canon_at = [&](int e) -> uint64_t {                       // canonical k-mer of window e, right-aligned (rare path)
    const uint32_t c = rel + (uint32_t)e;
    const int cw = (int)(c >> 4), csh = (int)(c & 15u) * 2;
    const uint32_t y0 = sm.packed[cw], y1 = sm.packed[cw + 1], y2 = sm.packed[cw + 2];
    const uint64_t v = ((((uint64_t)__funnelshift_l(y1, y0, csh)) << 32) | __funnelshift_l(y2, y1, csh)) >> (64 - 2 * k);
    const uint64_t rc = revcomp2(v, k);
    return v < rc ? v : rc;
  };
  // one step of the three-deep probe pipeline: test the word loaded three windows ago, then issue this window's load
  auto probe_step = [&](int e, uint32_t top_hash, bool valid) {    // top_hash: the hash left-aligned in 32 bits
    if (e >= 3) suspects |= ((bw0 & bm0) == bm0 ? 1u : 0u) << (e - 3);
    bw0 = bw1; bm0 = bm1; bw1 = bw2; bm1 = bm2;
    bm2 = bg_bit_mask(top_hash);
    bw2 = valid ? __ldg(fold.bg_bits + __funnelshift_rc(top_hash, 0u, 37 - fold.bg_log_bits)) : 0u;
  };
  if (KEY64) {
    const uint32_t x0 = sm.packed[w], x1 = sm.packed[w + 1], x2 = sm.packed[w + 2];
    uint32_t fh = __funnelshift_l(x1, x0, sh), fl = __funnelshift_l(x2, x1, sh) & low_clear;    // window 0, left-aligned
    uint32_t rh, rl;
    {
      const uint64_t f0 = ((uint64_t)fh << 32) | fl;
      const uint64_t r0 = revcomp2(f0 >> j, k) << j;
      rh = (uint32_t)(r0 >> 32); rl = (uint32_t)r0;
    }
#pragma unroll
    for (int e = 0; e < kSplitItems; ++e) {
      if (e > 0) {                                               // roll both strands by one base
        const uint32_t in_f = __funnelshift_l(TF, TF, 2 * e) & in_mask;
        const uint32_t in_r = __funnelshift_l(TR, TR, 2 * e) & 0xC0000000u;
        fh = __funnelshift_l(fl, fh, 2);
        fl = (fl << 2) | in_f;
        rl = __funnelshift_r(rl, rh, 2) & low_clear;
        rh = (rh >> 2) | in_r;
      }
      const bool valid = e < n_e;
      const bool f_less = fh < rh || (fh == rh && fl < rl);
      const uint32_t ch_ = f_less ? fh : rh, cl = f_less ? fl : rl;     // canonical k-mer, left-aligned
      if (FOLD && fold.check_pal && valid && fh == rh && fl == rl) fbits |= kFlagPalindrome;
      // bijective hash (BijHash::fwd) in left-aligned arithmetic: multiply, fold, multiply -- mod 2^2k for free
      uint64_t p1 = (uint64_t)cl * C1;
      uint32_t hl = (uint32_t)p1, hh = ch_ * C1 + (uint32_t)(p1 >> 32);
      hl ^= __funnelshift_rc(hl, hh, s_fold) & low_clear;
      hh ^= __funnelshift_rc(hh, 0u, s_fold);
      const uint64_t p2 = (uint64_t)hl * C2;
      hl = (uint32_t)p2; hh = hh * C2 + (uint32_t)(p2 >> 32);
      uint32_t d = __funnelshift_rc(hh, 0u, 32 - dbits);
      if (mix_owner) {
        const uint64_t canon = (((uint64_t)ch_ << 32) | cl) >> j;
        d |= owner_of(canon, (uint32_t)n_owners) << r1;
      }
      if (FOLD) {
        if (probe) probe_step(e, hh, valid);
        dr[e] = d;                                               // ranked after the part's flags are known
      } else {
        const uint32_t rank = atomicAdd(&s.cnt[valid ? d : (uint32_t)kLevelBuckets + lane], 1u);
        dr[e] = valid ? ((d << 16) | rank) : 0xffffffffu;
      }
      const uint32_t kh = __funnelshift_l(hl, hh, sdrop), kl = (hl << sdrop) | (v0 + (uint32_t)e * step);
      key[e] = ((uint64_t)kh << 32) | kl;
      val[e] = 0;
    }
  } else {
    const uint32_t x0 = sm.packed[w], x1 = sm.packed[w + 1];
    uint32_t f = __funnelshift_l(x1, x0, sh) & low_clear;
    uint32_t r = (uint32_t)(revcomp2((uint64_t)(f >> j), k) << j);
#pragma unroll
    for (int e = 0; e < kSplitItems; ++e) {
      if (e > 0) {
        f = (f << 2) | (__funnelshift_l(TF, TF, 2 * e) & in_mask);
        r = ((r >> 2) & low_clear) | (__funnelshift_l(TR, TR, 2 * e) & 0xC0000000u);
      }
      const bool valid = e < n_e;
      const uint32_t c = min(f, r);
      if (FOLD && fold.check_pal && valid && f == r) fbits |= kFlagPalindrome;
      uint32_t h = c * C1;
      h ^= (h >> s_fold) & low_clear;
      h *= C2;
      uint32_t d = __funnelshift_rc(h, 0u, 32 - dbits);
      if (mix_owner) d |= owner_of((uint64_t)(c >> j), (uint32_t)n_owners) << r1;
      if (FOLD) {
        if (probe) probe_step(e, h, valid);
        dr[e] = d;
      } else {
        const uint32_t rank = atomicAdd(&s.cnt[valid ? d : (uint32_t)kLevelBuckets + lane], 1u);
        dr[e] = valid ? ((d << 16) | rank) : 0xffffffffu;
      }
      key[e] = ((uint64_t)(h << sdrop) << 32) | (v0 + (uint32_t)e * step);
      val[e] = 0;
    }
  }
  if (FOLD) {
    if (probe) {                                                 // drain the probe pipeline (windows 5, 6, 7), then confirm the suspects
      suspects |= ((bw0 & bm0) == bm0 ? 1u : 0u) << (kSplitItems - 3);
      suspects |= ((bw1 & bm1) == bm1 ? 1u : 0u) << (kSplitItems - 2);
      suspects |= ((bw2 & bm2) == bm2 ? 1u : 0u) << (kSplitItems - 1);
      while (suspects != 0) {
        const int e = __ffs(suspects) - 1;
        suspects &= suspects - 1u;
        const uint64_t canon = canon_at(e);
        const bool hit = fold.bg_key_bytes == 4 ? bg_contains<uint32_t>(reinterpret_cast<const uint32_t*>(fold.bg_table), fold.bg_log_slots, canon)
                                                : bg_contains<uint64_t>(reinterpret_cast<const uint64_t*>(fold.bg_table), fold.bg_log_slots, canon);
        if (hit) { fbits |= kFlagBackground; break; }
      }
    }
    if (fbits != 0) atomicOr(&sm.pflag[pl], fbits);
    __syncthreads();
    if (n_e > 0) {
      const uint32_t found = sm.pflag[pl];
      if (found != 0) {
        if (j0 == 0) fold.flags[part] |= (uint8_t)found;          // the part's first thread (always one of its active threads)
        n_e = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < kSplitItems; ++e) {
      const bool valid = e < n_e;
      const uint32_t d = dr[e];
      const uint32_t rank = atomicAdd(&s.cnt[valid ? d : (uint32_t)kLevelBuckets + lane], 1u);
      dr[e] = valid ? ((d << 16) | rank) : 0xffffffffu;
    }
  }
  __syncthreads();
  const int nb = MODE == 0 ? (1 << r1) : (n_owners << r1);
  split_reorder<uint64_t, true>(s, key, val, dr, nb, cursor, lim);
  if (MODE == 0) split_flush<uint64_t, true>(s, out_keys, nullptr);
  else split_flush<uint64_t, true>(s, out_keys, nullptr, &peers, r1);
}

}  // namespace nrp
