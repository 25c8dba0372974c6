// General path: parts over ANY byte alphabet and / or windows longer than 32 bases (keys wider than 64 bits).
//
// The 2-bit kernels cover A/C/G/T and k <= 32.  The reference itself has no such limits (nrpcalc/base/utils.py:10,37-50,
// hgraph.py:23-94): `get_comp` maps 'ATGCU' -> 'TACGA' and leaves every other character alone -- so an RNA part keeps its U's,
// the reverse complement of a U is an A (and NOT the other way round), an N is its own complement -- keys are compared as
// strings, and k is unbounded.  This path restates exactly that on bytes:
//
//   * a window is identified by a 64-bit hash of its bytes; the canonical key of a window w is the hash of min(w, rc(w)) in
//     byte order (= Python's string order on ASCII);
//   * hashes only PROPOSE equalities -- every equality that changes a result is verified on the bytes themselves: inside a
//     part by the extraction kernel, between parts by k_generic_verify on the sorted candidate records.  A pair of different
//     strings with one hash raises a status bit and the host repeats the build with another seed.  So the result is exact;
//   * the internal-repeat test is the reference's own (hgraph.py:50-69), which for an alphabet whose complement is not an
//     involution differs from "two windows share a canonical key": a part is dropped iff two of its FORWARD windows are equal
//     (direct repeat) or the reverse complement of one window equals a forward window (inverted repeat; a window equal to its own
//     reverse complement is the case i == j);
//   * the background test follows kmerSetDB.__contains__ / multicheck (kmerSetDB.py:177-207): U -> T first, then the canonical
//     K-window; windows that then still hold other characters can only match keys that the host layer keeps as strings and are
//     resolved there (preset flags).
//
// After the extraction the records (canonical hash, (part << j_bits) | window) take the ordinary bulk path as a flat list of
// 64-bit keys (hash-radix partition, duplicate filter, grouping, edges).
#pragma once
#include "common.cuh"
#include "stages.cuh"

namespace nrp {

constexpr int kGenThreads = 128;
constexpr int kGenMaxWindows = 4096;          // windows of one part (shared-memory tables of the in-part tests)
constexpr uint32_t kGenPartCollision = 1u;    // status: two different windows of one part share a hash
constexpr uint32_t kGenGroupCollision = 2u;   // status: two records with equal keys but different canonical strings

__device__ __forceinline__ uint8_t gen_comp(uint8_t c) {        // utils.py:10 -- str.maketrans('ATGCU', 'TACGA')
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    case 'U': return 'A';
    default: return c;
  }
}

__device__ __forceinline__ uint64_t gen_finish(uint64_t h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h == ~0ull ? 0ull : h;                                 // all ones marks an empty slot of the device hash sets
}

struct GenWindow { uint64_t hf, hr; int cmp; };                 // hashes of w and rc(w); cmp < 0: w < rc(w), 0: equal, > 0: w > rc(w)

// seed: the low 58 bits seed the hash; the top 6 bits, when non-zero, keep only that many hash bits (tests of the collision path)
__device__ __forceinline__ GenWindow gen_window(const uint8_t* w, int k, uint64_t seed) {
  constexpr uint64_t P = 0x100000001b3ull;
  uint64_t hf = seed, hr = seed;
  int cmp = 0;
  for (int i = 0; i < k; ++i) {
    const uint8_t a = w[i], b = gen_comp(w[k - 1 - i]);
    hf = (hf ^ a) * P;
    hr = (hr ^ b) * P;
    if (cmp == 0 && a != b) cmp = a < b ? -1 : 1;
  }
  const int keep = (int)(seed >> 58);
  const uint64_t m = keep ? ((1ull << keep) - 1ull) : ~0ull;
  return GenWindow{gen_finish(hf) & m, gen_finish(hr) & m, cmp};
}

// canonical 2-bit K-mer of a window after U -> T (kmerSetDB.py:183-184); false when another character remains
__device__ __forceinline__ bool gen_bg_canon(const uint8_t* w, int K, uint64_t* canon) {
  uint64_t fwd = 0, rc = 0;
  for (int i = 0; i < K; ++i) {
    uint8_t c = w[i];
    if (c == 'U') c = 'T';
    uint32_t code;
    switch (c) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break; default: return false; }
    fwd = (fwd << 2) | code;
    rc |= (uint64_t)(3u - code) << (2 * i);
  }
  *canon = fwd < rc ? fwd : rc;
  return true;
}

struct GenArgs {
  const uint8_t* ascii; const int64_t* off; int64_t n_parts;
  int k; uint64_t seed; int internal_repeats; int j_bits; uint32_t part_base;
  const void* bg_table; int bg_log_slots, bg_key_bytes, bg_K;      // bg_K == 0: no background probe on the device
  uint8_t* flags;                                                   // in: flags the host preset; out: |= what this kernel finds
  uint64_t* out_keys; uint32_t* out_vals; unsigned long long* n_out;
  uint32_t* status;
  int max_windows, table_slots;                                     // table_slots: power of two >= 2 * max_windows
};

// One CTA per part (grid-stride).  Dynamic shared memory: F[max_windows], R[max_windows] (64-bit hashes), C[max_windows] (sign of
// the comparison), table[table_slots] (window indices, open addressing on F).
__global__ void __launch_bounds__(kGenThreads) k_generic_extract(GenArgs a) {
  extern __shared__ __align__(16) unsigned char gen_smem[];
  uint64_t* F = reinterpret_cast<uint64_t*>(gen_smem);
  uint64_t* R = F + a.max_windows;
  uint32_t* table = reinterpret_cast<uint32_t*>(R + a.max_windows);
  int8_t* C = reinterpret_cast<int8_t*>(table + a.table_slots);
  __shared__ uint32_t s_found;
  __shared__ unsigned long long s_base;
  const uint32_t mask = (uint32_t)a.table_slots - 1u;
  const int k = a.k;
  for (int64_t p = blockIdx.x; p < a.n_parts; p += gridDim.x) {
    __syncthreads();                                                // the previous part is done with the shared arrays
    const uint8_t* seq = a.ascii + a.off[p];
    const int64_t len = a.off[p + 1] - a.off[p];
    const int W = len >= k ? (int)(len - k + 1) : 0;
    if (W == 0 || a.flags[p] != 0) continue;                       // uniform across the CTA
    if (threadIdx.x == 0) s_found = 0;
    if (!a.internal_repeats)
      for (int i = threadIdx.x; i < a.table_slots; i += kGenThreads) table[i] = 0xffffffffu;
    for (int j = threadIdx.x; j < W; j += kGenThreads) {
      const GenWindow g = gen_window(seq + j, k, a.seed);
      F[j] = g.hf; R[j] = g.hr; C[j] = (int8_t)g.cmp;
    }
    __syncthreads();                                                // hashes, the cleared table and s_found are in place
    uint32_t found = 0;
    if (!a.internal_repeats) {
      // direct repeat: two equal forward windows (hgraph.py:53-56)
      for (int j = threadIdx.x; j < W; j += kGenThreads) {
        uint32_t h = (uint32_t)F[j] & mask;
        for (;;) {
          const uint32_t old = atomicCAS(&table[h], 0xffffffffu, (uint32_t)j);
          if (old == 0xffffffffu) break;
          if (F[old] == F[j]) {
            bool same = true;
            for (int t = 0; same && t < k; ++t) same = seq[old + t] == seq[j + t];
            if (same) found |= kFlagRepeat; else atomicOr(a.status, kGenPartCollision);
            break;
          }
          h = (h + 1u) & mask;
        }
      }
      __syncthreads();
      // inverted repeat: the reverse complement of window i is a forward window j (hgraph.py:58-69; i == j: a palindrome)
      for (int i = threadIdx.x; i < W; i += kGenThreads) {
        uint32_t h = (uint32_t)R[i] & mask;
        for (;;) {
          const uint32_t cur = table[h];
          if (cur == 0xffffffffu) break;
          if (F[cur] == R[i]) {
            bool same = true;
            for (int t = 0; same && t < k; ++t) same = gen_comp(seq[i + k - 1 - t]) == seq[cur + t];
            if (same) found |= (cur == (uint32_t)i) ? kFlagPalindrome : kFlagRepeat; else atomicOr(a.status, kGenPartCollision);
            break;
          }
          h = (h + 1u) & mask;
        }
      }
    }
    if (a.bg_K > 0 && len >= a.bg_K) {
      for (int t = threadIdx.x; t + a.bg_K <= len; t += kGenThreads) {
        uint64_t canon;
        if (!gen_bg_canon(seq + t, a.bg_K, &canon)) continue;
        const bool hit = a.bg_key_bytes == 4 ? bg_contains<uint32_t>(reinterpret_cast<const uint32_t*>(a.bg_table), a.bg_log_slots, canon)
                                             : bg_contains<uint64_t>(reinterpret_cast<const uint64_t*>(a.bg_table), a.bg_log_slots, canon);
        if (hit) found |= kFlagBackground;
      }
    }
    if (found != 0) atomicOr(&s_found, found);
    __syncthreads();
    if (s_found != 0) {
      if (threadIdx.x == 0) a.flags[p] |= (uint8_t)s_found;
      continue;
    }
    if (threadIdx.x == 0) s_base = atomicAdd(a.n_out, (unsigned long long)W);
    __syncthreads();
    const unsigned long long base = s_base;
    const uint32_t v0 = (a.part_base + (uint32_t)p) << a.j_bits;
    for (int j = threadIdx.x; j < W; j += kGenThreads) {
      a.out_keys[base + j] = C[j] <= 0 ? F[j] : R[j];
      a.out_vals[base + j] = v0 | (uint32_t)j;
    }
  }
}

// byte t of the canonical string of the k-window w: w itself when w <= rc(w), else rc(w)
__device__ __forceinline__ uint8_t gen_canon_byte(const uint8_t* w, int k, int cmp, int t) { return cmp <= 0 ? w[t] : gen_comp(w[k - 1 - t]); }
__device__ __forceinline__ int gen_compare(const uint8_t* w, int k) {
  for (int i = 0; i < k; ++i) {
    const uint8_t a = w[i], b = gen_comp(w[k - 1 - i]);
    if (a != b) return a < b ? -1 : 1;
  }
  return 0;
}

// candidate records ordered by (key, value): neighbours with equal keys must have equal canonical strings
__global__ void k_generic_verify(const uint8_t* ascii, const int64_t* off, const uint64_t* keys, const uint32_t* vals, int64_t n, int k, int j_bits,
                                 uint32_t part_base, uint32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1 || i >= n || keys[i] != keys[i - 1]) return;
  const uint32_t jm = (1u << j_bits) - 1u;
  const uint8_t* wa = ascii + off[(vals[i] >> j_bits) - part_base] + (vals[i] & jm);
  const uint8_t* wb = ascii + off[(vals[i - 1] >> j_bits) - part_base] + (vals[i - 1] & jm);
  const int ca = gen_compare(wa, k), cb = gen_compare(wb, k);
  for (int t = 0; t < k; ++t)
    if (gen_canon_byte(wa, k, ca, t) != gen_canon_byte(wb, k, cb, t)) { atomicOr(status, kGenGroupCollision); return; }
}

// k_last_single for this path: last window of each surviving part whose canonical hash is not a shared key
__global__ void k_generic_last_single(const uint8_t* ascii, const int64_t* off, const uint8_t* flags, int64_t n_parts, int k, uint64_t seed,
                                      const uint64_t* shared_table, int log_slots, int32_t* last_single) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_parts) return;
  const int64_t len = off[p + 1] - off[p];
  int32_t result = -1;
  if (len >= k && flags[p] == 0) {
    const uint8_t* seq = ascii + off[p];
    for (int64_t j = len - k; j >= 0; --j) {
      const GenWindow g = gen_window(seq + j, k, seed);
      const uint64_t key = g.cmp <= 0 ? g.hf : g.hr;
      if (!(log_slots > 0 && bg_contains<uint64_t>(shared_table, log_slots, key))) { result = (int32_t)j; break; }
    }
  }
  last_single[p] = result;
}

}  // namespace nrp
