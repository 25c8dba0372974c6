// libnrpb200: context, device buffers, pipeline orchestration and the C ABI (include/nrp_b200.h).
//
// Pipeline of one build (single GPU; the sharded path reuses the same stages around an NCCL all-to-all):
//
//   flags      k_flag_windows          palindromic windows / background hits            (reads ASCII)
//   histogram  k_hist_from_ascii       leaf-bucket sizes of hash(canonical k-mer)       (reads ASCII)
//   split 1    k_split_from_ascii      extraction fused with the level-1 scatter        (reads ASCII, writes records)
//   split 2    k_split_from_records    level-2 scatter inside every level-1 bucket      (reads + writes records)
//   filter     k_filter                per leaf bucket: keep records whose key repeats  (reads records)
//   sort       lsd_sort                candidates by (k-mer, part)                      (small)
//   group      k_group_*               key runs -> groups CSR, internal-repeat detection (small)
//   order      k_group_first_window, k_last_single   ranks for the order-exact host replay
//   edges      k_edge_pairs + lsd_sort + compaction  unique part-pair conflict edges
//
// If grouping finds a part with an internal repeat (and internal_repeats == 0) the part is flagged and the
// pipeline runs once more without it -- the reference drops such parts before they touch the table
// (hgraph.py:50-69), so their records must not be seen by any group.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nrp_b200.h"
#include "common.cuh"
#include "scan.cuh"
#include "lsd_sort.cuh"
#include "partition.cuh"
#include "extract.cuh"
#include "filter_warp.cuh"
#include "stages.cuh"
#include "uniquify.cuh"
#include "generic.cuh"

using namespace nrp;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(e_ == cudaErrorMemoryAllocation ? NRP_E_OOM : NRP_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                         \
  } while (0)
#define TRY(call)               \
  do {                          \
    int r_ = (call);            \
    if (r_ != NRP_OK) return r_; \
  } while (0)

struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return NRP_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NRP_E_OOM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); }
    cap = want;
    return NRP_OK;
  }
  // grow to at least `bytes`, keeping the first `keep` bytes
  int grow_keep(size_t bytes, size_t keep) {
    if (bytes <= cap) return NRP_OK;
    void* fresh = nullptr;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&fresh, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NRP_E_OOM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); }
    if (ptr && keep) cudaMemcpy(fresh, ptr, keep, cudaMemcpyDeviceToDevice);
    if (ptr) cudaFree(ptr);
    ptr = fresh; cap = want;
    return NRP_OK;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

// CNT_EDGES = 8 and CNT_SCATTER_OVERFLOW = 9 live outside the block [0, CNT_N) that every bulk stage clears
// page-locked host buffer that grows on demand
struct HostBuf {
  void* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return NRP_OK;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaHostAlloc(&ptr, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(NRP_E_OOM, "cudaHostAlloc(%zu bytes) failed: %s", want, cudaGetErrorString(e)); }
    cap = want;
    return NRP_OK;
  }
  void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

enum { CNT_SCATTER_OVERFLOW = 9, CNT_SENT = 10, CNT_NSORTED = 11, CNT_GROUP = 12 /* 5 words */, CNT_FLAG_OVERFLOW = 18, CNT_WORDS = 20 };
enum { CNT_CAND = 0, CNT_REPEAT = 1, CNT_FLAGGED = 2, CNT_WORK = 3, CNT_OVERSIZED = 4, CNT_UNSORTED = 5, CNT_OVERFLOW = 6, CNT_RECORDS = 7, CNT_N = 8, CNT_EDGES = 8, CNT_TOTAL = 12 };

constexpr int kFilterThreads = 1024;
constexpr int kFilterItems = 8;
constexpr int kFilterCap = kFilterThreads * kFilterItems;   // 8192 records per leaf bucket in shared memory (1 CTA/SM for 64-bit keys)
constexpr int kFilterItemsSmall = 4;
constexpr int kFilterCapSmall = 4096;   // half-size table: two 1024-thread CTAs per SM
constexpr bool kFilterStaged = false;   // cp.async staging of the next leaf measured slower than direct loads (profiles/README.md)

template <typename KeyT> constexpr size_t filter_smem_large() { return (sizeof(KeyT) + 1) * 2 * (size_t)kFilterCap; }
template <typename KeyT> constexpr size_t filter_smem_small() {
  return (sizeof(KeyT) + 1) * 2 * (size_t)(kFilterThreads * kFilterItemsSmall) + (kFilterStaged ? (sizeof(KeyT) + 4) * (size_t)kFilterCapSmall : 0);
}
constexpr int64_t kDefaultLeafTarget = 2750;                // records per leaf bucket aimed for (two 512-thread filter CTAs per SM, capacity 4096)
constexpr int64_t kLargeLeafTarget = 6400;                  // when the bits run out: one 1024-thread filter CTA per SM, capacity 8192
constexpr int kMaxLeafBits = 15;                            // exact path: 2^15 u32 counters = 128 KB of shared memory
constexpr int kMaxDirectBits = 2 * kMaxLevelBits;           // direct path: two levels of <= 9 bits

}  // namespace

constexpr double kWarpSuspects = 56.0, kWarpSuspectsWide = 150.0;     // expected suspects per leaf the warp-private filter takes (capacity 96 / 192)
constexpr int64_t kWarpLeafTarget = 800;     // records per leaf the warp-private filter is planned for (filter_warp.cuh)
struct nrp_ctx {
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;
  bool configured[2] = {false, false};

  // options
  int64_t opt_filter_cap = kFilterCap;
  int64_t opt_leaf_target = kDefaultLeafTarget;
  int64_t opt_max_level_bits = kMaxLevelBits;
  int64_t opt_max_pairs = (int64_t)1 << 31;
  int64_t opt_no_window_carry = 0;
  int64_t opt_no_bloom = 0;
  int64_t opt_pass_records = 210000000;   // exact path: more records than this per device are processed in several hash-slice passes
  int64_t opt_persistent_split = 0;       // extraction kernel: persistent CTAs with register prefetch (1) or one CTA per byte tile (0)
  int64_t opt_r1_bias = 0;                // added to the number of level-1 bits (tuning)
  int64_t opt_nvlink_digits = 128;        // direct exchange: most (owner, bucket) digits per tile the senders scatter over NVLink
  int64_t opt_filter_tiny = 0;            // packed leaves that average <= 1400 records: four 256-thread filter CTAs per SM (pair with leaf_target <= 1400)
  int64_t opt_coop_group = 1;             // grouping stage as one cooperative kernel: 1 = up to 2^20 candidates (measured: no faster than
                                          // the kernel sequence at 8e5 candidates, slower at 2e7), 2 = always, 0 = never
  int64_t opt_group_walk = 1;             // grouping with two grid barriers (run heads walk their runs) instead of nine; falls back on long runs
  int64_t opt_packed = 1;                 // one 64-bit word per record on the direct path whenever key and value fit
  int64_t opt_direct = 1;                 // fixed-region partition without a histogram pass (falls back to the exact path on overflow)
  int64_t opt_filter_wide = 1;            // packed leaves of <= 4096 records: three 256-thread CTAs per SM with 16 records per thread (half the
                                          // per-leaf fixed work per record: config 2 filter 0.437 -> 0.381 ms) when few suspects are expected;
                                          // 0 = two 512-thread CTAs
  int64_t opt_filter_warp = 1;            // packed leaves that average <= 800 records: the warp-private filter (filter_warp.cuh); the level planner
                                          // then aims at such leaves
  int64_t opt_warp_leaf_target = kWarpLeafTarget;     // mean records per leaf the planner aims at for the warp-private filter (halved while too many suspects
                                          // are expected); <= 400 selects its small-map variant (4.2 KB per warp).  Smaller is not faster:
                                          // config 2 with 2^18 leaves of 320 records, filter 0.30 -> 0.39 ms (r02p)
  int64_t opt_early_fetch = 0;            // single-GPU build: results that are final after the grouping stage are copied to the page-locked
                                          // result buffer beside the rest of the build (nrp_result_host finds them there).  For callers
                                          // that fetch: config 2 end to end 3.35 -> 3.28 ms, config 3 10.1 -> 8.8 ms; the build itself
                                          // ends later (its last count read-back queues behind the copy on PCIe), hence off by default
  int64_t opt_overlap_tail = 1;           // single-GPU build: the last-unshared-window pass runs on a second stream beside the edge expansion
  int64_t opt_fold_flags = 1;             // part-aligned kernel: palindrome test and (K == k) background probe inside the extraction instead
                                          // of a pass of their own over the parts
  int64_t opt_uniform_kernel = 1;         // parts of one length: the part-aligned extraction kernel (extract.cuh); 0 = the byte-position kernel

  // background
  DevBuf bg_table; int bg_K = 0; int bg_log_slots = 0; int bg_key_bytes = 0; int64_t bg_n = 0;
  DevBuf bg_bits; int bg_log_bits = 0;      // bit map by the top bits of the bijective hash (folded probe, extract.cuh)
  bool fold_pal = false, fold_bg = false;   // this job's extraction kernel computes these flags itself
  // background build on the device: staged sequences, extracted / sorted / distinct keys
  DevBuf bgb_ascii, bgb_off, bgb_tile_first, bgb_keysA, bgb_keysB, bgb_head, bgb_head_scan;
  uint64_t* bgb_result = nullptr; int64_t bgb_n = 0;

  // device-side uniquify (nrp_uniquify): the caller's parts, the surviving parts in processing order, their ids / sources / offsets
  DevBuf uq_in, uq_off, uq_out, uq_rep, uq_run_first, uq_source, uq_ids, uq_lens, uq_newoff, uq_status;
  int64_t uq_n = 0, uq_unique = -1, uq_total = 0, uq_uniform = 0;

  // general path (generic.cuh): records keyed by a seeded 64-bit hash of the canonical window string
  bool generic = false; uint64_t gen_seed = 0;
  DevBuf gen_keys, gen_vals, gen_status;

  // loaded parts
  DevBuf ascii, off, part_id, tile_first;
  const int64_t* h_off = nullptr;            // host copy of the offsets (lives in h_upload until the next batch)
  int64_t n_parts = 0, total_len = 0, n_tiles = 0, max_len = 0;
  int uniform_len = 0;
  bool have_parts = false, have_ids = false;
  int64_t staged_lo = 0, staged_hi = 0;      // parts whose bytes are on this device (shard-only staging: a sub-range)
  PartHomes homes{};                         // shard-only staging: the ranks' part buffers (CUDA IPC) and shard bounds
  bool flag_lists_applied = false;           // nrp_flags_repeat_apply ran since the last nrp_shard_finish
  bool flags_set = false;                    // the flag pass of the current job ran a kernel (some flag byte may be non-zero)
  int coop_blocks_per_sm[2] = {0, 0};
  bool group_counts_coop = false;
  bool group_walk_gave_up = false;           // the walk variant met a run longer than kGroupWalkMax: this job repeats the stage without it
  bool cand_compacted = false;               // the grouping stage left the candidate list without the records of flagged parts
  // streamed staging: the ASCII copy is in flight on copy_stream in pieces (contiguous part ranges); pending_chunks > 0 until a
  // build (or any other consumer) has waited for them
  struct Chunk { int64_t part_lo, part_hi, tile_lo, n_tiles; cudaEvent_t landed; };
  std::vector<Chunk> chunks;
  cudaStream_t copy_stream = nullptr;
  int pending_chunks = 0;
  cudaStream_t aux_stream = nullptr;         // second compute stream of the tail (last unshared windows beside the edge expansion)
  cudaEvent_t aux_begin = nullptr, aux_end = nullptr;
  uint32_t part_base = 0;

  // pipeline buffers
  DevBuf flags, hist, leaf_off, cursor1, cursor2, seg_off, tile_prefix, tile_desc, region_ends, counters, scan_scratch, lsd_scratch;
  DevBuf keysA, valsA, keysB, valsB, cand_keys_buf, cand_vals_buf;
  DevBuf head, head_scan, group_start, valid, valid_scan, mem_cnt, mem_scan, pair_cnt, pair_scan;
  DevBuf g_member_end;
  DevBuf g_keys, g_indptr, g_pair_off, g_members, g_first_part, g_first_window, last_single;
  DevBuf e_codesA, e_codesB, e_head, e_head_scan, e_u, e_v, e_u_alt, e_v_alt, shared_keysA, shared_keysB;
  unsigned long long* h_counters = nullptr;   // pinned
  HostBuf h_upload, h_results;                // pinned staging: offsets + ids on the way in, all results on the way out
  // early fetch (single-GPU build): everything that is final after the grouping stage (members, group offsets, first windows, keys,
  // flags) starts its way to h_results on fetch_stream while the edge expansion and the last-unshared-window pass still run;
  // nrp_result_host then copies only what came later.  early_* describe what is (being) copied and for which sizes.
  cudaStream_t fetch_stream = nullptr;
  cudaEvent_t fetch_ready = nullptr, fetch_done = nullptr;
  bool early_valid = false;
  int64_t early_n = 0, early_ng = 0, early_nm = 0;

  // shard window (whole batch unless nrp_shard_extract narrowed it)
  int64_t shard_lo = 0, shard_hi = 0, shard_tile_lo = 0, shard_tiles = 0, shard_m_max = 0;
  int n_owners = 0;
  int j_bits = 0;                // > 0: record values are (part << j_bits) | window
  int vbits = 32;                // bits of a record value
  DevBuf fb_keys, fb_vals;       // unpacked copy of a packed source for the exact-path fallback
  // receive buffers of the fused exchange (exported with CUDA IPC) and the peers' buffers opened here
  DevBuf recv_keys, recv_vals;
  // direct fused exchange: fixed regions (level-1 bucket x sender) + the end position of every region
  DevBuf recv_keys_d, recv_vals_d, recv_ends;
  PeerBuffers peers_d{};
  uint32_t* peer_ends[kMaxPeers] = {nullptr};
  int sh_world = 0, sh_r1 = 0, sh_bits = 0, sh_key_bytes = 0, n_peers_d = 0, sh_owner_bits = 0;
  bool sh_packed = false, sh_no_window = false;
  uint32_t sh_cap1s = 0;
  int64_t sh_windows_total = 0, sh_m_owner = 0, sh_sent = 0;
  int64_t recv_capacity = 0;
  PeerBuffers peers{};
  int n_peers = 0, my_rank = 0;
  std::vector<std::vector<unsigned char>> opened_handles;
  std::vector<void*> opened_ptrs;
  bool external_stream = false;

  // candidate records between the stages (type-erased: uint32_t* or uint64_t* keys)
  void* cand_keys = nullptr; uint32_t* cand_vals = nullptr;     // candidates / sorted candidates
  void* spare_keys = nullptr; uint32_t* spare_vals = nullptr;   // scratch of the same size
  int64_t n_cand = 0, n_sorted = 0, n_records = 0, n_repeat_records = 0;
  bool cand_unsorted = false;    // the filter could not order some leaf's candidates: the global sort is needed
  int64_t n_groups = 0, n_members = 0, n_pairs = 0;

  // result of the last build
  bool have_result = false;
  int k = 0, key_bytes = 0;
  int64_t counts[NRP_N_COUNTS] = {0};
  double timings[NRP_N_TIMINGS] = {0};
};

namespace {

Parts parts_view(const nrp_ctx* c) {
  Parts P;
  P.ascii = c->ascii.as<uint8_t>();
  P.off = c->off.as<int64_t>();
  P.tile_first = c->tile_first.as<uint32_t>();
  P.n_parts = c->n_parts;
  P.total_len = c->total_len;
  P.uniform_len = c->uniform_len;
  P.part_lo = c->shard_lo;
  P.part_hi = c->shard_hi;
  P.tile_lo = c->shard_tile_lo;
  return P;
}

__global__ void k_uniform_offsets(int64_t* off, int64_t n_parts, int64_t len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_parts) off[i] = i * len;
}

// byte offset of part i on the host (uniform batches keep no host copy of the offsets)
inline int64_t part_offset(const nrp_ctx* c, int64_t i) { return c->uniform_len > 0 ? i * (int64_t)c->uniform_len : c->h_off[i]; }

__global__ void k_init_cursors(const uint32_t* leaf_off, int r1, int r2, uint32_t* cursor1, uint32_t* cursor2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (1 << (r1 + r2))) cursor2[i] = leaf_off[i];
  if (i < (1 << r1)) cursor1[i] = leaf_off[(size_t)i << r2];
}

__global__ void k_gather_first_part(const int64_t* indptr, const uint32_t* members, int64_t n_groups, uint32_t* first_part) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_groups) first_part[g] = members[indptr[g]];
}

inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
// persistent CTAs of the extraction kernel: two per SM (registers and shared memory allow no more)
inline unsigned split_grid(const nrp_ctx* c) {
  const int64_t cap = c->opt_persistent_split ? (int64_t)c->n_sms * 2 : (int64_t)0x7fffffff;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(c->shard_tiles, cap));
}

struct StageTimer {
  nrp_ctx* c;
  size_t next = 0;
  std::vector<std::pair<int, size_t>> marks;   // (timing slot, event index of the stage's start)
  explicit StageTimer(nrp_ctx* ctx) : c(ctx) {}
  size_t stamp() {
    if (next >= c->events.size()) { cudaEvent_t e; cudaEventCreate(&e); c->events.push_back(e); }
    cudaEventRecord(c->events[next], c->stream);
    return next++;
  }
  void begin(int slot) { marks.push_back({slot, stamp()}); }
  void finish(double* out) {
    size_t end = stamp();
    cudaEventSynchronize(c->events[end]);
    for (size_t i = 0; i < marks.size(); ++i) {
      size_t stop = i + 1 < marks.size() ? marks[i + 1].second : end;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->events[marks[i].second], c->events[stop]);
      out[marks[i].first] += ms;
    }
  }
};

template <typename KeyT>
int configure_kernels(nrp_ctx* c) {
  bool& done = c->configured[sizeof(KeyT) == 4 ? 0 : 1];
  if (done) return NRP_OK;
  const int split_smem = (int)(sizeof(SplitSmem<KeyT>) + sizeof(TileStage));
  CU(cudaFuncSetAttribute((k_split_from_ascii<KeyT, 0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, split_smem));
  CU(cudaFuncSetAttribute((k_split_from_ascii<KeyT, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, split_smem));
  CU(cudaFuncSetAttribute((k_split_from_ascii<KeyT, 1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, split_smem));
  CU(cudaFuncSetAttribute((k_split_from_ascii<KeyT, 2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, split_smem));
  CU(cudaFuncSetAttribute((k_split_from_records<KeyT, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem<KeyT>)));
  CU(cudaFuncSetAttribute((k_split_from_records<KeyT, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem<KeyT>)));
  CU(cudaFuncSetAttribute((k_filter_bloom<KeyT, 512, 8, 2048, 2, 512, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_bloom_smem<(int)sizeof(KeyT), false, 4096, 2048, 512>()));
  CU(cudaFuncSetAttribute((k_filter_bloom<KeyT, kFilterThreads, 8, 4096, 1, 1024, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_bloom_smem<(int)sizeof(KeyT), false, 8192, 4096, 1024>()));
  CU(cudaFuncSetAttribute((k_filter_bloom<KeyT, 512, 8, 2048, 2, 512, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_bloom_smem<8, true, 4096, 2048, 512>()));
  CU(cudaFuncSetAttribute((k_filter_bloom<KeyT, 256, 8, 1024, 4, 256, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_bloom_smem<8, true, 2048, 1024, 256>()));
  CU(cudaFuncSetAttribute((k_filter_bloom<KeyT, 256, 16, 2048, 3, 256, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_bloom_smem<8, true, 4096, 2048, 256>()));
  CU(cudaFuncSetAttribute((k_filter_bloom<KeyT, kFilterThreads, 8, 4096, 1, 1024, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_bloom_smem<8, true, 8192, 4096, 1024>()));
  CU(cudaFuncSetAttribute((k_filter_warp<KeyT, 512, 192, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFwWarps * sizeof(FwWarp<512, 192>))));
  CU(cudaFuncSetAttribute((k_filter_warp<KeyT, 512, 192, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFwWarps * sizeof(FwWarp<512, 192>))));
  CU(cudaFuncSetAttribute((k_filter_warp<KeyT, 512, 96, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFwWarps * sizeof(FwWarp<512, 96>))));
  CU(cudaFuncSetAttribute((k_filter_warp<KeyT, 512, 96, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFwWarps * sizeof(FwWarp<512, 96>))));
  CU(cudaFuncSetAttribute((k_filter_warp<KeyT, 256, 64, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFwWarps * sizeof(FwWarp<256, 64>))));
  CU(cudaFuncSetAttribute((k_filter_warp<KeyT, 256, 64, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kFwWarps * sizeof(FwWarp<256, 64>))));
  CU(cudaFuncSetAttribute((k_split_uniform<false, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UniSmem)));
  CU(cudaFuncSetAttribute((k_split_uniform<true, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UniSmem)));
  CU(cudaFuncSetAttribute((k_split_uniform<false, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UniSmem)));
  CU(cudaFuncSetAttribute((k_split_uniform<true, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UniSmem)));
  CU(cudaFuncSetAttribute((k_split_uniform<false, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UniSmem)));
  CU(cudaFuncSetAttribute((k_split_uniform<true, 0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(UniSmem)));
  const int packed_smem = (int)(sizeof(SplitSmem<uint64_t, true>) + sizeof(TileStage));
  CU(cudaFuncSetAttribute((k_split_from_ascii<uint64_t, 0, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, packed_smem));
  CU(cudaFuncSetAttribute((k_split_from_ascii<uint64_t, 2, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, packed_smem));
  CU(cudaFuncSetAttribute((k_split_from_records<uint64_t, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)sizeof(SplitSmem<uint64_t, true>)));
  CU(cudaFuncSetAttribute((k_filter<KeyT, kFilterThreads, kFilterItems, kFilterCap, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_smem_large<KeyT>()));
  CU(cudaFuncSetAttribute((k_filter<KeyT, kFilterThreads, kFilterItemsSmall, kFilterCapSmall, kFilterStaged>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)filter_smem_small<KeyT>()));
  CU(cudaFuncSetAttribute(k_hist_from_ascii<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << kMaxLeafBits));
  CU(cudaFuncSetAttribute(k_hist_from_ascii<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << kMaxLeafBits));
  CU(cudaFuncSetAttribute(k_hist_from_tiles<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << kMaxLeafBits));
  done = true;
  return NRP_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int nrp_version(void) { return 100; }
int64_t nrp_launches_total(void) { return launch_total(); }
const char* nrp_last_error(void) { return g_last_error.c_str(); }

int nrp_ctx_create(int device, nrp_ctx** out) {
  if (!out) return fail(NRP_E_ARG, "out is NULL");
  int n_dev = 0;
  CU(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) return fail(NRP_E_ARG, "device %d out of range (%d devices)", device, n_dev);
  CU(cudaSetDevice(device));
  nrp_ctx* c = new nrp_ctx();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->n_sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaHostAlloc(&c->h_counters, sizeof(unsigned long long) * 64, cudaHostAllocDefault));
  int r = c->counters.ensure(sizeof(unsigned long long) * CNT_WORDS);
  if (r != NRP_OK) return r;
  *out = c;
  return NRP_OK;
}

int nrp_ctx_destroy(nrp_ctx* c) {
  if (!c) return NRP_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf* all[] = {&c->bgb_ascii, &c->bgb_off, &c->bgb_tile_first, &c->bgb_keysA, &c->bgb_keysB, &c->bgb_head, &c->bgb_head_scan, &c->bg_table, &c->ascii, &c->off, &c->part_id, &c->tile_first, &c->flags, &c->hist, &c->leaf_off, &c->cursor1,
                   &c->cursor2, &c->seg_off, &c->tile_prefix, &c->tile_desc, &c->region_ends, &c->recv_ends, &c->recv_keys_d, &c->recv_vals_d, &c->fb_keys, &c->fb_vals, &c->counters, &c->scan_scratch, &c->lsd_scratch, &c->keysA, &c->valsA, &c->cand_keys_buf, &c->cand_vals_buf,
                   &c->keysB, &c->valsB, &c->head, &c->head_scan, &c->group_start, &c->valid, &c->valid_scan, &c->mem_cnt,
                   &c->mem_scan, &c->pair_cnt, &c->pair_scan, &c->g_member_end, &c->g_keys, &c->g_indptr, &c->g_pair_off, &c->g_members,
                   &c->g_first_part, &c->g_first_window, &c->last_single, &c->e_codesA, &c->e_codesB, &c->e_head, &c->e_head_scan,
                   &c->e_u, &c->e_v, &c->e_u_alt, &c->e_v_alt, &c->shared_keysA, &c->shared_keysB, &c->recv_keys, &c->recv_vals};
  for (void* p : c->opened_ptrs) cudaIpcCloseMemHandle(p);
  for (DevBuf* b : all) b->release();
  for (cudaEvent_t e : c->events) cudaEventDestroy(e);
  for (auto& ch : c->chunks) cudaEventDestroy(ch.landed);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->fetch_stream) { cudaStreamSynchronize(c->fetch_stream); cudaStreamDestroy(c->fetch_stream); cudaEventDestroy(c->fetch_ready); cudaEventDestroy(c->fetch_done); }
  if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); cudaEventDestroy(c->aux_begin); cudaEventDestroy(c->aux_end); }
  if (c->h_counters) cudaFreeHost(c->h_counters);
  c->h_upload.release(); c->h_results.release();
  if (!c->external_stream) cudaStreamDestroy(c->stream);
  delete c;
  return NRP_OK;
}

int nrp_ctx_device(nrp_ctx* c) { return c ? c->device : -1; }

int nrp_ctx_set_option(nrp_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return fail(NRP_E_ARG, "null argument");
  std::string n(name);
  if (n == "filter_cap") c->opt_filter_cap = (value <= 0 || value > kFilterCap) ? kFilterCap : value;
  else if (n == "leaf_target") c->opt_leaf_target = value <= 0 ? kDefaultLeafTarget : value;
  else if (n == "max_level_bits") c->opt_max_level_bits = (value <= 0 || value > kMaxLevelBits) ? kMaxLevelBits : value;
  else if (n == "max_pairs") c->opt_max_pairs = value <= 0 ? ((int64_t)1 << 31) : value;
  else if (n == "no_window_carry") c->opt_no_window_carry = value > 0 ? 1 : 0;
  else if (n == "no_bloom") c->opt_no_bloom = value > 0 ? 1 : 0;
  else if (n == "pass_records") c->opt_pass_records = value <= 0 ? 210000000 : value;
  else if (n == "persistent_split") c->opt_persistent_split = value > 0 ? 1 : 0;
  else if (n == "r1_bias") c->opt_r1_bias = value;
  else if (n == "nvlink_digits") c->opt_nvlink_digits = value <= 0 ? 128 : value;
  else if (n == "filter_tiny") c->opt_filter_tiny = value > 0 ? 1 : 0;
  else if (n == "coop_group") c->opt_coop_group = value < 0 ? 1 : std::min<int64_t>(value, 2);
  else if (n == "group_walk") c->opt_group_walk = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "packed") c->opt_packed = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "direct") c->opt_direct = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "filter_wide") c->opt_filter_wide = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "uniform_kernel") c->opt_uniform_kernel = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "fold_flags") c->opt_fold_flags = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "overlap_tail") c->opt_overlap_tail = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "early_fetch") c->opt_early_fetch = value > 0 ? 1 : 0;
  else if (n == "filter_warp") c->opt_filter_warp = value < 0 ? 1 : (value > 0 ? 1 : 0);
  else if (n == "warp_leaf_target") c->opt_warp_leaf_target = value <= 0 ? kWarpLeafTarget : std::min<int64_t>(value, kWarpLeafTarget);
  else return fail(NRP_E_ARG, "unknown option '%s'", name);
  return NRP_OK;
}

int nrp_bg_clear(nrp_ctx* c) {
  if (!c) return fail(NRP_E_ARG, "null context");
  c->bg_n = 0; c->bg_K = 0;
  return NRP_OK;
}

// probe table (open addressing, >= 2n slots) from n canonical K-mers in device memory
static int bg_install_device(nrp_ctx* c, const uint64_t* dev_keys, int64_t n, int K) {
  if (n == 0) return nrp_bg_clear(c);
  int log_slots = 6;
  while (((int64_t)1 << log_slots) < 2 * n) ++log_slots;
  const int key_bytes = K <= 16 ? 4 : 8;
  const size_t table_bytes = (size_t)key_bytes << log_slots;
  TRY(c->bg_table.ensure(table_bytes));
  CU(cudaMemsetAsync(c->bg_table.ptr, 0xff, table_bytes, c->stream));
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, c->n_sms * 8);
  if (key_bytes == 4) NRP_LAUNCH((k_bg_insert<uint32_t><<<grid, 256, 0, c->stream>>>(dev_keys, n, c->bg_table.as<uint32_t>(), log_slots)));
  else NRP_LAUNCH((k_bg_insert<uint64_t><<<grid, 256, 0, c->stream>>>(dev_keys, n, c->bg_table.as<uint64_t>(), log_slots)));
  // bit map for the folded probe: >= 16 bits per key, at most 2K (all the hash has) and 2^30 bits
  int log_bits = 10;
  while (log_bits < 30 && ((int64_t)1 << log_bits) < 16 * n) ++log_bits;
  log_bits = std::min(log_bits, 2 * K);
  const size_t bit_bytes = ((((size_t)1 << log_bits) + 31) / 32) * 4;
  TRY(c->bg_bits.ensure(bit_bytes));
  CU(cudaMemsetAsync(c->bg_bits.ptr, 0, bit_bytes, c->stream));
  NRP_LAUNCH((k_bg_bitmap<<<grid, 256, 0, c->stream>>>(dev_keys, n, make_bijhash(K), log_bits, c->bg_bits.as<uint32_t>())));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->bg_n = n; c->bg_K = K; c->bg_log_slots = log_slots; c->bg_key_bytes = key_bytes; c->bg_log_bits = log_bits;
  return NRP_OK;
}

int nrp_bg_set(nrp_ctx* c, const uint64_t* canon_kmers, int64_t n, int K) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (n < 0 || (n > 0 && !canon_kmers)) return fail(NRP_E_ARG, "bad background array");
  if (K < 1 || K > 32) return fail(NRP_E_UNSUPPORTED, "background K=%d outside [1, 32]", K);
  CU(cudaSetDevice(c->device));
  if (n == 0) return nrp_bg_clear(c);
  DevBuf staging;
  TRY(staging.ensure(sizeof(uint64_t) * (size_t)n));
  CU(cudaMemcpyAsync(staging.ptr, canon_kmers, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  int r = bg_install_device(c, staging.as<uint64_t>(), n, K);
  staging.release();
  return r;
}

int nrp_bg_build(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, int64_t n_seqs, int K, int install, int64_t* n_unique) {
  if (!c || !offsets || !n_unique) return fail(NRP_E_ARG, "null argument");
  if (K < 1 || K > 32) return fail(NRP_E_UNSUPPORTED, "background K=%d outside [1, 32]", K);
  if (n_seqs < 0 || offsets[0] != 0 || (n_seqs > 0 && !ascii && offsets[n_seqs] > 0)) return fail(NRP_E_ARG, "bad sequence buffers");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->bgb_n = 0; c->bgb_result = nullptr;
  const int64_t total = offsets[n_seqs];
  int64_t m = 0;
  int uniform = n_seqs > 0 ? (int)std::min<int64_t>(offsets[1] - offsets[0], (int64_t)1 << 30) : 0;
  for (int64_t i = 0; i < n_seqs; ++i) {
    const int64_t len = offsets[i + 1] - offsets[i];
    if (len < 0) return fail(NRP_E_ARG, "offsets must be non-decreasing (sequence %lld)", (long long)i);
    if (len != uniform) uniform = 0;
    if (len >= K) m += len - K + 1;
  }
  if (uniform >= (1 << 30)) uniform = 0;
  if (m >= 0xffffe000ll) return fail(NRP_E_UNSUPPORTED, "%lld windows exceed the 32-bit index of one device", (long long)m);
  *n_unique = 0;
  if (m == 0) { if (install) return nrp_bg_clear(c); return NRP_OK; }
  const int64_t n_tiles = (total + kTileBytes - 1) / kTileBytes;
  const size_t padded = (size_t)n_tiles * kTileBytes + 256;
  TRY(c->bgb_ascii.ensure(padded));
  TRY(c->bgb_off.ensure(sizeof(int64_t) * (size_t)(n_seqs + 1)));
  TRY(c->bgb_tile_first.ensure(sizeof(uint32_t) * (size_t)(n_tiles + 2)));
  TRY(c->bgb_keysA.ensure(8 * (size_t)(m + 16))); TRY(c->bgb_keysB.ensure(8 * (size_t)(m + 16)));
  TRY(c->bgb_head.ensure(4 * (size_t)(m + 2))); TRY(c->bgb_head_scan.ensure(4 * (size_t)(m + 2)));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems(m + 2)));
  TRY(c->lsd_scratch.ensure(sizeof(uint32_t) * (size_t)lsd_scratch_elems(m)));
  CU(cudaMemsetAsync(c->bgb_ascii.as<uint8_t>() + total, 0, padded - (size_t)total, st));
  CU(cudaMemcpyAsync(c->bgb_ascii.ptr, ascii, (size_t)total, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(c->bgb_off.ptr, offsets, sizeof(int64_t) * (size_t)(n_seqs + 1), cudaMemcpyHostToDevice, st));
  if (uniform == 0)
    NRP_LAUNCH((k_tile_first_part<<<grid_for(n_tiles + 1, 256), 256, 0, st>>>(c->bgb_off.as<int64_t>(), n_seqs, total, n_tiles, c->bgb_tile_first.as<uint32_t>())));
  Parts P;
  P.ascii = c->bgb_ascii.as<uint8_t>(); P.off = c->bgb_off.as<int64_t>(); P.tile_first = c->bgb_tile_first.as<uint32_t>();
  P.n_parts = n_seqs; P.total_len = total; P.uniform_len = uniform; P.part_lo = 0; P.part_hi = n_seqs; P.tile_lo = 0;
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  CU(cudaMemsetAsync(cnt + CNT_WORK, 0, 8, st));
  NRP_LAUNCH((k_extract_canon<<<(unsigned)n_tiles, kTileThreads, 0, st>>>(P, K, c->bgb_keysA.as<uint64_t>(), cnt + CNT_WORK)));
  uint64_t* sorted = nullptr;
  LsdPlan plan; plan.n_passes = 0;
  lsd_plan_add(plan, 0, 0, 2 * K);
  lsd_sort<uint64_t, false>(c->bgb_keysA.as<uint64_t>(), nullptr, c->bgb_keysB.as<uint64_t>(), nullptr, m, plan, c->lsd_scratch.as<uint32_t>(), st,
                            &sorted, nullptr);
  uint64_t* spare = sorted == c->bgb_keysA.as<uint64_t>() ? c->bgb_keysB.as<uint64_t>() : c->bgb_keysA.as<uint64_t>();
  NRP_LAUNCH((k_unique_heads<<<grid_for(m, 256), 256, 0, st>>>(sorted, m, c->bgb_head.as<uint32_t>())));
  exclusive_scan<uint32_t, uint32_t>(c->bgb_head.as<uint32_t>(), c->bgb_head_scan.as<uint32_t>(), m, c->scan_scratch.as<uint32_t>(), st);
  NRP_LAUNCH((k_unique_compact<<<grid_for(m, 256), 256, 0, st>>>(sorted, c->bgb_head.as<uint32_t>(), c->bgb_head_scan.as<uint32_t>(), m, spare)));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_counters + 48, c->bgb_head_scan.as<uint32_t>() + m, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_counters + 49, cnt + CNT_WORK, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if ((int64_t)c->h_counters[49] != m) return fail(NRP_E_CUDA, "background extraction produced %lld of %lld windows", (long long)c->h_counters[49], (long long)m);
  c->bgb_n = (int64_t)(uint32_t)c->h_counters[48];
  c->bgb_result = spare;
  *n_unique = c->bgb_n;
  if (install) return bg_install_device(c, spare, c->bgb_n, K);
  return NRP_OK;
}

int nrp_bg_build_fetch(nrp_ctx* c, uint64_t* keys_out) {
  if (!c || (!keys_out && c->bgb_n > 0)) return fail(NRP_E_ARG, "null argument");
  if (c->bgb_n > 0 && !c->bgb_result) return fail(NRP_E_STATE, "no finished background build");
  CU(cudaSetDevice(c->device));
  if (c->bgb_n > 0) CU(cudaMemcpyAsync(keys_out, c->bgb_result, 8 * (size_t)c->bgb_n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

// ---- device-side uniquify (uniquify.cuh) ----------------------------------------------------------------------------------------
int nrp_uniquify(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, int part_len, int64_t n_parts, int64_t out[4]) {
  if (!c || !out) return fail(NRP_E_ARG, "null argument");
  if (n_parts < 0 || (!offsets && part_len < 1 && n_parts > 0)) return fail(NRP_E_ARG, "offsets or a positive part_len is required");
  if (n_parts >= 0xfffffff0ll) return fail(NRP_E_UNSUPPORTED, "%lld parts exceed the 32-bit part index", (long long)n_parts);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->uq_unique = -1;
  int64_t total = 0, min_len = 0, max_len = 0;
  if (offsets) {
    if (offsets[0] != 0) return fail(NRP_E_ARG, "offsets[0] must be 0");
    min_len = n_parts > 0 ? INT64_MAX : 0;
    for (int64_t i = 0; i < n_parts; ++i) {
      const int64_t len = offsets[i + 1] - offsets[i];
      min_len = std::min(min_len, len); max_len = std::max(max_len, len);
    }
    if (min_len < 0) return fail(NRP_E_ARG, "offsets must be non-decreasing");
    total = offsets[n_parts];
  } else {
    total = n_parts * (int64_t)part_len; min_len = max_len = n_parts > 0 ? part_len : 0;
  }
  if (total > 0 && !ascii) return fail(NRP_E_ARG, "null part buffer");
  if (max_len >= ((int64_t)1 << 31)) return fail(NRP_E_UNSUPPORTED, "parts of 2^31 bytes or more");
  const int64_t uniform = (n_parts > 0 && min_len == max_len) ? max_len : 0;
  out[0] = 0; out[1] = 0; out[2] = 0; out[3] = max_len;
  c->uq_n = n_parts; c->uq_uniform = uniform; c->uq_total = 0;
  if (n_parts == 0) { c->uq_unique = 0; return NRP_OK; }
  const size_t n1 = (size_t)n_parts + 2;
  TRY(c->uq_in.ensure((size_t)total + 64)); TRY(c->uq_out.ensure((size_t)total + 64));
  TRY(c->uq_off.ensure(8 * n1)); TRY(c->uq_newoff.ensure(8 * n1));
  TRY(c->uq_rep.ensure(4 * n1)); TRY(c->uq_run_first.ensure(4 * n1)); TRY(c->uq_source.ensure(4 * n1));
  TRY(c->uq_ids.ensure(4 * n1)); TRY(c->uq_lens.ensure(4 * n1)); TRY(c->uq_status.ensure(16));
  TRY(c->keysA.ensure(8 * n1)); TRY(c->keysB.ensure(8 * n1)); TRY(c->valsA.ensure(4 * n1)); TRY(c->valsB.ensure(4 * n1));
  TRY(c->head.ensure(4 * n1)); TRY(c->head_scan.ensure(4 * n1)); TRY(c->valid.ensure(4 * n1)); TRY(c->valid_scan.ensure(4 * n1));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems((int64_t)n1 + 1)));
  TRY(c->lsd_scratch.ensure(sizeof(uint32_t) * (size_t)lsd_scratch_elems(n_parts)));
  CU(cudaStreamSynchronize(st));
  if (total > 0) CU(cudaMemcpyAsync(c->uq_in.ptr, ascii, (size_t)total, cudaMemcpyHostToDevice, st));
  if (offsets && uniform == 0) CU(cudaMemcpyAsync(c->uq_off.ptr, offsets, 8 * (size_t)(n_parts + 1), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(c->uq_status.ptr, 0, 16, st));
  CU(cudaMemsetAsync(c->uq_rep.ptr, 0, 4 * n1, st));
  const PartSpan S{c->uq_in.as<uint8_t>(), c->uq_off.as<int64_t>(), uniform};
  uint32_t* status = c->uq_status.as<uint32_t>();
  const unsigned warp_grid = (unsigned)std::min<int64_t>((n_parts + 7) / 8, (int64_t)c->n_sms * 16);
  NRP_LAUNCH((k_part_hash<<<warp_grid, 256, 0, st>>>(S, n_parts, c->keysA.as<uint64_t>(), c->valsA.as<uint32_t>(), status)));
  LsdPlan plan; plan.n_passes = 0;
  lsd_plan_add(plan, 0, 0, 64);
  uint64_t* sk = nullptr; uint32_t* sv = nullptr;
  lsd_sort<uint64_t, true>(c->keysA.as<uint64_t>(), c->valsA.as<uint32_t>(), c->keysB.as<uint64_t>(), c->valsB.as<uint32_t>(), n_parts, plan,
                           c->lsd_scratch.as<uint32_t>(), st, &sk, &sv);
  const unsigned grid = grid_for(n_parts, 256);
  NRP_LAUNCH((k_uniq_heads<<<grid, 256, 0, st>>>(S, sk, sv, n_parts, c->head.as<uint32_t>(), status)));
  exclusive_scan<uint32_t, uint32_t>(c->head.as<uint32_t>(), c->head_scan.as<uint32_t>(), n_parts, c->scan_scratch.as<uint32_t>(), st);
  NRP_LAUNCH((k_uniq_runs<<<grid, 256, 0, st>>>(sk, sv, c->head.as<uint32_t>(), c->head_scan.as<uint32_t>(), n_parts, c->uq_run_first.as<uint32_t>(),
                                                c->uq_rep.as<uint32_t>())));
  NRP_LAUNCH((k_uniq_keep<<<grid, 256, 0, st>>>(c->uq_rep.as<uint32_t>(), n_parts, c->valid.as<uint32_t>())));
  exclusive_scan<uint32_t, uint32_t>(c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(), n_parts, c->scan_scratch.as<uint32_t>(), st);
  NRP_LAUNCH((k_uniq_emit<<<grid, 256, 0, st>>>(S, c->uq_rep.as<uint32_t>(), c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(),
                                                c->uq_run_first.as<uint32_t>(), n_parts, c->uq_source.as<uint32_t>(), c->uq_ids.as<uint32_t>(),
                                                c->uq_lens.as<uint32_t>())));
  CU(cudaMemcpyAsync(c->h_counters + 58, c->valid_scan.as<uint32_t>() + n_parts, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_counters + 59, status, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  const int64_t n_unique = (int64_t)(uint32_t)c->h_counters[58];
  const uint32_t bits = (uint32_t)c->h_counters[59];
  int64_t total_unique = n_unique * uniform;
  if (uniform == 0) {
    exclusive_scan<uint32_t, unsigned long long>(c->uq_lens.as<uint32_t>(), c->uq_newoff.as<unsigned long long>(), n_unique,
                                                 c->scan_scratch.as<unsigned long long>(), st);
    CU(cudaMemcpyAsync(c->h_counters + 58, c->uq_newoff.as<unsigned long long>() + n_unique, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    total_unique = (int64_t)c->h_counters[58];
  }
  if (n_unique > 0)
    NRP_LAUNCH((k_uniq_gather<<<(unsigned)std::min<int64_t>((n_unique + 7) / 8, (int64_t)c->n_sms * 16), 256, 0, st>>>(
        S, c->uq_source.as<uint32_t>(), c->uq_newoff.as<unsigned long long>(), uniform, n_unique, c->uq_out.as<uint8_t>())));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  c->uq_unique = n_unique; c->uq_total = total_unique;
  out[0] = n_unique; out[1] = total_unique; out[2] = (int64_t)bits; out[3] = max_len;
  return NRP_OK;
}

int nrp_uniquify_fetch(nrp_ctx* c, uint32_t* ids, uint32_t* source, int64_t* offsets, void** dev_ascii) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (c->uq_unique < 0) return fail(NRP_E_STATE, "no finished nrp_uniquify");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int64_t n = c->uq_unique;
  if (n > 0 && ids) CU(cudaMemcpyAsync(ids, c->uq_ids.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (n > 0 && source) CU(cudaMemcpyAsync(source, c->uq_source.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (offsets) {
    if (c->uq_uniform > 0 || n == 0) { for (int64_t i = 0; i <= n; ++i) offsets[i] = i * c->uq_uniform; }
    else CU(cudaMemcpyAsync(offsets, c->uq_newoff.ptr, 8 * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  if (dev_ascii) *dev_ascii = c->uq_out.ptr;
  return NRP_OK;
}

// uniform_hint > 0: every part has exactly this many bytes and `offsets` is NULL (nothing to validate or to stage: part i
// starts at i * uniform_hint)
// copy_lo / copy_hi: only the bytes of the parts [copy_lo, copy_hi) are copied (shard-only staging; -1 = all).  The device buffer
// still spans ALL parts and every part keeps its global offset, so that record values, offsets and flags stay global.
static int load_parts_impl(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, const uint32_t* part_id, int64_t n_parts, int ascii_on_device,
                           int n_chunks, int64_t uniform_hint = 0, int64_t copy_lo = -1, int64_t copy_hi = -1) {
  if (!c) return fail(NRP_E_ARG, "null context");
  const bool hinted = uniform_hint > 0;
  if (n_parts < 0 || (!hinted && !offsets) || (hinted && uniform_hint >= ((int64_t)1 << 30))) return fail(NRP_E_ARG, "bad part buffers");
  if (n_parts >= 0xfffffff0ll) return fail(NRP_E_UNSUPPORTED, "more than 2^32-16 parts");
  if (!hinted && offsets[0] != 0) return fail(NRP_E_ARG, "offsets[0] must be 0");
  const int64_t total = hinted ? n_parts * uniform_hint : offsets[n_parts];
  if (total < 0) return fail(NRP_E_ARG, "negative total length");
  if (n_parts > 0 && !ascii && total > 0) return fail(NRP_E_ARG, "bad part buffers");
  if (copy_lo < 0) { copy_lo = 0; copy_hi = n_parts; }
  if (copy_lo > copy_hi || copy_hi > n_parts) return fail(NRP_E_ARG, "bad staging range");
  CU(cudaSetDevice(c->device));
  c->have_parts = false; c->have_result = false;
  if (c->early_valid) { CU(cudaStreamSynchronize(c->fetch_stream)); c->early_valid = false; }     // an early fetch nobody collected still reads the old result
  auto first_byte = [&](int64_t part) -> int64_t { return hinted ? part * uniform_hint : offsets[part]; };
  const int64_t range_lo = first_byte(copy_lo), range_hi = first_byte(copy_hi);     // bytes that travel
  // the big copy goes first: the DMA runs while the host validates the offsets and stages the small arrays
  const int64_t n_tiles = (total + kTileBytes - 1) / kTileBytes;
  const size_t padded = (size_t)n_tiles * kTileBytes + 256;
  if (padded > c->ascii.cap && c->homes.n > 0)
    return fail(NRP_E_STATE, "the part buffer is mapped by peers and would be re-allocated: call nrp_peer_close on every rank first");
  TRY(c->ascii.ensure(padded));
  if (c->pending_chunks > 0) { CU(cudaStreamSynchronize(c->copy_stream)); c->pending_chunks = 0; }     // an earlier streamed batch nobody consumed
  CU(cudaStreamSynchronize(c->stream));                     // nothing of the previous batch is still in flight (staging buffers are reused)
  const bool streamed = n_chunks > 1 && !ascii_on_device && range_hi > range_lo && copy_hi - copy_lo > 1;
  if (!streamed) {
    CU(cudaMemsetAsync(c->ascii.as<uint8_t>() + total, 0, padded - (size_t)total, c->stream));
    if (range_hi > range_lo)
      CU(cudaMemcpyAsync(c->ascii.as<uint8_t>() + range_lo, ascii + range_lo, (size_t)(range_hi - range_lo),
                         ascii_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  } else {
    // pieces = contiguous part ranges of about equal bytes, copied on a second stream; each records an event when it has landed
    if (!c->copy_stream) CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaMemsetAsync(c->ascii.as<uint8_t>() + total, 0, padded - (size_t)total, c->copy_stream));
    int64_t p_lo = copy_lo;
    int used = 0;
    for (int i = 0; i < n_chunks && p_lo < copy_hi; ++i) {
      int64_t p_hi = copy_hi;
      if (i + 1 < n_chunks) {
        const int64_t target = range_lo + (range_hi - range_lo) / n_chunks * (i + 1);
        p_hi = hinted ? (target + uniform_hint - 1) / uniform_hint : std::lower_bound(offsets + p_lo + 1, offsets + copy_hi, target) - offsets;
        if (p_hi <= p_lo) p_hi = p_lo + 1;
        if (p_hi > copy_hi) p_hi = copy_hi;
      }
      const int64_t b_lo = first_byte(p_lo), b_hi = first_byte(p_hi);
      if ((size_t)used >= c->chunks.size()) { nrp_ctx::Chunk ch{}; CU(cudaEventCreateWithFlags(&ch.landed, cudaEventDisableTiming)); c->chunks.push_back(ch); }
      nrp_ctx::Chunk& ch = c->chunks[used++];
      ch.part_lo = p_lo; ch.part_hi = p_hi;
      ch.tile_lo = b_lo / kTileBytes;
      ch.n_tiles = b_hi > b_lo ? (b_hi + kTileBytes - 1) / kTileBytes - ch.tile_lo : 0;
      if (b_hi > b_lo) CU(cudaMemcpyAsync(c->ascii.as<uint8_t>() + b_lo, ascii + b_lo, (size_t)(b_hi - b_lo), cudaMemcpyHostToDevice, c->copy_stream));
      CU(cudaEventRecord(ch.landed, c->copy_stream));
      p_lo = p_hi;
    }
    c->pending_chunks = used;
  }
  // offsets + ids travel through one page-locked staging buffer (the caller's arrays are usually pageable; ids that already
  // are page-locked are copied from where they lie)
  bool ids_pinned = false;
  if (part_id && n_parts > 0) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, part_id) == cudaSuccess) ids_pinned = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  const size_t off_bytes = sizeof(int64_t) * (size_t)(n_parts + 1), id_bytes = part_id ? sizeof(uint32_t) * (size_t)n_parts : 0;
  const size_t off_stage = hinted ? 0 : off_bytes;
  TRY(c->h_upload.ensure(off_stage + (ids_pinned ? 0 : id_bytes) + 16));
  int64_t* st_off = c->h_upload.as<int64_t>();
  uint32_t* st_id = reinterpret_cast<uint32_t*>(c->h_upload.as<unsigned char>() + off_stage);
  // shortest / longest part in one branch-free sweep (the compiler vectorises it), then one memcpy into the staging buffer,
  // which also serves as the context's host copy of the offsets until the next batch
  int64_t min_len = n_parts > 0 ? INT64_MAX : 0, max_len = n_parts > 0 ? INT64_MIN : 0;
  if (hinted) min_len = max_len = n_parts > 0 ? uniform_hint : 0;
  else
    for (int64_t i = 0; i < n_parts; ++i) {
      const int64_t len = offsets[i + 1] - offsets[i];
      min_len = len < min_len ? len : min_len;
      max_len = len > max_len ? len : max_len;
    }
  if (min_len < 0) {
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    c->pending_chunks = 0;
    return fail(NRP_E_ARG, "offsets must be non-decreasing");
  }
  int uniform = (n_parts > 0 && min_len == max_len && max_len < ((int64_t)1 << 30)) ? (int)max_len : 0;
  TRY(c->off.ensure(off_bytes));
  if (uniform > 0) {
    // parts of one length: the offsets are i * L -- written on the device, nothing to copy or to keep on the host
    NRP_LAUNCH((k_uniform_offsets<<<grid_for(n_parts + 1, 256), 256, 0, c->stream>>>(c->off.as<int64_t>(), n_parts, (int64_t)uniform)));
    c->h_off = nullptr;
  } else {
    memcpy(st_off, offsets, off_bytes);
    CU(cudaMemcpyAsync(c->off.ptr, st_off, off_bytes, cudaMemcpyHostToDevice, c->stream));
    c->h_off = st_off;
  }
  c->n_parts = n_parts; c->total_len = total; c->uniform_len = uniform; c->max_len = max_len;
  c->n_tiles = n_tiles;
  c->have_ids = part_id != nullptr;
  if (part_id && n_parts > 0) {
    const uint32_t* src = part_id;
    if (!ids_pinned) { memcpy(st_id, part_id, id_bytes); src = st_id; }
    TRY(c->part_id.ensure(id_bytes));
    if (streamed) {
      // behind the pieces on the copy stream (the ids are read last, by the edge expansion); an empty piece carries the event
      if ((size_t)c->pending_chunks >= c->chunks.size()) { nrp_ctx::Chunk ch{}; CU(cudaEventCreateWithFlags(&ch.landed, cudaEventDisableTiming)); c->chunks.push_back(ch); }
      nrp_ctx::Chunk& ch = c->chunks[c->pending_chunks++];
      ch.part_lo = ch.part_hi = copy_hi; ch.tile_lo = 0; ch.n_tiles = 0;
      CU(cudaMemcpyAsync(c->part_id.ptr, src, id_bytes, cudaMemcpyHostToDevice, c->copy_stream));
      CU(cudaEventRecord(ch.landed, c->copy_stream));
    } else {
      CU(cudaMemcpyAsync(c->part_id.ptr, src, id_bytes, cudaMemcpyHostToDevice, c->stream));
    }
  }
  TRY(c->tile_first.ensure(sizeof(uint32_t) * (size_t)(c->n_tiles + 2)));
  if (uniform == 0 && n_parts > 0)
    NRP_LAUNCH((k_tile_first_part<<<grid_for(c->n_tiles + 1, 256), 256, 0, c->stream>>>(c->off.as<int64_t>(), n_parts, total, c->n_tiles,
                                                                            c->tile_first.as<uint32_t>())));
  CU(cudaGetLastError());
  // offsets and ids were staged in library memory; a blocking load also waits for the caller's ASCII buffer to be consumed
  if (!streamed) CU(cudaStreamSynchronize(c->stream));
  c->part_base = 0;
  c->shard_lo = 0; c->shard_hi = n_parts; c->shard_tile_lo = 0; c->shard_tiles = c->n_tiles;
  c->staged_lo = copy_lo; c->staged_hi = copy_hi;
  c->n_owners = 0;
  c->have_parts = true;
  return NRP_OK;
}

int nrp_load_parts(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, const uint32_t* part_id, int64_t n_parts, int ascii_on_device) {
  return load_parts_impl(c, ascii, offsets, part_id, n_parts, ascii_on_device, 1);
}

int nrp_load_parts_streamed(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, const uint32_t* part_id, int64_t n_parts, int n_chunks) {
  if (n_chunks < 1 || n_chunks > 64) return fail(NRP_E_ARG, "n_chunks must be in [1, 64]");
  return load_parts_impl(c, ascii, offsets, part_id, n_parts, 0, n_chunks);
}

int nrp_load_parts_uniform(nrp_ctx* c, const uint8_t* ascii, int64_t n_parts, int part_len, const uint32_t* part_id, int n_chunks) {
  if (n_chunks < 1 || n_chunks > 64) return fail(NRP_E_ARG, "n_chunks must be in [1, 64]");
  if (part_len < 1) return fail(NRP_E_ARG, "part_len must be positive");
  return load_parts_impl(c, ascii, nullptr, part_id, n_parts, 0, n_chunks, part_len);
}

int nrp_load_parts_shard(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, int part_len, const uint32_t* part_id, int64_t n_parts,
                         int64_t stage_lo, int64_t stage_hi, int n_chunks) {
  if (n_chunks < 1 || n_chunks > 64) return fail(NRP_E_ARG, "n_chunks must be in [1, 64]");
  if (!offsets && part_len < 1) return fail(NRP_E_ARG, "offsets or a positive part_len is required");
  if (stage_lo < 0 || stage_hi < stage_lo || stage_hi > n_parts) return fail(NRP_E_ARG, "bad staging range");
  return load_parts_impl(c, ascii, offsets, part_id, n_parts, 0, n_chunks, offsets ? 0 : part_len, stage_lo, stage_hi);
}

}  // extern "C"

namespace {

int64_t windows_in_range(const nrp_ctx* c, int64_t lo, int64_t hi, int k) {
  if (c->uniform_len > 0) return c->uniform_len >= k ? (hi - lo) * (int64_t)(c->uniform_len - k + 1) : 0;
  int64_t m = 0;
  for (int64_t i = lo; i < hi; ++i) { int64_t w = c->h_off[i + 1] - c->h_off[i] - k + 1; if (w > 0) m += w; }
  return m;
}

int ensure_record_buffers(nrp_ctx* c, int64_t n_records, size_t key_bytes, int64_t n_segments = 0) {
  const size_t rec = (size_t)n_records + 16;
  TRY(c->keysA.ensure(key_bytes * rec)); TRY(c->valsA.ensure(4 * rec));
  TRY(c->keysB.ensure(key_bytes * rec)); TRY(c->valsB.ensure(4 * rec));
  const size_t leaf_max = (size_t)1 << kMaxDirectBits;
  TRY(c->hist.ensure(4 * (((size_t)1 << kMaxLeafBits) + 2))); TRY(c->leaf_off.ensure(4 * (((size_t)1 << kMaxLeafBits) + 2)));
  TRY(c->cursor1.ensure(4 * (leaf_max + 2))); TRY(c->cursor2.ensure(4 * (leaf_max + 2)));
  TRY(c->region_ends.ensure(4 * (size_t)(kLevelBuckets * kMaxPeers + 4)));
  const size_t segs = (size_t)std::max<int64_t>(n_segments, kLevelBuckets * kMaxPeers) + 4;
  TRY(c->seg_off.ensure(4 * segs)); TRY(c->tile_prefix.ensure(4 * segs));
  TRY(c->tile_desc.ensure(16 * ((size_t)n_records / kSplitTile + segs + 4)));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems(((int64_t)1 << kMaxLeafBits) + 2)));
  return NRP_OK;
}

// a streamed batch (nrp_load_parts_streamed): make the context's stream wait for every piece still in flight
int wait_chunks(nrp_ctx* c) {
  for (int i = 0; i < c->pending_chunks; ++i) CU(cudaStreamWaitEvent(c->stream, c->chunks[i].landed, 0));
  c->pending_chunks = 0;
  return NRP_OK;
}

bool uniform_kernel_runs(const nrp_ctx* c, int k);

// ---- stage: exclusion flags that need no grouping (palindromes, background) over the shard's tiles.
// allow_fold (single-GPU build): when the part-aligned extraction kernel is going to run, the palindrome test and -- for a
// background with K == k -- the background probe are left to it (FOLD in extract.cuh): no pass of their own over the parts.
int stage_flags(nrp_ctx* c, int k, int internal_repeats, StageTimer& tm, bool allow_fold = false) {
  cudaStream_t st = c->stream;
  tm.begin(NRP_T_FLAGS);
  const bool want_pal = !internal_repeats && (k % 2) == 0, want_bg = c->bg_n > 0;
  const bool fold = allow_fold && c->opt_fold_flags && (want_pal || want_bg) && uniform_kernel_runs(c, k);
  c->fold_pal = fold && want_pal;
  c->fold_bg = fold && want_bg && c->bg_K == k;
  const bool pass_pal = want_pal && !c->fold_pal, pass_bg = want_bg && !c->fold_bg;
  if (pass_pal || pass_bg || c->shard_lo != 0 || c->shard_hi != c->n_parts) TRY(wait_chunks(c));
  const Parts P = parts_view(c);
  TRY(c->flags.ensure((size_t)c->n_parts + 16));
  TRY(c->last_single.ensure(sizeof(int32_t) * (size_t)(c->n_parts + 1)));
  CU(cudaMemsetAsync(c->flags.ptr, 0, (size_t)c->n_parts + 16, st));
  c->flags_set = c->shard_tiles > 0 && (pass_pal || pass_bg);
  if (c->shard_tiles > 0) {
    if (pass_pal)
      NRP_LAUNCH((k_flag_windows<uint32_t><<<(unsigned)c->shard_tiles, kTileThreads, 0, st>>>(P, k, 0, nullptr, 0, c->flags.as<uint8_t>())));
    if (pass_bg) {
      if (c->bg_key_bytes == 4)
        NRP_LAUNCH((k_flag_windows<uint32_t><<<(unsigned)c->shard_tiles, kTileThreads, 0, st>>>(P, c->bg_K, 1, c->bg_table.as<uint32_t>(),
                                                                                              c->bg_log_slots, c->flags.as<uint8_t>())));
      else
        NRP_LAUNCH((k_flag_windows<uint64_t><<<(unsigned)c->shard_tiles, kTileThreads, 0, st>>>(P, c->bg_K, 1, c->bg_table.as<uint64_t>(),
                                                                                              c->bg_log_slots, c->flags.as<uint8_t>())));
    }
  }
  CU(cudaGetLastError());
  return NRP_OK;
}

// ---- stage: records -> leaf buckets -> candidates (records whose key repeats inside its leaf bucket).
// Where the records of a bulk stage come from: the staged part buffer (extraction fused into the first split), or a
// segmented record list in HBM -- a flat array (one segment), or the fixed regions the senders of the direct sharded
// path filled, which are already partitioned by the first r_done hash bits.
template <typename KeyT>
struct BulkSource {
  bool from_parts = true;              // records are extracted from the staged parts
  bool packed = false;                 // the record list holds packed words (direct sharded path), not (key, value) arrays
  int drop = 0;                        // packed words: top hash bits implied by the segment's level-1 bucket
  int owner_bits = 0;                  // packed words of the sharded path: the owner is the top hash bits (this rank's records)
  const void* keys = nullptr;
  const uint32_t* vals = nullptr;
  SegExtents ext{nullptr, 0, nullptr, 0};
  uint32_t n_seg = 0;                  // segments of the record list
  uint32_t segs_per_bucket = 1;        // pre-partitioned input: consecutive segments that belong to one level-1 bucket
  bool prepartitioned = false;         // the segments are (level-1 bucket, sender) regions: level 1 is done
  int r_done = 0;                      // pre-partitioned input: hash bits already consumed
  int64_t m_upper = 0;                 // hard upper bound of the number of records (buffer sizes)
  int64_t m_plan = 0;                  // expected number of records (level planning); 0 = m_upper
  int bits_hint = 0;                   // pre-partitioned input: total leaf bits the senders planned for (0 = decide here)
};

struct LevelPlan { int bits, r1, r2; uint32_t cap1, cap2; };

// region size of the direct path: mean + 12.5 % + 8 sigma (Poisson) + slack, a multiple of 4 records
inline uint32_t region_cap(double mean, int slack) {
  const double v = mean + mean / 8.0 + 8.0 * std::sqrt(mean) + slack;
  return (uint32_t)(((int64_t)v + 4) & ~3ll);
}

// number of leaf bits for m records when at most max_bits are available (and at least min_bits must be used)
int leaf_bits_for(const nrp_ctx* c, int64_t m, int min_bits, int max_bits, int64_t target = 0) {
  if (target <= 0) target = c->opt_leaf_target;
  int bits = std::max(min_bits, 1);
  while (bits < max_bits && (m >> bits) > target) ++bits;
  // the small-leaf target is out of reach: take the fewest bits that still fit the large filter variant
  if (target == kDefaultLeafTarget && (m >> bits) > kDefaultLeafTarget)
    while (bits > std::max(min_bits, 1) && (m >> (bits - 1)) <= kLargeLeafTarget) --bits;
  return bits;
}

// Exact path: leaf bits (<= 15, the shared-memory histogram) and their division into two levels.
LevelPlan plan_levels(const nrp_ctx* c, int64_t m, bool /*direct*/, int /*fixed_r1*/ = -1) {
  const int lv = (int)c->opt_max_level_bits;
  const int bits = leaf_bits_for(c, m, 1, std::min(kMaxLeafBits, 2 * lv));
  LevelPlan p;
  p.bits = bits;
  p.r1 = bits <= std::min(lv, 8) ? bits : std::max(1, std::min(std::min(lv, bits - 1), (bits + 1) / 2 + (int)c->opt_r1_bias));
  if (bits - p.r1 > lv) p.r1 = bits - lv;
  p.r2 = bits - p.r1;
  p.cap1 = p.cap2 = 0;
  return p;
}

// Direct path: `pre` hash bits are already consumed by a pre-partitioned source (0 otherwise); the remaining leaf bits are
// split over one or two levels of at most opt_max_level_bits (<= 9) bits.  cum[i] = bits consumed after level i.
struct DirectPlan { int bits, pre, n_levels; int r[2]; int cum[2]; uint32_t cap[2]; };
DirectPlan plan_direct(const nrp_ctx* c, int64_t m, int pre, bool must_split, int bits_hint = 0, int64_t leaf_target = 0) {
  const int lv = (int)c->opt_max_level_bits;
  DirectPlan p;
  p.pre = pre;
  p.bits = bits_hint > 0 ? bits_hint : leaf_bits_for(c, m, pre + (must_split ? 1 : 0), std::min(kMaxDirectBits, pre + 2 * lv), leaf_target);
  const int rest = p.bits - pre;
  if (rest <= std::min(lv, 8) || (pre > 0 && rest <= lv)) {
    p.n_levels = 1; p.r[0] = rest; p.r[1] = 0;
  } else {
    p.n_levels = 2;
    p.r[0] = std::max(1, std::min(std::min(lv, rest - 1), (rest + 1) / 2 + (int)c->opt_r1_bias));
    if (rest - p.r[0] > lv) p.r[0] = rest - lv;
    p.r[1] = rest - p.r[0];
  }
  p.cum[0] = pre + p.r[0]; p.cum[1] = p.cum[0] + p.r[1];
  for (int i = 0; i < 2; ++i)
    p.cap[i] = region_cap((double)m / (double)(1ll << p.cum[i]), i + 1 == p.n_levels ? 64 : 1024);
  return p;
}

// Can the records of this job travel as packed words?  n = 2k hash bits, `drop` of them implied by the level-1 region,
// `bits` leaf bits in total (the filter mixes the >= 11 hash bits below them into its bit-map and table indices).
// (owner_bits: sharded path, the top hash bits that select the owner; `bits` then counts the owner-local leaf bits)
bool packed_fits(const nrp_ctx* c, int k, int drop, int bits, int owner_bits = 0, int vbits = -1) {
  const int n = 2 * k - owner_bits;
  if (vbits < 0) vbits = c->vbits;
  return c->opt_packed && !c->opt_no_bloom && n - drop + vbits <= 64 && n >= bits + 11 && n - drop >= 1 && 2 * k < 64;
}
PackFmt pack_format(const nrp_ctx* c, int k, int drop, int owner_bits = 0, uint32_t owner = 0) {
  // the stored hash bits go to the top of the word (see PackFmt); packed_fits has checked that the value fits below them
  const int stored = std::max(1, 2 * k - owner_bits - drop);
  return PackFmt{make_bijhash(k), std::max(c->vbits, 64 - stored), drop, owner_bits, owner};
}

// expected fraction of records whose k-mer occurs again: all records of the job over the canonical k-mer space (random DNA)
inline double shared_fraction_of(const nrp_ctx* c, int k, int64_t m) {
  const double all_records = c->n_owners > 1 ? (double)m * c->n_owners : (double)m;
  return k >= 31 ? 0.0 : std::min(1.0, 2.0 * all_records / std::ldexp(1.0, 2 * k));
}

// suspects per leaf the warp-private filter may expect: shared records + bit-map collisions (mean^2 / bits)
inline double warp_filter_suspects(double mean_leaf, double shared_fraction, int map_bits) {
  return mean_leaf * shared_fraction * 1.5 + mean_leaf * mean_leaf / (double)map_bits;
}

// Level plan of a direct build whose records come from the staged parts.  *warp_filter: the leaves are planned small (<= 800
// records on average) and filtered warp-privately -- when the records pack, the leaf bits allow it and few suspects are expected.
DirectPlan choose_direct_plan(const nrp_ctx* c, int k, int64_t m_max, bool* packed_out, bool* warp_filter) {
  *warp_filter = false;
  if (c->opt_filter_warp && !c->opt_no_bloom && c->opt_leaf_target == kDefaultLeafTarget && c->opt_filter_cap == kFilterCap && m_max > 0) {
    const double shared = shared_fraction_of(c, k, m_max);
    for (int64_t target = c->opt_warp_leaf_target; target >= 100; target /= 2) {      // smaller leaves when many suspects are expected
      const DirectPlan small = plan_direct(c, m_max, 0, false, 0, target);
      const double mean = (double)m_max / (double)((int64_t)1 << small.bits);
      int64_t total_max = 0;
      for (int i = 0; i < small.n_levels; ++i) total_max = std::max(total_max, ((int64_t)1 << small.cum[i]) * small.cap[i]);
      if (mean > (double)target || total_max >= 0xffffe000ll || !packed_fits(c, k, small.cum[0], small.bits)) break;
      // (up to 56 expected suspects per leaf: room for 96; up to 150: the variant with room for 192 -- config 3, where 6 % of the
      // records share their k-mer and two 9-bit levels cannot make the leaves smaller than ~700 records)
      if (warp_filter_suspects(mean, shared, 512 * 32) <= kWarpSuspectsWide) {
        *warp_filter = true; *packed_out = true;
        return small;
      }
    }
  }
  const DirectPlan dp = plan_direct(c, m_max, 0, false, 0);
  *packed_out = packed_fits(c, k, dp.cum[0], dp.bits);
  return dp;
}

// Will the first split of a single-GPU build of the staged parts be the part-aligned kernel?  (The same decisions as
// stage_bulk_direct takes; the extraction kernel then also computes the flags stage_flags left to it.)
bool uniform_kernel_runs(const nrp_ctx* c, int k) {
  UniPlan up;
  if (!c->opt_direct || !c->opt_uniform_kernel || c->uniform_len <= 0 || c->n_owners > 0 || !uni_plan(c->uniform_len, k, &up)) return false;
  if (c->shard_m_max <= 0) return false;
  bool packed = false, warp_filter = false;
  const DirectPlan dp = choose_direct_plan(c, k, c->shard_m_max, &packed, &warp_filter);
  int64_t total_max = 0;
  for (int i = 0; i < dp.n_levels; ++i) total_max = std::max(total_max, ((int64_t)1 << dp.cum[i]) * dp.cap[i]);
  if (total_max >= 0xffffe000ll) return false;
  return packed;
}

template <typename KeyT>
int launch_filter(nrp_ctx* c, bool packed, PackFmt pf, const void* leaf_keys, const uint32_t* leaf_vals, LeafExtents leaves, uint32_t n_leaf, int bits,
                  int64_t mean_leaf, KeyT* out_keys, uint32_t* out_vals, double shared_fraction = 1.0, bool warp_filter = false) {
  cudaStream_t st = c->stream;
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  if (warp_filter && packed) {
    // small leaves, one warp each (filter_warp.cuh): as many ten-warp CTAs per SM as shared memory allows
    uint32_t* over = reinterpret_cast<uint32_t*>(cnt + CNT_OVERSIZED);
    uint32_t* unsrt = reinterpret_cast<uint32_t*>(cnt + CNT_UNSORTED);
    const bool tiny = mean_leaf <= 400 && warp_filter_suspects((double)mean_leaf, shared_fraction, 256 * 32) <= 36.0;
    const bool wide = !tiny && warp_filter_suspects((double)mean_leaf, shared_fraction, 512 * 32) > kWarpSuspects;
    const size_t smem = tiny ? kFwWarps * sizeof(FwWarp<256, 64>) : wide ? kFwWarps * sizeof(FwWarp<512, 192>) : kFwWarps * sizeof(FwWarp<512, 96>);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (228 * 1024) / (smem + 1024)));      // 1 KB per CTA is reserved by the system
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(((int64_t)n_leaf + kFwWarps - 1) / kFwWarps, (int64_t)c->n_sms * per_sm));
#define NRP_FW(MAPW_, SUSP_, VHI_)                                                                                                              \
  NRP_LAUNCH((k_filter_warp<KeyT, MAPW_, SUSP_, VHI_><<<grid, kFwThreads, smem, st>>>((const uint64_t*)leaf_keys, leaves, n_leaf, bits, pf, out_keys, \
                                                                                     out_vals, cnt + CNT_CAND, over, unsrt)))
    const bool vhi = pf.vbits >= 32;       // where the stored k-mer bits begin: in the high word of the record, or below it
    if (tiny) { if (vhi) NRP_FW(256, 64, true); else NRP_FW(256, 64, false); }
    else if (wide) { if (vhi) NRP_FW(512, 192, true); else NRP_FW(512, 192, false); }
    else { if (vhi) NRP_FW(512, 96, true); else NRP_FW(512, 96, false); }
#undef NRP_FW
    CU(cudaGetLastError());
    c->counts[NRP_C_FILTER_WARP] = tiny ? 2 : wide ? 3 : 1;
    return NRP_OK;
  }
  // leaves that average <= 2750 records use the half-size variant (two CTAs per SM); larger ones the full one
  const bool small = mean_leaf <= 2750;
  const uint32_t cap = (uint32_t)std::min<int64_t>(c->opt_filter_cap, small ? kFilterCapSmall : kFilterCap);
  uint32_t* n_over = reinterpret_cast<uint32_t*>(cnt + CNT_OVERSIZED);
  uint32_t* unsorted = reinterpret_cast<uint32_t*>(cnt + CNT_UNSORTED);
  const unsigned grid_small = (unsigned)std::min<int64_t>(n_leaf, (int64_t)c->n_sms * 2), grid_large = (unsigned)std::min<int64_t>(n_leaf, (int64_t)c->n_sms);
  if (!c->opt_no_bloom) {
    if (c->opt_filter_tiny && packed && mean_leaf <= 1400)    // four 256-thread CTAs per SM, leaves of up to 2048 records
      NRP_LAUNCH((k_filter_bloom<KeyT, 256, 8, 1024, 4, 256, true><<<(unsigned)std::min<int64_t>(n_leaf, (int64_t)c->n_sms * 4), 256,
                                                                    filter_bloom_smem<8, true, 2048, 1024, 256>(), st>>>(
          leaf_keys, leaf_vals, leaves, n_leaf, bits, pf, std::min<uint32_t>(cap, 2048u), out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
    else if (small && packed && c->opt_filter_wide && (double)mean_leaf * (shared_fraction + (double)mean_leaf / 65536.0) <= 160.0)
      // few suspects expected per leaf (shared records + bit-map collisions): three 256-thread CTAs per SM, 16 records per thread
      NRP_LAUNCH((k_filter_bloom<KeyT, 256, 16, 2048, 3, 256, true><<<(unsigned)std::min<int64_t>(n_leaf, (int64_t)c->n_sms * 3), 256,
                                                                     filter_bloom_smem<8, true, 4096, 2048, 256>(), st>>>(
          leaf_keys, leaf_vals, leaves, n_leaf, bits, pf, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
    else if (small && packed)      // two 512-thread CTAs per SM, leaves of up to 4096 records
      NRP_LAUNCH((k_filter_bloom<KeyT, 512, 8, 2048, 2, 512, true><<<grid_small, 512, filter_bloom_smem<8, true, 4096, 2048, 512>(), st>>>(
          leaf_keys, leaf_vals, leaves, n_leaf, bits, pf, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
    else if (small)
      NRP_LAUNCH((k_filter_bloom<KeyT, 512, 8, 2048, 2, 512, false><<<grid_small, 512, filter_bloom_smem<(int)sizeof(KeyT), false, 4096, 2048, 512>(), st>>>(
          leaf_keys, leaf_vals, leaves, n_leaf, bits, pf, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
    else if (packed)               // one 1024-thread CTA per SM, leaves of up to 8192 records
      NRP_LAUNCH((k_filter_bloom<KeyT, kFilterThreads, 8, 4096, 1, 1024, true><<<grid_large, kFilterThreads, filter_bloom_smem<8, true, 8192, 4096, 1024>(), st>>>(
          leaf_keys, leaf_vals, leaves, n_leaf, bits, pf, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
    else
      NRP_LAUNCH((k_filter_bloom<KeyT, kFilterThreads, 8, 4096, 1, 1024, false><<<grid_large, kFilterThreads,
                                                                                 filter_bloom_smem<(int)sizeof(KeyT), false, 8192, 4096, 1024>(), st>>>(
          leaf_keys, leaf_vals, leaves, n_leaf, bits, pf, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
  } else {
    const size_t fsm = small ? filter_smem_small<KeyT>() : filter_smem_large<KeyT>();
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (220 * 1024) / (fsm + 18 * 1024)));   // + static smem of the leaf ordering
    const unsigned fgrid = (unsigned)std::min<int64_t>(n_leaf, (int64_t)c->n_sms * per_sm);
    if (small)
      NRP_LAUNCH((k_filter<KeyT, kFilterThreads, kFilterItemsSmall, kFilterCapSmall, kFilterStaged><<<fgrid, kFilterThreads, fsm, st>>>(
          (const KeyT*)leaf_keys, leaf_vals, leaves, n_leaf, bits, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
    else
      NRP_LAUNCH((k_filter<KeyT, kFilterThreads, kFilterItems, kFilterCap, false><<<fgrid, kFilterThreads, fsm, st>>>(
          (const KeyT*)leaf_keys, leaf_vals, leaves, n_leaf, bits, cap, out_keys, out_vals, cnt + CNT_CAND, n_over, unsorted)));
  }
  CU(cudaGetLastError());
  return NRP_OK;
}

// tile descriptors of a segmented record list; the number of tiles ends up in tile_prefix[n_seg] (device)
int build_tiles(nrp_ctx* c, SegExtents ext, uint32_t n_seg, uint32_t segs_per_bucket, int r_child, unsigned long long* n_records_dev) {
  cudaStream_t st = c->stream;
  NRP_LAUNCH((k_segment_scan<<<1, 1024, 0, st>>>(ext, n_seg, c->tile_prefix.as<uint32_t>(), n_records_dev)));
  NRP_LAUNCH((k_segment_fill<<<(n_seg + 7) / 8, 256, 0, st>>>(ext, n_seg, c->tile_prefix.as<uint32_t>(), segs_per_bucket, r_child, c->tile_desc.as<uint4>())));
  CU(cudaGetLastError());
  return NRP_OK;
}

// Direct path: fixed regions, no histogram.  *overflowed is set when a region was too small (nothing usable was produced).
// One or two split levels; the first reads the staged parts (extraction fused in), a flat record list, or the (bucket, sender)
// regions of a pre-partitioned source; every level scatters into the fixed regions of the next buffer.
template <typename KeyT>
int stage_bulk_direct(nrp_ctx* c, int k, const BulkSource<KeyT>& src, StageTimer& tm, bool* overflowed) {
  cudaStream_t st = c->stream;
  const bool from_ascii = src.from_parts;
  const bool prepart = !from_ascii && src.prepartitioned;
  const int64_t m_max = from_ascii ? c->shard_m_max : (src.m_plan > 0 ? src.m_plan : src.m_upper);
  const Parts P = parts_view(c);
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  uint32_t* overflow = reinterpret_cast<uint32_t*>(cnt + CNT_OVERFLOW);
  bool chosen_packed = false, warp_filter = false;
  const DirectPlan dp = from_ascii ? choose_direct_plan(c, k, m_max, &chosen_packed, &warp_filter)
                                   : plan_direct(c, m_max, prepart ? src.r_done : 0, prepart, prepart ? src.bits_hint : 0);
  const int bits = dp.bits;
  const uint32_t n_leaf = 1u << bits;
  *overflowed = false;
  int64_t total_max = 0;
  for (int i = 0; i < dp.n_levels; ++i) total_max = std::max(total_max, ((int64_t)1 << dp.cum[i]) * dp.cap[i]);
  if (total_max >= 0xffffe000ll) { *overflowed = true; return NRP_OK; }   // beyond 32-bit record indices: exact path
  c->counts[NRP_C_BUCKETS] = (int64_t)n_leaf;
  // records as single 64-bit words when hash and value fit (a packed source dictates it, with the bits its regions imply)
  const bool packed = from_ascii ? chosen_packed : src.packed;
  const PackFmt pf = from_ascii ? pack_format(c, k, dp.cum[0]) : pack_format(c, k, src.drop, src.owner_bits, (uint32_t)c->my_rank);
  const int64_t n_seg_max = std::max<int64_t>((int64_t)src.n_seg, dp.n_levels > 1 ? ((int64_t)1 << dp.cum[0]) : 0);
  TRY(ensure_record_buffers(c, std::max<int64_t>(total_max, m_max), packed ? 8 : sizeof(KeyT), n_seg_max));
  CU(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * CNT_N, st));
  c->n_records = 0;

  void* buf_keys[2] = {c->keysA.ptr, c->keysB.ptr};
  uint32_t* buf_vals[2] = {c->valsA.as<uint32_t>(), c->valsB.as<uint32_t>()};
  // the records a level reads: the source first, then the regions the previous level filled
  const void* in_keys = src.keys; const uint32_t* in_vals = src.vals;
  SegExtents in_ext = src.ext;
  uint32_t in_seg = src.n_seg, in_spb = prepart ? src.segs_per_bucket : 0xffffffffu;
  int in_done = dp.pre;
  bool counted = false;
  for (int lvl = 0; lvl < dp.n_levels && m_max > 0; ++lvl) {
    tm.begin(lvl == 0 ? NRP_T_SPLIT1 : NRP_T_SPLIT2);
    const bool last = lvl + 1 == dp.n_levels;
    const int r = dp.r[lvl];
    const uint32_t n_regions = 1u << dp.cum[lvl];
    uint32_t* cursor = last ? c->cursor2.as<uint32_t>() : c->cursor1.as<uint32_t>();
    NRP_LAUNCH((k_init_region_cursors<<<grid_for(n_regions, 256), 256, 0, st>>>(cursor, n_regions, dp.cap[lvl])));
    const RegionLimit lim{dp.cap[lvl], 0u, overflow, nullptr};
    void* out_k = buf_keys[lvl & 1]; uint32_t* out_v = buf_vals[lvl & 1];
    if (lvl == 0 && from_ascii) {
      // a streamed batch is extracted piece by piece, each as soon as its copy has landed (pieces are part ranges; the
      // kernel only forms windows of the parts in its range, so a byte tile shared by two pieces is visited by both)
      const int n_pieces = c->pending_chunks > 0 ? c->pending_chunks : 1;
      for (int piece = 0; piece < n_pieces; ++piece) {
        Parts Q = P;
        int64_t q_tiles = c->shard_tiles;
        if (c->pending_chunks > 0) {
          const nrp_ctx::Chunk& ch = c->chunks[piece];
          CU(cudaStreamWaitEvent(st, ch.landed, 0));
          Q.part_lo = ch.part_lo; Q.part_hi = ch.part_hi; Q.tile_lo = ch.tile_lo; q_tiles = ch.n_tiles;
          if (q_tiles == 0) continue;
        }
        UniPlan up;
        if (packed && c->opt_uniform_kernel && c->uniform_len > 0 && uni_plan(c->uniform_len, k, &up)) {
          // parts of one length: the part-aligned kernel (threads never cross a part boundary, left-aligned rolling values)
          const int64_t n_q = Q.part_hi - Q.part_lo;
          const unsigned ugrid = (unsigned)((n_q + up.G - 1) / up.G);
          const uint8_t* fl = c->flags_set ? c->flags.as<uint8_t>() : nullptr;
          if (n_q <= 0) continue;
          if (c->fold_pal || c->fold_bg) {
            // the flags stage_flags left to this kernel: palindromic windows and / or the background probe (K == k)
            FoldArgs fa{c->flags.as<uint8_t>(), c->fold_bg ? c->bg_bits.as<uint32_t>() : nullptr, c->bg_table.ptr, c->bg_log_bits, c->bg_log_slots,
                        c->bg_key_bytes, c->fold_pal ? 1 : 0};
            if (k > 16)
              NRP_LAUNCH((k_split_uniform<true, 0, true><<<ugrid, kSplitThreads, sizeof(UniSmem), st>>>(
                  Q.ascii, fl, up, Q.part_lo, Q.part_hi, k, r, 0, c->part_base, c->j_bits, pf, cursor, (uint64_t*)out_k, lim, PeerBuffers{}, fa)));
            else
              NRP_LAUNCH((k_split_uniform<false, 0, true><<<ugrid, kSplitThreads, sizeof(UniSmem), st>>>(
                  Q.ascii, fl, up, Q.part_lo, Q.part_hi, k, r, 0, c->part_base, c->j_bits, pf, cursor, (uint64_t*)out_k, lim, PeerBuffers{}, fa)));
            c->counts[NRP_C_UNIFORM] = 2;
            continue;
          }
          if (k > 16)
            NRP_LAUNCH((k_split_uniform<true, 0><<<ugrid, kSplitThreads, sizeof(UniSmem), st>>>(
                Q.ascii, fl, up, Q.part_lo, Q.part_hi, k, r, 0, c->part_base, c->j_bits, pf, cursor, (uint64_t*)out_k, lim, PeerBuffers{})));
          else
            NRP_LAUNCH((k_split_uniform<false, 0><<<ugrid, kSplitThreads, sizeof(UniSmem), st>>>(
                Q.ascii, fl, up, Q.part_lo, Q.part_hi, k, r, 0, c->part_base, c->j_bits, pf, cursor, (uint64_t*)out_k, lim, PeerBuffers{})));
          c->counts[NRP_C_UNIFORM] = 1;
        } else if (packed)
          NRP_LAUNCH((k_split_from_ascii<uint64_t, 0, false, true><<<(unsigned)q_tiles, kTileThreads, sizeof(SplitSmem<uint64_t, true>) + sizeof(TileStage), st>>>(
              Q, c->flags.as<uint8_t>(), k, r, 0, q_tiles, c->part_base, c->j_bits, pf, PassSel{0, 0u, 0u}, cursor,
              (uint64_t*)out_k, out_v, lim, PeerBuffers{}, 0)));
        else
          NRP_LAUNCH((k_split_from_ascii<KeyT, 0, false><<<(unsigned)q_tiles, kTileThreads, sizeof(SplitSmem<KeyT>) + sizeof(TileStage), st>>>(
              Q, c->flags.as<uint8_t>(), k, r, 0, q_tiles, c->part_base, c->j_bits, pf, PassSel{0, 0u, 0u}, cursor,
              (KeyT*)out_k, out_v, lim, PeerBuffers{}, 0)));
      }
      c->pending_chunks = 0;
    } else {
      // children of segment g: the cursors of bucket (g / in_spb), i.e. region index << r (a flat list has one bucket)
      TRY(build_tiles(c, in_ext, in_seg, in_spb, in_spb == 0xffffffffu ? 0 : r, cnt + CNT_RECORDS));
      counted = true;
      const unsigned grid = (unsigned)std::min<int64_t>(m_max / kSplitTile + in_seg, (int64_t)c->n_sms * 4);
      if (packed)
        NRP_LAUNCH((k_split_from_records<uint64_t, false, true><<<grid, kSplitThreads, sizeof(SplitSmem<uint64_t, true>), st>>>(
            (const uint64_t*)in_keys, in_vals, c->tile_desc.as<uint4>(), c->tile_prefix.as<uint32_t>() + in_seg, in_done, r, pf,
            PassSel{0, 0u, 0u}, cursor, (uint64_t*)out_k, out_v, lim)));
      else
        NRP_LAUNCH((k_split_from_records<KeyT, false><<<grid, kSplitThreads, sizeof(SplitSmem<KeyT>), st>>>(
            (const KeyT*)in_keys, in_vals, c->tile_desc.as<uint4>(), c->tile_prefix.as<uint32_t>() + in_seg, in_done, r, pf, PassSel{0, 0u, 0u},
            cursor, (KeyT*)out_k, out_v, lim)));
    }
    in_keys = out_k; in_vals = out_v;
    in_ext = SegExtents{nullptr, 0, cursor, dp.cap[lvl]};
    in_seg = n_regions; in_spb = 1u; in_done = dp.cum[lvl];
  }
  // leaves: the last level's output; candidates go to the other buffer (free, or consumed by the last level)
  const int leaf_buf = (dp.n_levels - 1) & 1;
  const void* leaf_keys = buf_keys[leaf_buf]; const uint32_t* leaf_vals = buf_vals[leaf_buf];
  void* out_keys = buf_keys[leaf_buf ^ 1]; uint32_t* out_vals = buf_vals[leaf_buf ^ 1];
  const uint32_t leaf_cap = dp.cap[dp.n_levels - 1];
  tm.begin(NRP_T_FILTER);
  if (m_max > 0) {
    if (!counted) NRP_LAUNCH((k_region_total<<<1, 1024, 0, st>>>(c->cursor2.as<uint32_t>(), n_leaf, leaf_cap, cnt + CNT_RECORDS)));
    const double shared_fraction = shared_fraction_of(c, k, m_max);
    TRY((launch_filter<KeyT>(c, packed, pf, leaf_keys, leaf_vals, LeafExtents{nullptr, c->cursor2.as<uint32_t>(), leaf_cap}, n_leaf, bits, m_max >> bits,
                             (KeyT*)out_keys, out_vals, shared_fraction, warp_filter)));
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_counters, cnt, sizeof(unsigned long long) * CNT_N, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if ((uint32_t)c->h_counters[CNT_OVERFLOW] != 0) { *overflowed = true; return NRP_OK; }
  c->n_cand = (int64_t)c->h_counters[CNT_CAND];
  c->n_records = (int64_t)c->h_counters[CNT_RECORDS];
  c->counts[NRP_C_OVERSIZED] = (int64_t)(uint32_t)c->h_counters[CNT_OVERSIZED];
  c->counts[NRP_C_CANDIDATES] = c->n_cand;
  c->counts[NRP_C_RUNS] = 1;
  c->counts[NRP_C_DIRECT] = 1;
  c->counts[NRP_C_PACKED] = packed ? 1 : 0;
  c->cand_unsorted = (uint32_t)c->h_counters[CNT_UNSORTED] != 0;
  c->cand_keys = out_keys; c->cand_vals = out_vals;
  c->spare_keys = const_cast<void*>(leaf_keys); c->spare_vals = const_cast<uint32_t*>(leaf_vals);
  return NRP_OK;
}

// Exact path: leaf sizes from a histogram pass, tight offsets; several passes over slices of the hash space when the
// batch is too large for two split levels of a 2^15-entry shared-memory histogram.
template <typename KeyT>
int stage_bulk_exact(nrp_ctx* c, int k, const BulkSource<KeyT>& src, StageTimer& tm) {
  cudaStream_t st = c->stream;
  const bool from_ascii = src.from_parts;
  const int64_t m_max = from_ascii ? c->shard_m_max : src.m_upper;
  const Parts P = parts_view(c);
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  TRY(wait_chunks(c));
  // passes over disjoint slices of the hash space (1 unless the batch is too large for two split levels)
  int pass_bits = 0;
  while (pass_bits < 6 && (m_max >> pass_bits) > c->opt_pass_records) ++pass_bits;
  const int n_passes = 1 << pass_bits;
  const int64_t m_pass = (m_max >> pass_bits) + (pass_bits ? (m_max >> (pass_bits + 3)) + (1 << 20) : 0);   // slack: slices are near-uniform
  const int64_t m_plan = from_ascii || src.m_plan <= 0 ? m_max : src.m_plan;
  const LevelPlan lp = plan_levels(c, m_plan >> pass_bits, false);
  const int bits = lp.bits, r1 = lp.r1, r2 = lp.r2;
  const uint32_t n_leaf = 1u << bits;
  c->counts[NRP_C_BUCKETS] = (int64_t)n_leaf * n_passes;
  c->counts[NRP_C_DIRECT] = 0;
  c->counts[NRP_C_PACKED] = 0;
  TRY(ensure_record_buffers(c, m_pass, sizeof(KeyT), src.n_seg));
  CU(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * CNT_N, st));
  c->n_records = 0;
  KeyT* leaf_keys = nullptr; uint32_t* leaf_vals = nullptr; KeyT* free_keys = nullptr; uint32_t* free_vals = nullptr;
  int64_t cand_so_far = 0;
  const RegionLimit no_limit{0u, 0u, nullptr, nullptr};
  const PackFmt no_pack = pack_format(c, k, 0);
  const KeyT* src_keys = (const KeyT*)src.keys; const uint32_t* src_vals = src.vals;
  if (!from_ascii && src.packed && m_max > 0) {
    // the source holds packed words (direct sharded path): the exact path works on (key, value) arrays at the same indices
    TRY(c->fb_keys.ensure(sizeof(KeyT) * (size_t)(src.m_upper + 16))); TRY(c->fb_vals.ensure(4 * (size_t)(src.m_upper + 16)));
    TRY(build_tiles(c, src.ext, src.n_seg, src.segs_per_bucket, 0, nullptr));          // descriptor .z = level-1 bucket of the segment
    NRP_LAUNCH((k_unpack_tiles<KeyT><<<(unsigned)std::min<int64_t>(m_max / kSplitTile + src.n_seg, (int64_t)c->n_sms * 4), 512, 0, st>>>(
        (const uint64_t*)src.keys, c->tile_desc.as<uint4>(), c->tile_prefix.as<uint32_t>() + src.n_seg,
        pack_format(c, k, src.drop, src.owner_bits, (uint32_t)c->my_rank),
        c->fb_keys.as<KeyT>(), c->fb_vals.as<uint32_t>())));
    src_keys = c->fb_keys.as<KeyT>(); src_vals = c->fb_vals.as<uint32_t>();
  }
  const uint32_t* n_src_tiles = c->tile_prefix.as<uint32_t>() + src.n_seg;
  const int64_t src_tiles_max = m_max / kSplitTile + src.n_seg;

  for (int pass = 0; pass < n_passes; ++pass) {
    const PassSel sel{64 - kMaxLeafBits - pass_bits, (uint32_t)n_passes - 1u, (uint32_t)pass};
    tm.begin(NRP_T_HISTOGRAM);
    CU(cudaMemsetAsync(c->hist.ptr, 0, 4 * ((size_t)n_leaf + 2), st));
    if (m_max > 0) {
      if (from_ascii) {
        if (bits >= 15) {
          const unsigned grid = (unsigned)std::min<int64_t>((c->shard_tiles + 1) / 2, (int64_t)c->n_sms);
          NRP_LAUNCH((k_hist_from_ascii<2><<<grid, 2 * kTileThreads, 4u << bits, st>>>(P, c->flags.as<uint8_t>(), k, bits, 0, c->shard_tiles, sel, c->hist.as<uint32_t>())));
        } else {
          const unsigned grid = (unsigned)std::min<int64_t>(c->shard_tiles, (int64_t)c->n_sms * (bits >= 14 ? 2 : 3));
          NRP_LAUNCH((k_hist_from_ascii<1><<<grid, kTileThreads, 4u << bits, st>>>(P, c->flags.as<uint8_t>(), k, bits, 0, c->shard_tiles, sel, c->hist.as<uint32_t>())));
        }
      } else {
        // the source's tile list is rebuilt every pass: the level-2 split below reuses the descriptor buffer
        TRY(build_tiles(c, src.ext, src.n_seg, 0xffffffffu, 0, nullptr));
        const unsigned grid = (unsigned)std::min<int64_t>(src_tiles_max, (int64_t)c->n_sms * (bits >= 15 ? 1 : 2));
        NRP_LAUNCH((k_hist_from_tiles<KeyT><<<grid, 512, 4u << bits, st>>>(src_keys, c->tile_desc.as<uint4>(), n_src_tiles, bits, sel, c->hist.as<uint32_t>())));
      }
    }
    exclusive_scan<uint32_t, uint32_t>(c->hist.as<uint32_t>(), c->leaf_off.as<uint32_t>(), n_leaf, c->scan_scratch.as<uint32_t>(), st);
    NRP_LAUNCH((k_init_cursors<<<grid_for(n_leaf, 256), 256, 0, st>>>(c->leaf_off.as<uint32_t>(), r1, r2, c->cursor1.as<uint32_t>(), c->cursor2.as<uint32_t>())));
    int64_t n_pass = m_max;
    if (n_passes > 1) {
      // the slice's exact size decides the buffers; its candidates are appended to a dedicated list
      CU(cudaMemcpyAsync(c->h_counters + 16, c->leaf_off.as<uint32_t>() + n_leaf, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n_pass = (int64_t)(uint32_t)c->h_counters[16];
      TRY(ensure_record_buffers(c, std::max(n_pass, m_pass), sizeof(KeyT), src.n_seg));
      TRY(c->cand_keys_buf.grow_keep(sizeof(KeyT) * (size_t)(cand_so_far + n_pass + 16), sizeof(KeyT) * (size_t)cand_so_far));
      TRY(c->cand_vals_buf.grow_keep(4 * (size_t)(cand_so_far + n_pass + 16), 4 * (size_t)cand_so_far));
    }

    tm.begin(NRP_T_SPLIT1);
    leaf_keys = c->keysA.as<KeyT>(); leaf_vals = c->valsA.as<uint32_t>();
    free_keys = c->keysB.as<KeyT>(); free_vals = c->valsB.as<uint32_t>();
    uint32_t* level1_cursor = r2 > 0 ? c->cursor1.as<uint32_t>() : c->cursor2.as<uint32_t>();
    if (m_max > 0) {
      if (from_ascii) {
        NRP_LAUNCH((k_split_from_ascii<KeyT, 0, true><<<split_grid(c), kTileThreads, sizeof(SplitSmem<KeyT>) + sizeof(TileStage), st>>>(
            P, c->flags.as<uint8_t>(), k, r1, 0, c->shard_tiles, c->part_base, c->j_bits, no_pack, sel, level1_cursor, leaf_keys, leaf_vals, no_limit, PeerBuffers{}, 0)));
      } else {
        const unsigned grid = (unsigned)std::min<int64_t>(src_tiles_max, (int64_t)c->n_sms * 4);
        NRP_LAUNCH((k_split_from_records<KeyT, true><<<grid, kSplitThreads, sizeof(SplitSmem<KeyT>), st>>>(
            src_keys, src_vals, c->tile_desc.as<uint4>(), n_src_tiles, 0, r1, no_pack, sel, level1_cursor, leaf_keys, leaf_vals, no_limit)));
      }
    }
    if (r2 > 0 && m_max > 0) {
      tm.begin(NRP_T_SPLIT2);
      const uint32_t n_l1 = 1u << r1;
      TRY(build_tiles(c, SegExtents{c->leaf_off.as<uint32_t>(), r2, nullptr, 0u}, n_l1, 1u, r2, nullptr));
      const unsigned grid2 = (unsigned)std::min<int64_t>(n_pass / kSplitTile + n_l1, (int64_t)c->n_sms * 4);
      NRP_LAUNCH((k_split_from_records<KeyT, false><<<grid2, kSplitThreads, sizeof(SplitSmem<KeyT>), st>>>(
          leaf_keys, leaf_vals, c->tile_desc.as<uint4>(), c->tile_prefix.as<uint32_t>() + n_l1, r1, r2, no_pack, PassSel{0, 0u, 0u}, c->cursor2.as<uint32_t>(),
          free_keys, free_vals, no_limit)));
      std::swap(leaf_keys, free_keys); std::swap(leaf_vals, free_vals);
    }
    tm.begin(NRP_T_FILTER);
    KeyT* out_keys = n_passes > 1 ? c->cand_keys_buf.as<KeyT>() : free_keys;
    uint32_t* out_vals = n_passes > 1 ? c->cand_vals_buf.as<uint32_t>() : free_vals;
    if (m_max > 0)
      TRY((launch_filter<KeyT>(c, false, no_pack, leaf_keys, leaf_vals, LeafExtents{c->leaf_off.as<uint32_t>(), nullptr, 0u}, n_leaf, bits,
                               (m_plan >> pass_bits) >> bits, out_keys, out_vals)));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_counters, cnt, sizeof(unsigned long long) * CNT_N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_counters + 16, c->leaf_off.as<uint32_t>() + n_leaf, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cand_so_far = (int64_t)c->h_counters[CNT_CAND];
    c->n_records += (int64_t)(uint32_t)c->h_counters[16];
  }
  c->n_cand = cand_so_far;
  c->counts[NRP_C_OVERSIZED] = (int64_t)(uint32_t)c->h_counters[CNT_OVERSIZED];
  c->counts[NRP_C_CANDIDATES] = c->n_cand;
  c->counts[NRP_C_RUNS] = n_passes;
  c->cand_unsorted = (uint32_t)c->h_counters[CNT_UNSORTED] != 0;
  if (n_passes > 1) {
    // candidates of all passes live in the dedicated list; the ping-pong buffers serve as scratch of the same size
    TRY(c->keysA.ensure(sizeof(KeyT) * (size_t)(c->n_cand + 16))); TRY(c->valsA.ensure(4 * (size_t)(c->n_cand + 16)));
    c->cand_keys = c->cand_keys_buf.ptr; c->cand_vals = c->cand_vals_buf.as<uint32_t>();
    c->spare_keys = c->keysA.ptr; c->spare_vals = c->valsA.as<uint32_t>();
  } else {
    c->cand_keys = free_keys; c->cand_vals = free_vals;         // candidates; the leaf buffers are free again
    c->spare_keys = leaf_keys; c->spare_vals = leaf_vals;
  }
  return NRP_OK;
}

template <typename KeyT>
int stage_bulk(nrp_ctx* c, int k, const BulkSource<KeyT>& src, StageTimer& tm) {
  if (c->opt_direct) {
    bool overflowed = false;
    TRY((stage_bulk_direct<KeyT>(c, k, src, tm, &overflowed)));
    if (!overflowed) return NRP_OK;
    c->counts[NRP_C_FALLBACKS] += 1;
  }
  return stage_bulk_exact<KeyT>(c, k, src, tm);
}

// a flat record array as a one-segment source (the two offsets live in the context's seg_off buffer)
template <typename KeyT>
int flat_source(nrp_ctx* c, const KeyT* keys, const uint32_t* vals, int64_t n, BulkSource<KeyT>* out) {
  TRY(c->seg_off.ensure(4 * (size_t)(kLevelBuckets * kMaxPeers + 4)));
  uint32_t* h = reinterpret_cast<uint32_t*>(c->h_counters + 40);     // pinned: no need to wait for the copy
  h[0] = 0u; h[1] = (uint32_t)n;
  CU(cudaMemcpyAsync(c->seg_off.ptr, h, 8, cudaMemcpyHostToDevice, c->stream));
  out->from_parts = false; out->keys = keys; out->vals = vals;
  out->ext = SegExtents{c->seg_off.as<uint32_t>(), 0, nullptr, 0u};
  out->n_seg = 1; out->segs_per_bucket = 1; out->r_done = 0; out->m_upper = n;
  return NRP_OK;
}

int ensure_group_buffers(nrp_ctx* c, int64_t n) {
  const size_t n1 = (size_t)n + 2;
  TRY(c->head.ensure(4 * n1)); TRY(c->head_scan.ensure(4 * n1)); TRY(c->group_start.ensure(4 * n1));
  TRY(c->valid.ensure(4 * n1)); TRY(c->valid_scan.ensure(4 * n1));
  TRY(c->mem_cnt.ensure(4 * n1)); TRY(c->mem_scan.ensure(8 * n1));
  TRY(c->pair_cnt.ensure(8 * n1)); TRY(c->pair_scan.ensure(8 * n1));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems((int64_t)n1 + 1)));
  return NRP_OK;
}

// ---- stage: order the candidates by (k-mer, part) if the filter could not, mark key-run heads, detect internal repeats
// (two equal adjacent records).  With mark_repeats the parts get NRP_FLAG_REPEAT in flags[value].  No host round trip.
template <typename KeyT>
int stage_sort_mark(nrp_ctx* c, int k, int64_t n_parts_total, bool mark_repeats, StageTimer& tm) {
  cudaStream_t st = c->stream;
  tm.begin(NRP_T_SORT);
  const int64_t n = c->n_cand;
  int part_bits = 1;
  while (((int64_t)1 << part_bits) < n_parts_total) ++part_bits;
  if (c->cand_unsorted) {
    // rare: a leaf bucket over the filter capacity or with more suspects than the in-leaf ordering takes (heavy duplicate keys)
    TRY(c->lsd_scratch.ensure(sizeof(uint32_t) * (size_t)lsd_scratch_elems(n)));
    LsdPlan plan = lsd_plan(c->generic ? 64 : 2 * k, part_bits + c->j_bits);
    KeyT* sk; uint32_t* sv;
    lsd_sort<KeyT, true>((KeyT*)c->cand_keys, c->cand_vals, (KeyT*)c->spare_keys, c->spare_vals, n, plan, c->lsd_scratch.as<uint32_t>(), st, &sk, &sv);
    if (sk != (KeyT*)c->cand_keys) { std::swap(c->cand_keys, c->spare_keys); std::swap(c->cand_vals, c->spare_vals); }
  }
  tm.begin(NRP_T_GROUP);
  TRY(ensure_group_buffers(c, n));
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  c->n_sorted = n;
  c->n_repeat_records = 0;
  if (n > 0)
    NRP_LAUNCH((k_group_heads<KeyT><<<grid_for(n, 256), 256, 0, st>>>((KeyT*)c->cand_keys, c->cand_vals, n, c->j_bits, c->head.as<uint32_t>(),
                                                                  mark_repeats ? c->flags.as<uint8_t>() : nullptr, cnt + CNT_REPEAT)));
  CU(cudaGetLastError());
  return NRP_OK;
}

// ---- stage: (optionally) drop every record of a part that carries the repeat flag, then key runs with >= 2 records ->
// groups CSR (keys, indptr, members, pair offsets).  Everything is enqueued with device-side counts (n_max = number of
// candidates bounds every grid); the totals are copied to pinned memory and become valid at the caller's next synchronisation
// (read them with group_counts_ready).
template <typename KeyT>
int stage_groups(nrp_ctx* c, bool drop_flagged, StageTimer& tm) {
  cudaStream_t st = c->stream;
  tm.begin(NRP_T_GROUP);
  const int64_t n_max = c->n_cand;
  c->n_groups = c->n_members = c->n_pairs = 0;
  for (int i = 17; i <= 22; ++i) c->h_counters[i] = 0;
  const size_t ng_max = (size_t)n_max / 2 + 2;
  TRY(c->g_keys.ensure(8 * ng_max)); TRY(c->g_indptr.ensure(8 * (ng_max + 1)));
  TRY(c->g_pair_off.ensure(8 * (ng_max + 1))); TRY(c->g_members.ensure(4 * (size_t)(n_max + 1)));
  TRY(c->g_member_end.ensure(4 * (size_t)(n_max + 1)));
  TRY(c->g_first_window.ensure(4 * ng_max)); TRY(c->g_first_part.ensure(4 * ng_max));
  if (n_max == 0) {
    CU(cudaMemsetAsync(c->g_indptr.ptr, 0, 16, st));
    CU(cudaMemsetAsync(c->g_pair_off.ptr, 0, 16, st));
    return NRP_OK;
  }
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  if (c->opt_coop_group == 2 || (c->opt_coop_group == 1 && n_max <= (1ll << 20))) {
    // the whole stage as one cooperative kernel (grid-wide barriers instead of ~20 launches)
    const int64_t n_tiles = (n_max + kScanTile - 1) / kScanTile;
    TRY(c->pair_cnt.ensure(sizeof(GroupTotals) * (size_t)(n_tiles + 2)));
    TRY(c->mem_cnt.ensure(4 * (size_t)(n_tiles + 2)));
    int& per_sm = c->coop_blocks_per_sm[sizeof(KeyT) == 4 ? 0 : 1];
    if (per_sm == 0) {
      int per_sm_walk = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_group_coop<KeyT>, kScanThreads, 0));
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_walk, k_group_walk<KeyT>, kScanThreads, 0));
      per_sm = std::min(per_sm, per_sm_walk);          // one grid size for both cooperative variants
      if (per_sm < 1) return fail(NRP_E_CUDA, "cooperative grouping kernel does not fit an SM");
    }
    GroupCoopArgs a;
    a.keys = c->cand_keys; a.vals = c->cand_vals; a.keys2 = c->spare_keys; a.vals2 = c->spare_vals;
    a.n = n_max; a.j_bits = c->j_bits; a.drop_flagged = drop_flagged ? 1 : 0; a.flags = c->flags.as<uint8_t>();
    a.head = c->head.as<uint32_t>(); a.head_scan = c->head_scan.as<uint32_t>(); a.group_start = c->group_start.as<uint32_t>();
    a.valid = c->valid.as<uint32_t>(); a.valid_scan = c->valid_scan.as<uint32_t>();
    a.mem_scan = c->mem_scan.as<unsigned long long>(); a.pair_scan = c->pair_scan.as<unsigned long long>();
    a.tile_u32 = c->mem_cnt.as<uint32_t>(); a.tile_tot = c->pair_cnt.as<GroupTotals>();
    a.out_keys = c->g_keys.as<uint64_t>(); a.out_indptr = c->g_indptr.as<int64_t>(); a.out_pair_off = c->g_pair_off.as<unsigned long long>();
    a.out_first_window = c->g_first_window.as<uint32_t>(); a.out_members = c->g_members.as<uint32_t>();
    a.out_member_end = c->g_member_end.as<uint32_t>();
    a.counts = cnt + CNT_GROUP;
    const unsigned cgrid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(n_tiles, (int64_t)per_sm * c->n_sms));
    void* args[] = {&a};
    ++launch_counter(); ++launch_total();
    c->h_counters[55] = 0;
    if (c->opt_group_walk && !c->group_walk_gave_up) {
      // two grid barriers: the run heads walk their runs on the uncompacted list (nothing is swapped); a run longer than
      // kGroupWalkMax raises counts[5] and the caller repeats the stage with the nine-barrier kernel (group_stage_retry)
      CU(cudaMemsetAsync(cnt + CNT_GROUP, 0, 6 * 8, st));
      CU(cudaLaunchCooperativeKernel((void*)k_group_walk<KeyT>, dim3(cgrid), dim3(kScanThreads), args, 0, st));
      CU(cudaMemcpyAsync(c->h_counters + 55, cnt + CNT_GROUP + 5, 8, cudaMemcpyDeviceToHost, st));
      c->cand_compacted = false;
    } else {
      CU(cudaLaunchCooperativeKernel((void*)k_group_coop<KeyT>, dim3(cgrid), dim3(kScanThreads), args, 0, st));
      if (drop_flagged) { std::swap(c->cand_keys, c->spare_keys); std::swap(c->cand_vals, c->spare_vals); }
      c->cand_compacted = drop_flagged;
    }
    CU(cudaMemcpyAsync(c->h_counters + 18, cnt + CNT_REPEAT, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_counters + 50, cnt + CNT_GROUP, 5 * 8, cudaMemcpyDeviceToHost, st));
    c->group_counts_coop = true;
    return NRP_OK;
  }
  c->group_counts_coop = false;
  c->h_counters[55] = 0;
  c->cand_compacted = drop_flagged;
  uint32_t* n_sorted_dev = reinterpret_cast<uint32_t*>(cnt + CNT_NSORTED);
  const unsigned grid = grid_for(n_max + 1, 256);
  if (drop_flagged) {
    NRP_LAUNCH((k_keep_unflagged<<<grid, 256, 0, st>>>(c->cand_vals, n_max, c->j_bits, c->flags.as<uint8_t>(), c->valid.as<uint32_t>())));
    exclusive_scan<uint32_t, uint32_t>(c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(), n_max, c->scan_scratch.as<uint32_t>(), st);
    NRP_LAUNCH((k_compact_records<KeyT><<<grid, 256, 0, st>>>((KeyT*)c->cand_keys, c->cand_vals, c->valid.as<uint32_t>(),
                                                            c->valid_scan.as<uint32_t>(), n_max, (KeyT*)c->spare_keys, c->spare_vals)));
    CU(cudaMemcpyAsync(n_sorted_dev, c->valid_scan.as<uint32_t>() + n_max, 4, cudaMemcpyDeviceToDevice, st));
    std::swap(c->cand_keys, c->spare_keys); std::swap(c->cand_vals, c->spare_vals);
    NRP_LAUNCH((k_group_heads_n<KeyT><<<grid, 256, 0, st>>>((KeyT*)c->cand_keys, c->cand_vals, n_sorted_dev, n_max, c->head.as<uint32_t>())));
  } else {
    uint32_t* h = reinterpret_cast<uint32_t*>(c->h_counters + 46);       // pinned: the copy reads it later, nobody rewrites it before the next job
    h[0] = (uint32_t)n_max;
    CU(cudaMemcpyAsync(n_sorted_dev, h, 4, cudaMemcpyHostToDevice, st));
  }
  exclusive_scan<uint32_t, uint32_t>(c->head.as<uint32_t>(), c->head_scan.as<uint32_t>(), n_max, c->scan_scratch.as<uint32_t>(), st);
  const uint32_t* n_runs_dev = c->head_scan.as<uint32_t>() + n_max;
  NRP_LAUNCH((k_group_starts_n<<<grid, 256, 0, st>>>(c->head.as<uint32_t>(), c->head_scan.as<uint32_t>(), n_sorted_dev, n_max, c->group_start.as<uint32_t>())));
  {  // run sizes + the three prefix sums (validity, members, pairs) in three launches
    const int64_t n_tiles = (n_max + kScanTile - 1) / kScanTile;
    TRY(c->pair_cnt.ensure(sizeof(GroupTotals) * (size_t)(n_tiles + 1)));          // tile totals / offsets
    GroupTotals* totals = c->pair_cnt.as<GroupTotals>();
    NRP_LAUNCH((k_sizes_scan_tiles<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(c->group_start.as<uint32_t>(), n_runs_dev, n_max, c->valid.as<uint32_t>(),
                                                                             c->valid_scan.as<uint32_t>(), c->mem_scan.as<unsigned long long>(),
                                                                             c->pair_scan.as<unsigned long long>(), totals)));
    NRP_LAUNCH((k_sizes_scan_top<<<1, kScanThreads, 0, st>>>(totals, n_tiles, n_max, c->valid_scan.as<uint32_t>(), c->mem_scan.as<unsigned long long>(),
                                                            c->pair_scan.as<unsigned long long>())));
    NRP_LAUNCH((k_sizes_scan_add<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(totals, n_max, c->valid_scan.as<uint32_t>(),
                                                                           c->mem_scan.as<unsigned long long>(), c->pair_scan.as<unsigned long long>())));
  }
  NRP_LAUNCH((k_group_emit_n<KeyT><<<grid, 256, 0, st>>>(
      (KeyT*)c->cand_keys, c->cand_vals, c->j_bits, c->group_start.as<uint32_t>(), c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(),
      c->mem_scan.as<unsigned long long>(), c->pair_scan.as<unsigned long long>(), n_runs_dev, n_max, c->g_keys.as<uint64_t>(),
      c->g_indptr.as<int64_t>(), c->g_pair_off.as<unsigned long long>(), c->g_first_window.as<uint32_t>())));
  NRP_LAUNCH((k_group_members_n<<<grid, 256, 0, st>>>(c->cand_vals, c->head_scan.as<uint32_t>(), c->head.as<uint32_t>(),
                                                    c->group_start.as<uint32_t>(), c->valid.as<uint32_t>(),
                                                    c->mem_scan.as<unsigned long long>(), n_sorted_dev, c->j_bits, c->g_members.as<uint32_t>(),
                                                    c->g_member_end.as<uint32_t>())));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_counters + 18, cnt + CNT_REPEAT, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_counters + 19, n_sorted_dev, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_counters + 20, c->valid_scan.as<uint32_t>() + n_max, 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_counters + 21, c->mem_scan.as<unsigned long long>() + n_max, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(c->h_counters + 22, c->pair_scan.as<unsigned long long>() + n_max, 8, cudaMemcpyDeviceToHost, st));
  return NRP_OK;
}

// after a synchronisation that follows stage_groups: true when the walk variant gave up and the stage has to be repeated
bool group_stage_retry(nrp_ctx* c) {
  if (c->group_counts_coop && c->n_cand > 0 && c->h_counters[55] != 0) { c->group_walk_gave_up = true; return true; }
  return false;
}

// after a synchronisation that follows stage_groups
void group_counts_ready(nrp_ctx* c) {
  c->n_repeat_records = (int64_t)c->h_counters[18];
  if (c->group_counts_coop && c->n_cand > 0) {
    c->n_sorted = (int64_t)c->h_counters[50];
    c->n_groups = (int64_t)c->h_counters[52];
    c->n_members = (int64_t)c->h_counters[53];
    c->n_pairs = (int64_t)c->h_counters[54];
  } else {
    c->n_sorted = c->n_cand > 0 ? (int64_t)(uint32_t)c->h_counters[19] : 0;
    c->n_groups = (int64_t)(uint32_t)c->h_counters[20];
    c->n_members = (int64_t)c->h_counters[21];
    c->n_pairs = (int64_t)c->h_counters[22];
  }
  c->counts[NRP_C_GROUPS] = c->n_groups;
  c->counts[NRP_C_MEMBERS] = c->n_members;
  c->counts[NRP_C_PAIRS] = c->n_pairs;
}

// ---- stage: rank of every group = first window of its first member that carries the group's key
int stage_first_window(nrp_ctx* c, int k, StageTimer& tm) {
  cudaStream_t st = c->stream;
  tm.begin(NRP_T_ORDER);
  TRY(c->g_first_part.ensure(4 * (size_t)(c->n_groups + 1))); TRY(c->g_first_window.ensure(4 * (size_t)(c->n_groups + 1)));
  PartHomes homes{};
  if (c->staged_lo != 0 || c->staged_hi != c->n_parts) {
    if (c->homes.n == 0 && c->n_groups > 0 && c->j_bits == 0)
      return fail(NRP_E_STATE, "shard-only staging needs the peers' part buffers (nrp_peer_open_parts) to rank groups without a carried window");
    homes = c->homes;
  }
  if (c->n_groups > 0 && c->j_bits == 0) {
    NRP_LAUNCH((k_gather_first_part<<<grid_for(c->n_groups, 256), 256, 0, st>>>(c->g_indptr.as<int64_t>(), c->g_members.as<uint32_t>(),
                                                                            c->n_groups, c->g_first_part.as<uint32_t>())));
    NRP_LAUNCH((k_group_first_window<<<grid_for(c->n_groups, 128), 128, 0, st>>>(c->ascii.as<uint8_t>(), c->off.as<int64_t>(),
                                                                             c->g_keys.as<uint64_t>(), c->g_first_part.as<uint32_t>(),
                                                                             c->n_parts, c->n_groups, k, c->g_first_window.as<uint32_t>(), homes)));
  }
  CU(cudaGetLastError());
  return NRP_OK;
}

// ---- stage: last unshared window of the parts in [lo, hi); the shared k-mers (device, any order) go into a hash set
int stage_last_single(nrp_ctx* c, int k, const uint64_t* shared_keys, int64_t n_shared, int64_t lo, int64_t hi, StageTimer& tm,
                      cudaStream_t side_stream = nullptr) {
  cudaStream_t st = side_stream ? side_stream : c->stream;
  if (!side_stream) tm.begin(NRP_T_ORDER);
  int log_slots = 0;
  if (n_shared > 0) {
    log_slots = 6;
    while (((int64_t)1 << log_slots) < 2 * n_shared) ++log_slots;
    TRY(c->shared_keysB.ensure((size_t)8 << log_slots));
    CU(cudaMemsetAsync(c->shared_keysB.ptr, 0xff, (size_t)8 << log_slots, st));
    const unsigned grid = (unsigned)std::min<int64_t>((n_shared + 255) / 256, c->n_sms * 8);
    NRP_LAUNCH((k_bg_insert<uint64_t><<<grid, 256, 0, st>>>(shared_keys, n_shared, c->shared_keysB.as<uint64_t>(), log_slots)));
  }
  if (c->n_parts > 0) CU(cudaMemsetAsync(c->last_single.ptr, 0xff, sizeof(int32_t) * (size_t)c->n_parts, st));
  if (hi > lo && c->generic)
    NRP_LAUNCH((k_generic_last_single<<<grid_for(c->n_parts, 128), 128, 0, st>>>(c->ascii.as<uint8_t>(), c->off.as<int64_t>(), c->flags.as<uint8_t>(),
                                                                             c->n_parts, k, c->gen_seed, c->shared_keysB.as<uint64_t>(), log_slots,
                                                                             c->last_single.as<int32_t>())));
  else if (hi > lo)
    NRP_LAUNCH((k_last_single<<<grid_for(hi - lo, 128), 128, 0, st>>>(c->ascii.as<uint8_t>(), c->off.as<int64_t>(), c->flags.as<uint8_t>(), lo, hi,
                                                                  k, c->shared_keysB.as<uint64_t>(), log_slots, c->last_single.as<int32_t>())));
  CU(cudaGetLastError());
  return NRP_OK;
}

// hash set for at most n_codes edge codes in e_codesA, cleared; output lists e_u / e_v sized for n_codes
int edge_set_prepare(nrp_ctx* c, int64_t n_codes, int* log_slots) {
  int ls = 6;
  while (((int64_t)1 << ls) < n_codes + n_codes / 2) ++ls;      // load <= 2/3 (config 3: 4.2e6 pairs in 64 MB, which the L2 holds, not 128)
  TRY(c->e_codesA.ensure((size_t)8 << ls));
  TRY(c->e_u.ensure(4 * (size_t)(n_codes + 1))); TRY(c->e_v.ensure(4 * (size_t)(n_codes + 1)));
  CU(cudaMemsetAsync(c->e_codesA.ptr, 0xff, (size_t)8 << ls, c->stream));
  CU(cudaMemsetAsync(c->counters.as<unsigned long long>() + CNT_EDGES, 0, 8, c->stream));
  *log_slots = ls;
  return NRP_OK;
}

int edge_set_count(nrp_ctx* c) {
  CU(cudaMemcpyAsync(c->h_counters + 24, c->counters.as<unsigned long long>() + CNT_EDGES, 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->counts[NRP_C_EDGES] = (int64_t)c->h_counters[24];
  return NRP_OK;
}

// ---- stage: distinct part-pair edges of this context's groups (order unspecified)
int stage_edges(nrp_ctx* c, StageTimer& tm) {
  cudaStream_t st = c->stream;
  tm.begin(NRP_T_EDGES);
  if (!c->have_ids) return fail(NRP_E_ARG, "want_edges needs part ids");
  if (c->n_pairs > c->opt_max_pairs)
    return fail(NRP_E_EDGES, "%lld member pairs exceed the limit of %lld", (long long)c->n_pairs, (long long)c->opt_max_pairs);
  c->counts[NRP_C_EDGES] = 0;
  if (c->n_pairs > 0) {
    int log_slots = 0;
    TRY(edge_set_prepare(c, c->n_pairs, &log_slots));
    NRP_LAUNCH((k_edge_pairs<<<grid_for(c->n_members, 128), 128, 0, st>>>(c->g_member_end.as<uint32_t>(), c->g_members.as<uint32_t>(),
                                                                      c->n_members, c->part_id.as<uint32_t>(), c->e_codesA.as<uint64_t>(),
                                                                      log_slots, c->e_u.as<uint32_t>(), c->e_v.as<uint32_t>(),
                                                                      c->counters.as<unsigned long long>() + CNT_EDGES)));
    CU(cudaGetLastError());
    TRY(edge_set_count(c));
  }
  return NRP_OK;
}

// flagged parts and the windows of the parts flagged for internal repeats (they were counted as records before the flag was
// known): enqueued; finish_counts_ready reads the numbers after the next synchronisation
int finish_counts_enqueue(nrp_ctx* c, StageTimer& tm) {
  cudaStream_t st = c->stream;
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  tm.begin(NRP_T_GROUP);
  CU(cudaMemsetAsync(cnt + CNT_FLAGGED, 0, 16, st));
  if (c->n_parts > 0) {
    NRP_LAUNCH((k_count_flagged<<<grid_for(c->n_parts, 256), 256, 0, st>>>(c->flags.as<uint8_t>(), c->n_parts, cnt + CNT_FLAGGED)));
    if (!c->generic)   // (the general path never emits the records of a part with an internal repeat: nothing to subtract)
      NRP_LAUNCH((k_flagged_windows<<<grid_for(c->shard_hi - c->shard_lo, 256), 256, 0, st>>>(c->off.as<int64_t>(), c->flags.as<uint8_t>(),
                                                                                          c->shard_lo, c->shard_hi, c->k, cnt + CNT_WORK)));
  }
  CU(cudaMemcpyAsync(c->h_counters + 26, cnt + CNT_FLAGGED, 16, cudaMemcpyDeviceToHost, st));
  CU(cudaGetLastError());
  return NRP_OK;
}

void finish_counts_ready(nrp_ctx* c) {
  c->counts[NRP_C_EXCLUDED] = (int64_t)c->h_counters[26];
  c->counts[NRP_C_RECORDS] = c->n_records - (int64_t)c->h_counters[27];
}

int finish_timing(nrp_ctx* c, StageTimer& tm) {
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  c->counts[NRP_C_LAUNCHES] = launch_counter();
  double total = 0;
  for (int i = NRP_T_FLAGS; i < NRP_N_TIMINGS; ++i) total += c->timings[i];
  c->timings[NRP_T_TOTAL] = total;
  return NRP_OK;
}

// layout of a record of the staged parts at window length k
void record_layout(nrp_ctx* c, int k, bool no_window = false) {
  // carry the window index in the record value when (part, window) fits 32 bits
  int part_bits = 1, win_bits = 1;
  while (((int64_t)1 << part_bits) < c->n_parts) ++part_bits;
  const int64_t max_windows = c->max_len - k + 1;
  while (((int64_t)1 << win_bits) < max_windows) ++win_bits;
  c->j_bits = (max_windows >= 1 && part_bits + win_bits <= 32) ? win_bits : 0;
  if (c->opt_no_window_carry || no_window) c->j_bits = 0;
  c->vbits = std::min(32, part_bits + c->j_bits);
}

void begin_job(nrp_ctx* c, int k, size_t key_bytes, int want_edges, bool no_window = false) {
  if (c->early_valid) { cudaStreamSynchronize(c->fetch_stream); c->early_valid = false; }     // an early fetch nobody collected: its sources are about to change
  memset(c->counts, 0, sizeof(c->counts));
  memset(c->timings, 0, sizeof(c->timings));
  launch_counter() = 0;
  c->counts[NRP_C_KEY_BYTES] = (int64_t)key_bytes;
  c->counts[NRP_C_EDGES] = want_edges ? 0 : -1;
  c->counts[NRP_C_RUNS] = 1;
  c->k = k; c->key_bytes = (int)key_bytes;
  c->have_result = false;
  record_layout(c, k, no_window);
}

// Layout of the page-locked result buffer (nrp_result_host).  The edge arrays come last, so that every other offset is known as
// soon as the grouping stage has reported its sizes.
struct ResultLayout {
  size_t indptr, keys, members, fw, last, flags, eu, ev, total;
  ResultLayout(int64_t n, int64_t ng, int64_t nm, int64_t ne) {
    auto up = [](size_t b) { return (b + 63) & ~(size_t)63; };
    indptr = 0; keys = indptr + up(8 * (size_t)(ng + 1)); members = keys + up(8 * (size_t)ng); fw = members + up(4 * (size_t)nm);
    last = fw + up(4 * (size_t)ng); flags = last + up(4 * (size_t)n); eu = flags + up((size_t)n); ev = eu + up(4 * (size_t)ne);
    total = ev + up(4 * (size_t)ne) + 64;
  }
};

// Device-to-host copy by a few CTAs storing straight into the page-locked result buffer (mapped into the device's address space).
// Not cudaMemcpyAsync: the copy engine that would move these 12 MB is also the one the rest of the build needs for its memsets and
// its small count read-backs, and they queue behind the bulk copies (measured, run R: the build grew by the copies' 0.23 ms).
struct FetchSegments {
  const unsigned char* src[5];
  unsigned char* dst[5];
  size_t bytes[5];
};
__global__ void __launch_bounds__(256) k_fetch_to_host(FetchSegments f) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
#pragma unroll 1
  for (int sgm = 0; sgm < 5; ++sgm) {
    const uint4* src = reinterpret_cast<const uint4*>(f.src[sgm]);            // device buffers and result offsets are 16-byte aligned
    uint4* dst = reinterpret_cast<uint4*>(f.dst[sgm]);
    const size_t n16 = f.bytes[sgm] >> 4;
    size_t i = tid;
    for (; i + 3 * nth < n16; i += 4 * nth) {                                 // four loads in flight per thread
      const uint4 a = src[i], b = src[i + nth], c4 = src[i + 2 * nth], d = src[i + 3 * nth];
      dst[i] = a; dst[i + nth] = b; dst[i + 2 * nth] = c4; dst[i + 3 * nth] = d;
    }
    for (; i < n16; i += nth) dst[i] = src[i];
    if (tid == 0)
      for (size_t b = n16 << 4; b < f.bytes[sgm]; ++b) f.dst[sgm][b] = f.src[sgm][b];
  }
}

// Start copying what the grouping stage has finished (the host knows the sizes: it has just synchronised).  The copy waits for the
// kernels already enqueued on the build's stream (the first-window pass, when the records do not carry the window) and runs on a
// stream of its own beside everything enqueued afterwards.
int early_fetch_enqueue(nrp_ctx* c) {
  c->early_valid = false;
  if (!c->opt_early_fetch || c->external_stream || c->n_parts <= 0) return NRP_OK;
  const int64_t n = c->n_parts, ng = c->n_groups, nm = c->n_members;
  const int64_t ne_bound = std::min<int64_t>(std::max<int64_t>(c->n_pairs, 0), (int64_t)1 << 24);     // more edges than this: nrp_result_host re-allocates and copies everything
  const ResultLayout L(n, ng, nm, ne_bound);
  TRY(c->h_results.ensure(L.total));
  unsigned char* hd = nullptr;                     // the result buffer as the device sees it
  if (cudaHostGetDevicePointer((void**)&hd, c->h_results.ptr, 0) != cudaSuccess || hd == nullptr) { cudaGetLastError(); return NRP_OK; }
  if (!c->fetch_stream) {
    CU(cudaStreamCreateWithFlags(&c->fetch_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->fetch_ready, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&c->fetch_done, cudaEventDisableTiming));
  }
  FetchSegments f;
  f.src[0] = c->g_members.as<unsigned char>(); f.dst[0] = hd + L.members; f.bytes[0] = 4 * (size_t)nm;
  f.src[1] = c->g_indptr.as<unsigned char>(); f.dst[1] = hd + L.indptr; f.bytes[1] = 8 * (size_t)(ng + 1);
  f.src[2] = c->g_first_window.as<unsigned char>(); f.dst[2] = hd + L.fw; f.bytes[2] = 4 * (size_t)ng;
  f.src[3] = c->flags.as<unsigned char>(); f.dst[3] = hd + L.flags; f.bytes[3] = (size_t)n;
  f.src[4] = c->g_keys.as<unsigned char>(); f.dst[4] = hd + L.keys; f.bytes[4] = 8 * (size_t)ng;
  cudaStream_t fs = c->fetch_stream;
  CU(cudaEventRecord(c->fetch_ready, c->stream));
  CU(cudaStreamWaitEvent(fs, c->fetch_ready, 0));
  NRP_LAUNCH((k_fetch_to_host<<<32, 256, 0, fs>>>(f)));
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->fetch_done, fs));
  c->early_valid = true; c->early_n = n; c->early_ng = ng; c->early_nm = nm;
  return NRP_OK;
}

template <typename KeyT>
int build_impl(nrp_ctx* c, int k, int internal_repeats, int want_edges) {
  TRY(configure_kernels<KeyT>(c));
  c->generic = false;
  c->shard_lo = 0; c->shard_hi = c->n_parts; c->shard_tile_lo = 0; c->shard_tiles = c->n_tiles; c->n_owners = 0;
  c->shard_m_max = windows_in_range(c, 0, c->n_parts, k);
  if (c->shard_m_max >= 0xffffe000ll)
    return fail(NRP_E_UNSUPPORTED, "%lld records exceed the 32-bit record index of one device", (long long)c->shard_m_max);
  begin_job(c, k, sizeof(KeyT), want_edges);
  c->counts[NRP_C_WINDOWS_MAX] = c->shard_m_max;
  StageTimer tm(c);
  TRY(stage_flags(c, k, internal_repeats, tm, true));
  TRY((stage_bulk<KeyT>(c, k, BulkSource<KeyT>{}, tm)));
  TRY((stage_sort_mark<KeyT>(c, k, c->n_parts, !internal_repeats, tm)));
  TRY((stage_groups<KeyT>(c, !internal_repeats, tm)));
  TRY(finish_counts_enqueue(c, tm));
  CU(cudaStreamSynchronize(c->stream));               // the one host round trip of the grouping: all counts at once
  if (group_stage_retry(c)) {                         // a very long key run: the nine-barrier kernel
    TRY((stage_groups<KeyT>(c, !internal_repeats, tm)));
    CU(cudaStreamSynchronize(c->stream));
  }
  c->group_walk_gave_up = false;
  group_counts_ready(c);
  finish_counts_ready(c);
  TRY(stage_first_window(c, k, tm));
  TRY(early_fetch_enqueue(c));
  if (c->opt_overlap_tail && want_edges && c->n_pairs > 0 && !c->external_stream) {
    // The host has just synchronised: the device is idle.  The last-unshared-window pass (reads the parts, probes the shared
    // k-mers) and the edge expansion (walks the member lists) touch different buffers and neither fills the device: two streams.
    if (!c->aux_stream) {
      CU(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
      CU(cudaEventCreate(&c->aux_begin)); CU(cudaEventCreate(&c->aux_end));
    }
    CU(cudaEventRecord(c->aux_begin, c->aux_stream));
    TRY(stage_last_single(c, k, c->g_keys.as<uint64_t>(), c->n_groups, 0, c->n_parts, tm, c->aux_stream));
    CU(cudaEventRecord(c->aux_end, c->aux_stream));
    TRY(stage_edges(c, tm));
    CU(cudaStreamWaitEvent(c->stream, c->aux_end, 0));
    TRY(finish_timing(c, tm));
    float ms = 0.f;
    CU(cudaEventSynchronize(c->aux_end));
    CU(cudaEventElapsedTime(&ms, c->aux_begin, c->aux_end));
    c->timings[NRP_T_ORDER] += ms;                           // reported, but it ran beside the edge stage: not in the total
  } else {
    TRY(stage_last_single(c, k, c->g_keys.as<uint64_t>(), c->n_groups, 0, c->n_parts, tm));
    if (want_edges) TRY(stage_edges(c, tm));
    TRY(finish_timing(c, tm));
  }
  c->have_result = true;
  return NRP_OK;
}

// ---- general path (generic.cuh): any byte alphabet, any k.  *status receives the collision bits (non-zero: repeat with another seed).
int build_generic_impl(nrp_ctx* c, int k, int internal_repeats, int want_edges, uint64_t seed, const uint8_t* preset_flags, int* status) {
  TRY(configure_kernels<uint64_t>(c));
  cudaStream_t st = c->stream;
  *status = 0;
  c->generic = true; c->gen_seed = seed;
  c->shard_lo = 0; c->shard_hi = c->n_parts; c->shard_tile_lo = 0; c->shard_tiles = c->n_tiles; c->n_owners = 0;
  c->shard_m_max = windows_in_range(c, 0, c->n_parts, k);
  if (c->shard_m_max >= 0xffffe000ll)
    return fail(NRP_E_UNSUPPORTED, "%lld records exceed the 32-bit record index of one device", (long long)c->shard_m_max);
  begin_job(c, k, 8, want_edges);
  c->fold_pal = c->fold_bg = false;
  const int64_t max_windows = std::max<int64_t>(1, c->max_len - k + 1);
  if (max_windows > kGenMaxWindows)
    return fail(NRP_E_UNSUPPORTED, "the general path takes parts of at most %d windows (longest part: %lld)", kGenMaxWindows, (long long)max_windows);
  if (c->j_bits == 0 && c->shard_m_max > 0)
    return fail(NRP_E_UNSUPPORTED, "the general path needs (part, window) to fit 32 bits");
  c->counts[NRP_C_WINDOWS_MAX] = c->shard_m_max;
  StageTimer tm(c);
  tm.begin(NRP_T_FLAGS);
  TRY(wait_chunks(c));
  TRY(c->flags.ensure((size_t)c->n_parts + 16));
  TRY(c->last_single.ensure(sizeof(int32_t) * (size_t)(c->n_parts + 1)));
  CU(cudaMemsetAsync(c->flags.ptr, 0, (size_t)c->n_parts + 16, st));
  if (preset_flags && c->n_parts > 0) CU(cudaMemcpyAsync(c->flags.ptr, preset_flags, (size_t)c->n_parts, cudaMemcpyHostToDevice, st));
  c->flags_set = true;
  TRY(c->gen_keys.ensure(8 * (size_t)(c->shard_m_max + 16))); TRY(c->gen_vals.ensure(4 * (size_t)(c->shard_m_max + 16)));
  TRY(c->gen_status.ensure(16));
  CU(cudaMemsetAsync(c->gen_status.ptr, 0, 16, st));
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  CU(cudaMemsetAsync(cnt + CNT_WORK, 0, 8, st));
  tm.begin(NRP_T_SPLIT1);
  int64_t n_records = 0;
  if (c->shard_m_max > 0) {
    int slots = 64;
    while (slots < 2 * max_windows) slots <<= 1;
    GenArgs a;
    a.ascii = c->ascii.as<uint8_t>(); a.off = c->off.as<int64_t>(); a.n_parts = c->n_parts;
    a.k = k; a.seed = seed; a.internal_repeats = internal_repeats; a.j_bits = c->j_bits; a.part_base = c->part_base;
    const bool bg_on = c->bg_n > 0;
    a.bg_table = c->bg_table.ptr; a.bg_log_slots = c->bg_log_slots; a.bg_key_bytes = c->bg_key_bytes; a.bg_K = bg_on ? c->bg_K : 0;
    a.flags = c->flags.as<uint8_t>();
    a.out_keys = c->gen_keys.as<uint64_t>(); a.out_vals = c->gen_vals.as<uint32_t>(); a.n_out = cnt + CNT_WORK;
    a.status = c->gen_status.as<uint32_t>();
    a.max_windows = (int)max_windows; a.table_slots = slots;
    const size_t smem = 16 * (size_t)max_windows + 4 * (size_t)slots + (size_t)max_windows + 16;
    CU(cudaFuncSetAttribute(k_generic_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>(c->n_parts, (int64_t)c->n_sms * 16);
    NRP_LAUNCH((k_generic_extract<<<grid, kGenThreads, smem, st>>>(a)));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_counters + 60, cnt + CNT_WORK, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_counters + 61, c->gen_status.ptr, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    n_records = (int64_t)c->h_counters[60];
    if ((uint32_t)c->h_counters[61] != 0) { *status = (int)(uint32_t)c->h_counters[61]; tm.finish(c->timings); return NRP_OK; }
  }
  BulkSource<uint64_t> src;
  TRY((flat_source<uint64_t>(c, c->gen_keys.as<uint64_t>(), c->gen_vals.as<uint32_t>(), n_records, &src)));
  TRY((stage_bulk<uint64_t>(c, std::min(k, 32), src, tm)));
  TRY((stage_sort_mark<uint64_t>(c, std::min(k, 32), c->n_parts, false, tm)));
  if (c->n_cand > 1) {
    NRP_LAUNCH((k_generic_verify<<<grid_for(c->n_cand, 256), 256, 0, st>>>(c->ascii.as<uint8_t>(), c->off.as<int64_t>(), (const uint64_t*)c->cand_keys,
                                                                      c->cand_vals, c->n_cand, k, c->j_bits, c->part_base, c->gen_status.as<uint32_t>())));
    CU(cudaMemcpyAsync(c->h_counters + 61, c->gen_status.ptr, 4, cudaMemcpyDeviceToHost, st));
  }
  TRY((stage_groups<uint64_t>(c, false, tm)));
  TRY(finish_counts_enqueue(c, tm));
  CU(cudaStreamSynchronize(st));
  if (c->n_cand > 1 && (uint32_t)c->h_counters[61] != 0) { *status = (int)(uint32_t)c->h_counters[61]; tm.finish(c->timings); return NRP_OK; }
  if (group_stage_retry(c)) {
    TRY((stage_groups<uint64_t>(c, false, tm)));
    CU(cudaStreamSynchronize(st));
  }
  c->group_walk_gave_up = false;
  group_counts_ready(c);
  finish_counts_ready(c);
  c->counts[NRP_C_RECORDS] = n_records;
  TRY(stage_first_window(c, k, tm));
  TRY(stage_last_single(c, k, c->g_keys.as<uint64_t>(), c->n_groups, 0, c->n_parts, tm));
  if (want_edges) TRY(stage_edges(c, tm));
  TRY(finish_timing(c, tm));
  c->have_result = true;
  return NRP_OK;
}

// ---- sharded path -------------------------------------------------------------------------------------
template <typename KeyT>
int shard_extract_impl(nrp_ctx* c, int64_t lo, int64_t hi, int k, int internal_repeats, int n_owners, int64_t* owner_counts) {
  TRY(configure_kernels<KeyT>(c));
  cudaStream_t st = c->stream;
  c->shard_lo = lo; c->shard_hi = hi; c->n_owners = n_owners;
  TRY(wait_chunks(c));                                   // a streamed batch must have landed completely
  const int64_t byte_lo = part_offset(c, lo), byte_hi = part_offset(c, hi);
  c->shard_tile_lo = byte_lo / kTileBytes;
  c->shard_tiles = hi > lo && byte_hi > byte_lo ? (byte_hi + kTileBytes - 1) / kTileBytes - c->shard_tile_lo : 0;
  c->shard_m_max = windows_in_range(c, lo, hi, k);
  if (c->shard_m_max >= 0xffffe000ll) return fail(NRP_E_UNSUPPORTED, "shard too large for 32-bit record indices");
  begin_job(c, k, sizeof(KeyT), 0);
  c->counts[NRP_C_WINDOWS_MAX] = c->shard_m_max;
  StageTimer tm(c);
  TRY(stage_flags(c, k, internal_repeats, tm));
  TRY(ensure_record_buffers(c, c->shard_m_max, sizeof(KeyT)));
  const Parts P = parts_view(c);
  tm.begin(NRP_T_HISTOGRAM);
  CU(cudaMemsetAsync(c->hist.ptr, 0, 4 * 260, st));
  if (c->shard_m_max > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(c->shard_tiles, (int64_t)c->n_sms * 2);
    NRP_LAUNCH((k_hist_from_ascii<1><<<grid, kTileThreads, 4u * 256, st>>>(P, c->flags.as<uint8_t>(), k, 8, n_owners, c->shard_tiles, PassSel{0, 0u, 0u}, c->hist.as<uint32_t>())));
  }
  exclusive_scan<uint32_t, uint32_t>(c->hist.as<uint32_t>(), c->cursor1.as<uint32_t>(), n_owners, c->scan_scratch.as<uint32_t>(), st);
  CU(cudaMemcpyAsync(c->h_counters + 32, c->hist.ptr, 4 * (size_t)n_owners, cudaMemcpyDeviceToHost, st));
  tm.begin(NRP_T_SPLIT1);
  if (c->shard_m_max > 0)
    NRP_LAUNCH((k_split_from_ascii<KeyT, 1, false><<<split_grid(c), kTileThreads, sizeof(SplitSmem<KeyT>) + sizeof(TileStage), st>>>(
        P, c->flags.as<uint8_t>(), k, 8, n_owners, c->shard_tiles, 0u, c->j_bits, pack_format(c, k, 0), PassSel{0, 0u, 0u}, c->cursor1.as<uint32_t>(), c->keysA.as<KeyT>(), c->valsA.as<uint32_t>(),
        RegionLimit{0u, 0u, nullptr, nullptr}, PeerBuffers{}, 0)));
  CU(cudaGetLastError());
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(st));
  const uint32_t* h = reinterpret_cast<const uint32_t*>(c->h_counters + 32);
  int64_t total = 0;
  for (int i = 0; i < n_owners; ++i) { owner_counts[i] = h[i]; total += h[i]; }
  c->n_records = total;
  return NRP_OK;
}

template <typename KeyT>
int shard_group_impl(nrp_ctx* c, const void* keys, const uint32_t* vals, int64_t n, int k, int internal_repeats) {
  TRY(configure_kernels<KeyT>(c));
  if (n >= 0xffffe000ll) return fail(NRP_E_UNSUPPORTED, "too many records for 32-bit record indices");
  // the received buffers are borrowed: they are consumed by the first split only
  StageTimer tm(c);
  BulkSource<KeyT> src;
  TRY((flat_source<KeyT>(c, (const KeyT*)keys, vals, n, &src)));
  TRY((stage_bulk<KeyT>(c, k, src, tm)));
  TRY((stage_sort_mark<KeyT>(c, k, c->n_parts, !internal_repeats, tm)));
  tm.finish(c->timings);
  return NRP_OK;
}

template <typename KeyT>
int shard_finish_impl(nrp_ctx* c, int k, int internal_repeats, int want_edges) {
  StageTimer tm(c);
  TRY((stage_groups<KeyT>(c, !internal_repeats, tm)));     // flags may come from other ranks: always filter
  CU(cudaStreamSynchronize(c->stream));
  if (group_stage_retry(c)) {
    TRY((stage_groups<KeyT>(c, !internal_repeats, tm)));
    CU(cudaStreamSynchronize(c->stream));
  }
  c->group_walk_gave_up = false;
  group_counts_ready(c);
  TRY(stage_first_window(c, k, tm));
  if (want_edges) { c->counts[NRP_C_EDGES] = 0; TRY(stage_edges(c, tm)); }
  tm.finish(c->timings);
  if (c->cand_compacted) c->n_cand = c->n_sorted;   // the list now holds the kept records only (a repeated finish starts from them)
  c->have_result = true;
  return NRP_OK;
}

// fused exchange, step 1: flags + owner histogram of the shard (counts go to the host for the offset exchange)
template <typename KeyT>
int shard_count_impl(nrp_ctx* c, int64_t lo, int64_t hi, int k, int internal_repeats, int n_owners, int64_t* owner_counts) {
  TRY(configure_kernels<KeyT>(c));
  cudaStream_t st = c->stream;
  c->shard_lo = lo; c->shard_hi = hi; c->n_owners = n_owners;
  TRY(wait_chunks(c));                                   // a streamed batch must have landed completely
  const int64_t byte_lo = part_offset(c, lo), byte_hi = part_offset(c, hi);
  c->shard_tile_lo = byte_lo / kTileBytes;
  c->shard_tiles = hi > lo && byte_hi > byte_lo ? (byte_hi + kTileBytes - 1) / kTileBytes - c->shard_tile_lo : 0;
  c->shard_m_max = windows_in_range(c, lo, hi, k);
  if (c->shard_m_max >= 0xffffe000ll) return fail(NRP_E_UNSUPPORTED, "shard too large for 32-bit record indices");
  begin_job(c, k, sizeof(KeyT), 0);
  c->counts[NRP_C_WINDOWS_MAX] = c->shard_m_max;
  StageTimer tm(c);
  TRY(stage_flags(c, k, internal_repeats, tm));
  TRY(c->hist.ensure(4 * (((size_t)1 << kMaxLeafBits) + 2)));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems(((int64_t)1 << kMaxLeafBits) + 2)));
  TRY(c->cursor1.ensure(4 * (size_t)(kLevelBuckets * kMaxPeers + 4)));
  const Parts P = parts_view(c);
  tm.begin(NRP_T_HISTOGRAM);
  CU(cudaMemsetAsync(c->hist.ptr, 0, 4 * 260, st));
  if (c->shard_m_max > 0) {
    const unsigned grid = (unsigned)std::min<int64_t>(c->shard_tiles, (int64_t)c->n_sms * 3);
    NRP_LAUNCH((k_hist_from_ascii<1><<<grid, kTileThreads, 4u * 256, st>>>(P, c->flags.as<uint8_t>(), k, 8, n_owners, c->shard_tiles, PassSel{0, 0u, 0u}, c->hist.as<uint32_t>())));
  }
  CU(cudaMemcpyAsync(c->h_counters + 32, c->hist.ptr, 4 * (size_t)n_owners, cudaMemcpyDeviceToHost, st));
  CU(cudaGetLastError());
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(st));
  const uint32_t* h = reinterpret_cast<const uint32_t*>(c->h_counters + 32);
  for (int i = 0; i < n_owners; ++i) owner_counts[i] = h[i];
  return NRP_OK;
}

// fused exchange, step 2: extraction + scatter straight into the owners' receive buffers (peer stores over NVLink).
// write_offsets[o] = record index in owner o's buffer where this rank's run starts.
template <typename KeyT>
int shard_scatter_impl(nrp_ctx* c, int k, const int64_t* write_offsets) {
  cudaStream_t st = c->stream;
  const Parts P = parts_view(c);
  uint32_t h_cursor[kMaxPeers];
  for (int o = 0; o < c->n_owners; ++o) h_cursor[o] = (uint32_t)write_offsets[o];
  CU(cudaMemcpyAsync(c->cursor1.ptr, h_cursor, 4 * (size_t)c->n_owners, cudaMemcpyHostToDevice, st));
  StageTimer tm(c);
  tm.begin(NRP_T_SPLIT1);
  if (c->shard_m_max > 0)
    NRP_LAUNCH((k_split_from_ascii<KeyT, 1, false><<<split_grid(c), kTileThreads, sizeof(SplitSmem<KeyT>) + sizeof(TileStage), st>>>(
        P, c->flags.as<uint8_t>(), k, 8, c->n_owners, c->shard_tiles, 0u, c->j_bits, pack_format(c, k, 0), PassSel{0, 0u, 0u}, c->cursor1.as<uint32_t>(), nullptr, nullptr,
        RegionLimit{0u, 0u, nullptr, nullptr}, c->peers, 1)));
  CU(cudaGetLastError());
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(st));
  return NRP_OK;
}

// ---- direct fused exchange ----------------------------------------------------------------------------------
// cursors and region ends of the sender: digit d = (owner << r1) | bucket writes into region (bucket * world + rank) of its owner
__global__ void k_init_owner_cursors(uint32_t* cursor, uint32_t* ends, uint32_t n_digits, int r1, uint32_t world, uint32_t rank, uint32_t cap) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_digits) return;
  const uint32_t b = d & ((1u << r1) - 1u);
  const uint32_t region = b * world + rank;
  cursor[d] = region * cap;
  ends[d] = (region + 1u) * cap;
}

// deliver the final cursors to the owners: ends of region (bucket * world + rank) in owner o's table (peer stores)
struct PeerEnds { uint32_t* ptr[kMaxPeers]; };
__global__ void k_push_region_ends(const uint32_t* cursor, uint32_t n_digits, int r1, uint32_t world, uint32_t rank, uint32_t cap, PeerEnds dst,
                                   unsigned long long* n_sent) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_digits) return;
  const uint32_t b = d & ((1u << r1) - 1u), o = d >> r1;
  const uint32_t region = b * world + rank;
  uint32_t e = cursor[d];
  if (e > (region + 1u) * cap) e = (region + 1u) * cap;
  dst.ptr[o][region] = e;
  atomicAdd(n_sent, (unsigned long long)(e - region * cap));
}

// level-1 bits of the direct sharded path: world * 2^r1 digits must fit one split level
int direct_shard_plan(nrp_ctx* c, int64_t windows_total, int world, int* r1_out, int* bits_out, uint32_t* cap1s_out, int64_t* m_owner_out) {
  const int64_t m_owner = windows_total / world + (windows_total / world) / 32 + 4096;
  // Runs stored over NVLink must stay long: at most opt_nvlink_digits (owner, bucket) digits per 4096-record tile.  If the
  // owner can finish with ONE more level of <= 9 bits -- with small leaves, or else with leaves for the large filter variant
  // -- the senders pre-partition by the missing bits; otherwise they scatter by owner only (or by the few bits two owner
  // levels cannot cover) and the owner runs two levels from the received regions.
  const int lv = (int)c->opt_max_level_bits;
  int r1_eff = 0;
  while (((int64_t)world << (r1_eff + 1)) <= c->opt_nvlink_digits) ++r1_eff;
  const int need = leaf_bits_for(c, m_owner, 1, kMaxDirectBits);
  int need_large = need;
  if (c->opt_leaf_target == kDefaultLeafTarget)
    while (need_large > 1 && (m_owner >> (need_large - 1)) <= kLargeLeafTarget) --need_large;
  int r1, bits;
  if (need - lv <= r1_eff) { bits = need; r1 = std::max(need - lv, 0); }
  else if (need_large - lv <= r1_eff) { bits = need_large; r1 = std::max(need_large - lv, 0); }
  else { bits = need; r1 = std::max(need - 2 * lv, 0); }
  while (r1 > 0 && (world << r1) > kLevelBuckets) --r1;
  if (world > kLevelBuckets) return fail(NRP_E_UNSUPPORTED, "%d owners do not fit one split level", world);
  *r1_out = r1;
  *bits_out = std::max(bits, r1 + 1);
  *cap1s_out = region_cap((double)windows_total / ((double)world * world * (double)(1 << r1)), 256);
  *m_owner_out = m_owner;
  return NRP_OK;
}

template <typename KeyT>
int shard_scatter_direct_impl(nrp_ctx* c, int64_t lo, int64_t hi, int k, int internal_repeats, int n_owners, int* overflow_out) {
  TRY(configure_kernels<KeyT>(c));
  cudaStream_t st = c->stream;
  c->shard_lo = lo; c->shard_hi = hi; c->n_owners = n_owners;
  TRY(wait_chunks(c));                                   // a streamed batch must have landed completely
  const int64_t byte_lo = part_offset(c, lo), byte_hi = part_offset(c, hi);
  c->shard_tile_lo = byte_lo / kTileBytes;
  c->shard_tiles = hi > lo && byte_hi > byte_lo ? (byte_hi + kTileBytes - 1) / kTileBytes - c->shard_tile_lo : 0;
  c->shard_m_max = windows_in_range(c, lo, hi, k);
  begin_job(c, k, sizeof(KeyT), 0, c->sh_no_window);
  c->counts[NRP_C_WINDOWS_MAX] = c->shard_m_max;
  StageTimer tm(c);
  TRY(stage_flags(c, k, internal_repeats, tm));
  TRY(c->cursor1.ensure(4 * (size_t)(kLevelBuckets * kMaxPeers + 4)));
  TRY(c->region_ends.ensure(4 * (size_t)(kLevelBuckets * kMaxPeers + 4)));
  const Parts P = parts_view(c);
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  tm.begin(NRP_T_SPLIT1);
  const uint32_t nd = (uint32_t)n_owners << c->sh_r1;
  CU(cudaMemsetAsync(cnt + CNT_SCATTER_OVERFLOW, 0, 16, st));
  NRP_LAUNCH((k_init_owner_cursors<<<grid_for(nd, 256), 256, 0, st>>>(c->cursor1.as<uint32_t>(), c->region_ends.as<uint32_t>(), nd, c->sh_r1,
                                                                    (uint32_t)n_owners, (uint32_t)c->my_rank, c->sh_cap1s)));
  if (packed_fits(c, k, c->sh_r1, c->sh_bits, c->sh_owner_bits) != c->sh_packed)
    return fail(NRP_E_STATE, "receive buffers were exported for another record layout");
  const RegionLimit lim{c->sh_cap1s, 0u, reinterpret_cast<uint32_t*>(cnt + CNT_SCATTER_OVERFLOW), c->region_ends.as<uint32_t>()};
  UniPlan up;
  if (c->shard_m_max > 0 && c->sh_packed && c->opt_uniform_kernel && c->uniform_len > 0 && uni_plan(c->uniform_len, k, &up)) {
    const unsigned ugrid = (unsigned)((hi - lo + up.G - 1) / up.G);
    const uint8_t* fl = c->flags_set ? c->flags.as<uint8_t>() : nullptr;
    const PackFmt pf = pack_format(c, k, c->sh_r1, c->sh_owner_bits, 0u);
    if (k > 16)
      NRP_LAUNCH((k_split_uniform<true, 2><<<ugrid, kSplitThreads, sizeof(UniSmem), st>>>(
          P.ascii, fl, up, lo, hi, k, c->sh_r1, n_owners, 0u, c->j_bits, pf, c->cursor1.as<uint32_t>(), nullptr, lim, c->peers_d)));
    else
      NRP_LAUNCH((k_split_uniform<false, 2><<<ugrid, kSplitThreads, sizeof(UniSmem), st>>>(
          P.ascii, fl, up, lo, hi, k, c->sh_r1, n_owners, 0u, c->j_bits, pf, c->cursor1.as<uint32_t>(), nullptr, lim, c->peers_d)));
    c->counts[NRP_C_UNIFORM] = 1;
  } else if (c->shard_m_max > 0) {
    if (c->sh_packed)
      NRP_LAUNCH((k_split_from_ascii<uint64_t, 2, false, true><<<split_grid(c), kTileThreads, sizeof(SplitSmem<uint64_t, true>) + sizeof(TileStage), st>>>(
          P, c->flags.as<uint8_t>(), k, c->sh_r1, n_owners, c->shard_tiles, 0u, c->j_bits, pack_format(c, k, c->sh_r1, c->sh_owner_bits, 0u), PassSel{0, 0u, 0u},
          c->cursor1.as<uint32_t>(), nullptr, nullptr, lim, c->peers_d, 1)));
    else
      NRP_LAUNCH((k_split_from_ascii<KeyT, 2, false><<<split_grid(c), kTileThreads, sizeof(SplitSmem<KeyT>) + sizeof(TileStage), st>>>(
          P, c->flags.as<uint8_t>(), k, c->sh_r1, n_owners, c->shard_tiles, 0u, c->j_bits, pack_format(c, k, 0), PassSel{0, 0u, 0u},
          c->cursor1.as<uint32_t>(), nullptr, nullptr, lim, c->peers_d, 1)));
  }
  PeerEnds pe;
  for (int o = 0; o < kMaxPeers; ++o) pe.ptr[o] = c->peer_ends[o];
  NRP_LAUNCH((k_push_region_ends<<<grid_for(nd, 256), 256, 0, st>>>(c->cursor1.as<uint32_t>(), nd, c->sh_r1, (uint32_t)n_owners, (uint32_t)c->my_rank, c->sh_cap1s, pe,
                                                                  cnt + CNT_SENT)));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_counters + 44, cnt + CNT_SCATTER_OVERFLOW, 16, cudaMemcpyDeviceToHost, st));
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(st));
  *overflow_out = (uint32_t)c->h_counters[44] != 0 ? 1 : 0;
  c->sh_sent = (int64_t)c->h_counters[45];
  return NRP_OK;
}

template <typename KeyT>
int shard_group_direct_impl(nrp_ctx* c, int k, int internal_repeats) {
  TRY(configure_kernels<KeyT>(c));
  StageTimer tm(c);
  BulkSource<KeyT> src;
  src.from_parts = false;
  src.packed = c->sh_packed;
  src.drop = c->sh_r1;
  src.owner_bits = c->sh_packed ? c->sh_owner_bits : 0;
  src.keys = c->recv_keys_d.ptr; src.vals = c->recv_vals_d.as<uint32_t>();
  src.n_seg = (uint32_t)c->sh_world << c->sh_r1;
  src.ext = SegExtents{nullptr, 0, c->recv_ends.as<uint32_t>(), c->sh_cap1s};
  src.segs_per_bucket = (uint32_t)c->sh_world;
  src.prepartitioned = true;
  src.r_done = c->sh_r1;
  src.m_upper = (int64_t)src.n_seg * c->sh_cap1s;
  src.m_plan = c->sh_m_owner;
  src.bits_hint = c->sh_bits;
  TRY((stage_bulk<KeyT>(c, k, src, tm)));
  TRY((stage_sort_mark<KeyT>(c, k, c->n_parts, !internal_repeats, tm)));
  tm.finish(c->timings);
  return NRP_OK;
}

}  // namespace

extern "C" {

int nrp_build(nrp_ctx* c, int k, int internal_repeats, int want_edges) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (k < 1 || k > 32) return fail(NRP_E_UNSUPPORTED, "k=%d outside [1, 32]", k);
  if (c->staged_lo != 0 || c->staged_hi != c->n_parts) return fail(NRP_E_STATE, "only the parts [%lld, %lld) are staged (nrp_load_parts_shard): nrp_build needs all of them",
                                                                    (long long)c->staged_lo, (long long)c->staged_hi);
  CU(cudaSetDevice(c->device));
  c->have_result = false;
  return k <= 16 ? build_impl<uint32_t>(c, k, internal_repeats, want_edges) : build_impl<uint64_t>(c, k, internal_repeats, want_edges);
}

int nrp_build_generic(nrp_ctx* c, int k, int internal_repeats, int want_edges, uint64_t seed, const uint8_t* preset_flags, int* status) {
  if (!c || !status) return fail(NRP_E_ARG, "null argument");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (k < 1) return fail(NRP_E_ARG, "k=%d must be positive", k);
  if (c->staged_lo != 0 || c->staged_hi != c->n_parts) return fail(NRP_E_STATE, "nrp_build_generic needs all parts staged");
  CU(cudaSetDevice(c->device));
  c->have_result = false;
  const int r = build_generic_impl(c, k, internal_repeats, want_edges, seed, preset_flags, status);
  if (r != NRP_OK || *status != 0) c->generic = false;
  return r;
}

int nrp_ctx_set_stream(nrp_ctx* c, void* cuda_stream) {
  if (!c) return fail(NRP_E_ARG, "null context");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  if (!c->external_stream) cudaStreamDestroy(c->stream);
  if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->external_stream = true; }
  else { CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->external_stream = false; }
  return NRP_OK;
}

int nrp_shard_extract(nrp_ctx* c, int64_t part_lo, int64_t part_hi, int k, int internal_repeats, int n_owners, int64_t* owner_counts) {
  if (!c || !owner_counts) return fail(NRP_E_ARG, "null argument");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (k < 1 || k > 32) return fail(NRP_E_UNSUPPORTED, "k=%d outside [1, 32]", k);
  if (part_lo < 0 || part_hi < part_lo || part_hi > c->n_parts) return fail(NRP_E_ARG, "bad shard range");
  if (n_owners < 1 || n_owners > 256) return fail(NRP_E_ARG, "n_owners must be in [1, 256]");
  if (part_lo < c->staged_lo || part_hi > c->staged_hi) return fail(NRP_E_STATE, "shard [%lld, %lld) is not staged on this device", (long long)part_lo, (long long)part_hi);
  CU(cudaSetDevice(c->device));
  return k <= 16 ? shard_extract_impl<uint32_t>(c, part_lo, part_hi, k, internal_repeats, n_owners, owner_counts)
                 : shard_extract_impl<uint64_t>(c, part_lo, part_hi, k, internal_repeats, n_owners, owner_counts);
}

int nrp_shard_send_buffers(nrp_ctx* c, void** keys, void** vals, int* key_bytes) {
  if (!c || !keys || !vals || !key_bytes) return fail(NRP_E_ARG, "null argument");
  *keys = c->keysA.ptr; *vals = c->valsA.ptr; *key_bytes = c->key_bytes;
  return NRP_OK;
}

int nrp_shard_group(nrp_ctx* c, const void* keys_dev, const uint32_t* vals_dev, int64_t n_records, int k, int internal_repeats) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (n_records < 0 || (n_records > 0 && (!keys_dev || !vals_dev))) return fail(NRP_E_ARG, "bad record buffers");
  CU(cudaSetDevice(c->device));
  return k <= 16 ? shard_group_impl<uint32_t>(c, keys_dev, vals_dev, n_records, k, internal_repeats)
                 : shard_group_impl<uint64_t>(c, keys_dev, vals_dev, n_records, k, internal_repeats);
}

int nrp_flags_export(nrp_ctx* c, uint8_t* dst_dev) {
  if (!c || !dst_dev) return fail(NRP_E_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  if (c->n_parts > 0) CU(cudaMemcpyAsync(dst_dev, c->flags.ptr, (size_t)c->n_parts, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_flags_import(nrp_ctx* c, const uint8_t* src_dev) {
  if (!c || !src_dev) return fail(NRP_E_ARG, "null argument");
  CU(cudaSetDevice(c->device));
  if (c->n_parts > 0) CU(cudaMemcpyAsync(c->flags.ptr, src_dev, (size_t)c->n_parts, cudaMemcpyDeviceToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_flags_repeat_list(nrp_ctx* c, uint32_t* list_dev, int cap) {
  if (!c || !list_dev || cap < 1) return fail(NRP_E_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  CU(cudaMemsetAsync(list_dev, 0, 4, c->stream));
  if (c->n_parts > 0)
    NRP_LAUNCH((k_repeat_list<<<grid_for((c->n_parts + 15) / 16, 256), 256, 0, c->stream>>>(c->flags.as<uint8_t>(), c->n_parts, list_dev, (uint32_t)cap)));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_flags_repeat_apply(nrp_ctx* c, const uint32_t* lists_dev, int world, int cap) {
  if (!c || !lists_dev || cap < 1 || world < 1) return fail(NRP_E_ARG, "bad argument");
  CU(cudaSetDevice(c->device));
  uint32_t* overflow = reinterpret_cast<uint32_t*>(c->counters.as<unsigned long long>() + CNT_FLAG_OVERFLOW);
  CU(cudaMemsetAsync(overflow, 0, 8, c->stream));
  NRP_LAUNCH((k_repeat_apply<<<32, 256, 0, c->stream>>>(lists_dev, world, (uint32_t)cap, c->flags.as<uint8_t>(), c->n_parts, overflow)));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_counters + 56, overflow, 8, cudaMemcpyDeviceToHost, c->stream));
  c->flag_lists_applied = true;
  return NRP_OK;          // asynchronous: nrp_shard_finish reports a truncated list (NRP_RETRY_DENSE_FLAGS)
}

int nrp_shard_finish(nrp_ctx* c, int k, int internal_repeats, int want_edges) {
  if (!c) return fail(NRP_E_ARG, "null context");
  CU(cudaSetDevice(c->device));
  const int r = k <= 16 ? shard_finish_impl<uint32_t>(c, k, internal_repeats, want_edges) : shard_finish_impl<uint64_t>(c, k, internal_repeats, want_edges);
  if (r == NRP_OK && c->flag_lists_applied) {
    c->flag_lists_applied = false;
    if ((uint32_t)c->h_counters[56] != 0) return NRP_RETRY_DENSE_FLAGS;     // (the stream was synchronised by the stage)
  }
  return r;
}

int nrp_shard_last_single(nrp_ctx* c, const uint64_t* shared_keys, int64_t n_shared, int k, int keys_on_device) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (n_shared < 0 || (n_shared > 0 && !shared_keys)) return fail(NRP_E_ARG, "bad key array");
  CU(cudaSetDevice(c->device));
  TRY(c->shared_keysA.ensure(8 * (size_t)(n_shared + 1)));
  if (n_shared > 0)
    CU(cudaMemcpyAsync(c->shared_keysA.ptr, shared_keys, 8 * (size_t)n_shared, keys_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  StageTimer tm(c);
  TRY(stage_last_single(c, k, c->shared_keysA.as<uint64_t>(), n_shared, c->shard_lo, c->shard_hi, tm));
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_edges_merge(nrp_ctx* c, const uint32_t* u_dev, const uint32_t* v_dev, int64_t n) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (n < 0 || (n > 0 && !u_dev)) return fail(NRP_E_ARG, "bad edge arrays");
  CU(cudaSetDevice(c->device));
  c->counts[NRP_C_EDGES] = 0;
  if (n > 0) {
    // the inputs may alias e_u / e_v of this context (zero-copy views): gather into fresh lists
    DevBuf& new_u = c->e_u_alt; DevBuf& new_v = c->e_v_alt;
    TRY(new_u.ensure(4 * (size_t)(n + 1))); TRY(new_v.ensure(4 * (size_t)(n + 1)));
    int log_slots = 6;
    while (((int64_t)1 << log_slots) < 2 * n) ++log_slots;
    TRY(c->e_codesA.ensure((size_t)8 << log_slots));
    CU(cudaMemsetAsync(c->e_codesA.ptr, 0xff, (size_t)8 << log_slots, c->stream));
    CU(cudaMemsetAsync(c->counters.as<unsigned long long>() + CNT_EDGES, 0, 8, c->stream));
    if (v_dev)
      NRP_LAUNCH((k_edge_merge<<<grid_for(n, 256), 256, 0, c->stream>>>(u_dev, v_dev, n, c->e_codesA.as<uint64_t>(), log_slots, new_u.as<uint32_t>(),
                                                                    new_v.as<uint32_t>(), c->counters.as<unsigned long long>() + CNT_EDGES)));
    else      // u_dev holds n 64-bit codes (u << 32) | v
      NRP_LAUNCH((k_edge_merge_codes<<<grid_for(n, 256), 256, 0, c->stream>>>(reinterpret_cast<const uint64_t*>(u_dev), n, c->e_codesA.as<uint64_t>(),
                                                                          log_slots, new_u.as<uint32_t>(), new_v.as<uint32_t>(),
                                                                          c->counters.as<unsigned long long>() + CNT_EDGES)));
    CU(cudaGetLastError());
    TRY(edge_set_count(c));
    std::swap(c->e_u, c->e_u_alt); std::swap(c->e_v, c->e_v_alt);
  }
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_result_device(nrp_ctx* c, int what, void** ptr, int64_t* count) {
  if (!c || !ptr || !count) return fail(NRP_E_ARG, "null argument");
  if (!c->have_result) return fail(NRP_E_STATE, "no finished build");
  switch (what) {
    case NRP_D_GROUP_KEYS: *ptr = c->g_keys.ptr; *count = c->counts[NRP_C_GROUPS]; break;
    case NRP_D_EDGE_U: *ptr = c->e_u.ptr; *count = c->counts[NRP_C_EDGES] < 0 ? 0 : c->counts[NRP_C_EDGES]; break;
    case NRP_D_EDGE_V: *ptr = c->e_v.ptr; *count = c->counts[NRP_C_EDGES] < 0 ? 0 : c->counts[NRP_C_EDGES]; break;
    case NRP_D_FLAGS: *ptr = c->flags.ptr; *count = c->n_parts; break;
    case NRP_D_LAST_SINGLE: *ptr = c->last_single.ptr; *count = c->n_parts; break;
    case NRP_D_INDPTR: *ptr = c->g_indptr.ptr; *count = c->counts[NRP_C_GROUPS] + 1; break;
    case NRP_D_MEMBERS: *ptr = c->g_members.ptr; *count = c->counts[NRP_C_MEMBERS]; break;
    case NRP_D_FIRST_WINDOW: *ptr = c->g_first_window.ptr; *count = c->counts[NRP_C_GROUPS]; break;
    default: return fail(NRP_E_ARG, "unknown result selector %d", what);
  }
  return NRP_OK;
}

int nrp_peer_export(nrp_ctx* c, int64_t capacity_records, int k, unsigned char* handle_keys, unsigned char* handle_vals) {
  if (!c || !handle_keys || !handle_vals) return fail(NRP_E_ARG, "null argument");
  if (capacity_records < 1 || capacity_records >= 0xffffe000ll) return fail(NRP_E_ARG, "bad capacity");
  CU(cudaSetDevice(c->device));
  const size_t key_bytes = k <= 16 ? 4 : 8;
  TRY(c->recv_keys.ensure(key_bytes * (size_t)(capacity_records + 16)));
  TRY(c->recv_vals.ensure(4 * (size_t)(capacity_records + 16)));
  c->recv_capacity = capacity_records;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hk, hv;
  CU(cudaIpcGetMemHandle(&hk, c->recv_keys.ptr));
  CU(cudaIpcGetMemHandle(&hv, c->recv_vals.ptr));
  memcpy(handle_keys, &hk, 64);
  memcpy(handle_vals, &hv, 64);
  return NRP_OK;
}

int nrp_peer_open(nrp_ctx* c, int n_peers, int my_rank, const unsigned char* handles) {
  if (!c || !handles) return fail(NRP_E_ARG, "null argument");
  if (n_peers < 1 || n_peers > kMaxPeers || my_rank < 0 || my_rank >= n_peers) return fail(NRP_E_ARG, "bad peer count");
  CU(cudaSetDevice(c->device));
  auto open_one = [&](const unsigned char* h, void** out) -> int {
    for (size_t i = 0; i < c->opened_handles.size(); ++i)
      if (memcmp(c->opened_handles[i].data(), h, 64) == 0) { *out = c->opened_ptrs[i]; return NRP_OK; }
    cudaIpcMemHandle_t handle;
    memcpy(&handle, h, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess));
    c->opened_handles.emplace_back(h, h + 64);
    c->opened_ptrs.push_back(ptr);
    *out = ptr;
    return NRP_OK;
  };
  for (int p = 0; p < n_peers; ++p) {
    if (p == my_rank) { c->peers.keys[p] = c->recv_keys.ptr; c->peers.vals[p] = c->recv_vals.as<uint32_t>(); continue; }
    void* pk = nullptr; void* pv = nullptr;
    TRY(open_one(handles + (size_t)p * 128, &pk));
    TRY(open_one(handles + (size_t)p * 128 + 64, &pv));
    c->peers.keys[p] = pk; c->peers.vals[p] = reinterpret_cast<uint32_t*>(pv);
  }
  c->n_peers = n_peers; c->my_rank = my_rank;
  return NRP_OK;
}

int nrp_shard_count(nrp_ctx* c, int64_t part_lo, int64_t part_hi, int k, int internal_repeats, int n_owners, int64_t* owner_counts) {
  if (!c || !owner_counts) return fail(NRP_E_ARG, "null argument");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (k < 1 || k > 32) return fail(NRP_E_UNSUPPORTED, "k=%d outside [1, 32]", k);
  if (part_lo < 0 || part_hi < part_lo || part_hi > c->n_parts) return fail(NRP_E_ARG, "bad shard range");
  if (n_owners < 1 || n_owners > kMaxPeers) return fail(NRP_E_ARG, "n_owners must be in [1, %d]", kMaxPeers);
  if (part_lo < c->staged_lo || part_hi > c->staged_hi) return fail(NRP_E_STATE, "shard [%lld, %lld) is not staged on this device", (long long)part_lo, (long long)part_hi);
  CU(cudaSetDevice(c->device));
  return k <= 16 ? shard_count_impl<uint32_t>(c, part_lo, part_hi, k, internal_repeats, n_owners, owner_counts)
                 : shard_count_impl<uint64_t>(c, part_lo, part_hi, k, internal_repeats, n_owners, owner_counts);
}

int nrp_shard_scatter_peers(nrp_ctx* c, int k, const int64_t* write_offsets) {
  if (!c || !write_offsets) return fail(NRP_E_ARG, "null argument");
  if (c->n_peers != c->n_owners || c->n_owners < 1) return fail(NRP_E_STATE, "peers are not open for %d owners", c->n_owners);
  CU(cudaSetDevice(c->device));
  return k <= 16 ? shard_scatter_impl<uint32_t>(c, k, write_offsets) : shard_scatter_impl<uint64_t>(c, k, write_offsets);
}

int nrp_shard_group_received(nrp_ctx* c, int64_t n_records, int k, int internal_repeats) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (n_records < 0 || n_records > c->recv_capacity) return fail(NRP_E_ARG, "record count exceeds the receive capacity");
  return nrp_shard_group(c, c->recv_keys.ptr, c->recv_vals.as<uint32_t>(), n_records, k, internal_repeats);
}

int nrp_peer_export_direct(nrp_ctx* c, int64_t windows_total, int n_owners, int k, unsigned char* handles) {
  if (!c || !handles) return fail(NRP_E_ARG, "null argument");
  if (n_owners < 1 || n_owners > kMaxPeers || windows_total < 0) return fail(NRP_E_ARG, "bad arguments");
  if (k < 1 || k > 32) return fail(NRP_E_UNSUPPORTED, "k=%d outside [1, 32]", k);
  CU(cudaSetDevice(c->device));
  int r1 = 0, bits = 0; uint32_t cap1s = 0; int64_t m_owner = 0;
  TRY(direct_shard_plan(c, windows_total, n_owners, &r1, &bits, &cap1s, &m_owner));
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  // Record layout of the exchange (every rank stages the same parts: the same decision everywhere).  With a power-of-two number
  // of owners the owner is the top bits of the bijective hash, which packed records then need not store; if the words still do
  // not fit with the window index in the value, it is left out (group ranks then come from a scan of the few shared k-mers).
  int owner_bits = 0;
  if ((n_owners & (n_owners - 1)) == 0) while ((1 << owner_bits) < n_owners) ++owner_bits;
  const size_t key_bytes = k <= 16 ? 4 : 8;
  record_layout(c, k, false);
  bool no_window = false;
  bool packed = packed_fits(c, k, r1, bits, owner_bits);
  if (!packed && c->j_bits > 0) {
    record_layout(c, k, true);
    if (packed_fits(c, k, r1, bits, owner_bits)) { packed = true; no_window = true; }
    else record_layout(c, k, false);
  }
  if (!packed) {
    // a few more pre-partition bits (still long runs over NVLink) may be all that is missing: they are implied by the region
    int r1_eff = 0;
    while (((int64_t)n_owners << (r1_eff + 1)) <= c->opt_nvlink_digits) ++r1_eff;
    for (int r = r1 + 1; r <= r1_eff && r < bits && !packed; ++r) {
      for (int nw = 0; nw < 2 && !packed; ++nw) {
        record_layout(c, k, nw != 0);
        if (nw == 1 && c->j_bits != 0) continue;
        if (packed_fits(c, k, r, bits, owner_bits)) { packed = true; no_window = nw != 0; r1 = r; }
      }
    }
    if (packed) cap1s = region_cap((double)windows_total / ((double)n_owners * n_owners * (double)(1 << r1)), 256);
    else record_layout(c, k, false);
  }
  if (!packed) owner_bits = 0;
  c->sh_owner_bits = owner_bits; c->sh_no_window = no_window;
  c->sh_bits = bits;
  const int64_t n_regions = (int64_t)n_owners << r1;
  const int64_t capacity = n_regions * cap1s;
  if (capacity >= 0xffffe000ll) return fail(NRP_E_UNSUPPORTED, "receive regions exceed 32-bit record indices");
  TRY(c->recv_keys_d.ensure((packed ? 8 : key_bytes) * (size_t)(capacity + 16)));
  TRY(c->recv_vals_d.ensure(4 * (size_t)(capacity + 16)));
  TRY(c->recv_ends.ensure(4 * (size_t)(kLevelBuckets * kMaxPeers + 4)));
  c->sh_world = n_owners; c->sh_r1 = r1; c->sh_cap1s = cap1s; c->sh_key_bytes = (int)key_bytes; c->sh_packed = packed;
  c->sh_windows_total = windows_total; c->sh_m_owner = m_owner;
  cudaIpcMemHandle_t h[3];
  CU(cudaIpcGetMemHandle(&h[0], c->recv_keys_d.ptr));
  CU(cudaIpcGetMemHandle(&h[1], c->recv_vals_d.ptr));
  CU(cudaIpcGetMemHandle(&h[2], c->recv_ends.ptr));
  for (int i = 0; i < 3; ++i) memcpy(handles + 64 * i, &h[i], 64);
  return NRP_OK;
}

int nrp_peer_open_direct(nrp_ctx* c, int n_peers, int my_rank, const unsigned char* handles) {
  if (!c || !handles) return fail(NRP_E_ARG, "null argument");
  if (n_peers < 1 || n_peers > kMaxPeers || my_rank < 0 || my_rank >= n_peers) return fail(NRP_E_ARG, "bad peer count");
  if (n_peers != c->sh_world) return fail(NRP_E_STATE, "nrp_peer_export_direct was called for %d owners", c->sh_world);
  CU(cudaSetDevice(c->device));
  auto open_one = [&](const unsigned char* h, void** out) -> int {
    for (size_t i = 0; i < c->opened_handles.size(); ++i)
      if (memcmp(c->opened_handles[i].data(), h, 64) == 0) { *out = c->opened_ptrs[i]; return NRP_OK; }
    cudaIpcMemHandle_t handle;
    memcpy(&handle, h, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess));
    c->opened_handles.emplace_back(h, h + 64);
    c->opened_ptrs.push_back(ptr);
    *out = ptr;
    return NRP_OK;
  };
  for (int p = 0; p < n_peers; ++p) {
    if (p == my_rank) {
      c->peers_d.keys[p] = c->recv_keys_d.ptr; c->peers_d.vals[p] = c->recv_vals_d.as<uint32_t>(); c->peer_ends[p] = c->recv_ends.as<uint32_t>();
      continue;
    }
    void* pk = nullptr; void* pv = nullptr; void* pe = nullptr;
    TRY(open_one(handles + (size_t)p * 192, &pk));
    TRY(open_one(handles + (size_t)p * 192 + 64, &pv));
    TRY(open_one(handles + (size_t)p * 192 + 128, &pe));
    c->peers_d.keys[p] = pk; c->peers_d.vals[p] = reinterpret_cast<uint32_t*>(pv); c->peer_ends[p] = reinterpret_cast<uint32_t*>(pe);
  }
  c->n_peers_d = n_peers; c->my_rank = my_rank;
  return NRP_OK;
}

int nrp_shard_scatter_direct(nrp_ctx* c, int64_t part_lo, int64_t part_hi, int k, int internal_repeats, int n_owners, int* overflow) {
  if (!c || !overflow) return fail(NRP_E_ARG, "null argument");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (k < 1 || k > 32) return fail(NRP_E_UNSUPPORTED, "k=%d outside [1, 32]", k);
  if (part_lo < 0 || part_hi < part_lo || part_hi > c->n_parts) return fail(NRP_E_ARG, "bad shard range");
  if (c->n_peers_d != n_owners || c->sh_world != n_owners) return fail(NRP_E_STATE, "direct peers are not open for %d owners", n_owners);
  if (c->sh_key_bytes != (k <= 16 ? 4 : 8)) return fail(NRP_E_STATE, "receive buffers were exported for another key width");
  if (part_lo < c->staged_lo || part_hi > c->staged_hi) return fail(NRP_E_STATE, "shard [%lld, %lld) is not staged on this device", (long long)part_lo, (long long)part_hi);
  CU(cudaSetDevice(c->device));
  return k <= 16 ? shard_scatter_direct_impl<uint32_t>(c, part_lo, part_hi, k, internal_repeats, n_owners, overflow)
                 : shard_scatter_direct_impl<uint64_t>(c, part_lo, part_hi, k, internal_repeats, n_owners, overflow);
}

int nrp_shard_group_direct(nrp_ctx* c, int k, int internal_repeats) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (c->n_peers_d < 1) return fail(NRP_E_STATE, "direct peers are not open");
  CU(cudaSetDevice(c->device));
  return k <= 16 ? shard_group_direct_impl<uint32_t>(c, k, internal_repeats) : shard_group_direct_impl<uint64_t>(c, k, internal_repeats);
}

int nrp_parts_capacity(nrp_ctx* c, int64_t* bytes) {
  if (!c || !bytes) return fail(NRP_E_ARG, "null argument");
  *bytes = (int64_t)c->ascii.cap;
  return NRP_OK;
}

int nrp_peer_export_parts(nrp_ctx* c, unsigned char* handle) {
  if (!c || !handle) return fail(NRP_E_ARG, "null argument");
  if (!c->have_parts || !c->ascii.ptr) return fail(NRP_E_STATE, "no parts are staged");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->ascii.ptr));
  memcpy(handle, &h, 64);
  return NRP_OK;
}

int nrp_peer_open_parts(nrp_ctx* c, int n_peers, int my_rank, const unsigned char* handles, const int64_t* part_bounds) {
  if (!c || !handles || !part_bounds) return fail(NRP_E_ARG, "null argument");
  if (n_peers < 1 || n_peers > kMaxPeers || my_rank < 0 || my_rank >= n_peers) return fail(NRP_E_ARG, "bad peer count");
  CU(cudaSetDevice(c->device));
  PartHomes homes{};
  for (int p = 0; p < n_peers; ++p) {
    homes.first[p] = part_bounds[p];
    if (p == my_rank) { homes.base[p] = c->ascii.as<uint8_t>(); continue; }
    const unsigned char* h = handles + (size_t)p * 64;
    void* ptr = nullptr;
    for (size_t i = 0; i < c->opened_handles.size(); ++i)
      if (memcmp(c->opened_handles[i].data(), h, 64) == 0) ptr = c->opened_ptrs[i];
    if (!ptr) {
      cudaIpcMemHandle_t handle;
      memcpy(&handle, h, 64);
      CU(cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess));
      c->opened_handles.emplace_back(h, h + 64);
      c->opened_ptrs.push_back(ptr);
    }
    homes.base[p] = reinterpret_cast<const uint8_t*>(ptr);
  }
  homes.first[n_peers] = part_bounds[n_peers];
  homes.n = n_peers;
  c->homes = homes;
  c->my_rank = my_rank;
  return NRP_OK;
}

int nrp_peer_close(nrp_ctx* c) {
  if (!c) return fail(NRP_E_ARG, "null context");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  c->homes = PartHomes{};
  for (void* p : c->opened_ptrs) cudaIpcCloseMemHandle(p);
  c->opened_ptrs.clear();
  c->opened_handles.clear();
  c->peers = PeerBuffers{}; c->peers_d = PeerBuffers{};
  for (int o = 0; o < kMaxPeers; ++o) c->peer_ends[o] = nullptr;
  c->n_peers = 0; c->n_peers_d = 0;
  return NRP_OK;
}

int nrp_shard_moved(nrp_ctx* c, int64_t out[2]) {
  if (!c || !out) return fail(NRP_E_ARG, "null argument");
  out[0] = c->sh_sent; out[1] = c->n_records;
  return NRP_OK;
}

int nrp_result_counts(nrp_ctx* c, int64_t out[NRP_N_COUNTS]) {
  if (!c || !out) return fail(NRP_E_ARG, "null argument");
  if (!c->have_result) return fail(NRP_E_STATE, "no finished build");
  memcpy(out, c->counts, sizeof(c->counts));
  return NRP_OK;
}

int nrp_result_timings(nrp_ctx* c, double out[NRP_N_TIMINGS]) {
  if (!c || !out) return fail(NRP_E_ARG, "null argument");
  if (!c->have_result) return fail(NRP_E_STATE, "no finished build");
  memcpy(out, c->timings, sizeof(c->timings));
  return NRP_OK;
}

#define NEED_RESULT()                                               \
  if (!c) return fail(NRP_E_ARG, "null context");                   \
  if (!c->have_result) return fail(NRP_E_STATE, "no finished build"); \
  CU(cudaSetDevice(c->device))

int nrp_result_flags(nrp_ctx* c, uint8_t* flags) {
  NEED_RESULT();
  if (c->n_parts > 0) CU(cudaMemcpyAsync(flags, c->flags.ptr, (size_t)c->n_parts, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_result_groups(nrp_ctx* c, int64_t* indptr, uint32_t* members, uint32_t* first_window, uint64_t* keys) {
  NEED_RESULT();
  const int64_t ng = c->counts[NRP_C_GROUPS], nm = c->counts[NRP_C_MEMBERS];
  if (indptr) CU(cudaMemcpyAsync(indptr, c->g_indptr.ptr, 8 * (size_t)(ng + 1), cudaMemcpyDeviceToHost, c->stream));
  if (members && nm > 0) CU(cudaMemcpyAsync(members, c->g_members.ptr, 4 * (size_t)nm, cudaMemcpyDeviceToHost, c->stream));
  if (first_window && ng > 0) CU(cudaMemcpyAsync(first_window, c->g_first_window.ptr, 4 * (size_t)ng, cudaMemcpyDeviceToHost, c->stream));
  if (keys && ng > 0) CU(cudaMemcpyAsync(keys, c->g_keys.ptr, 8 * (size_t)ng, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_result_host(nrp_ctx* c, int want_keys, nrp_host_result* out) {
  NEED_RESULT();
  if (!out) return fail(NRP_E_ARG, "null argument");
  const int64_t n = c->n_parts, ng = c->counts[NRP_C_GROUPS], nm = c->counts[NRP_C_MEMBERS];
  const int64_t ne = c->counts[NRP_C_EDGES] < 0 ? 0 : c->counts[NRP_C_EDGES];
  const ResultLayout L(n, ng, nm, ne);
  // what the build already sent on its way (early_fetch_enqueue) is not copied again -- unless the buffer has to grow for the edges
  bool early = c->early_valid && c->early_n == n && c->early_ng == ng && c->early_nm == nm;
  if (early && L.total > c->h_results.cap) { CU(cudaStreamSynchronize(c->fetch_stream)); early = false; }
  c->early_valid = false;
  TRY(c->h_results.ensure(L.total));
  unsigned char* h = c->h_results.as<unsigned char>();
  cudaStream_t st = c->stream;
  if (!early) {
    CU(cudaMemcpyAsync(h + L.indptr, c->g_indptr.ptr, 8 * (size_t)(ng + 1), cudaMemcpyDeviceToHost, st));
    if (want_keys && ng > 0) CU(cudaMemcpyAsync(h + L.keys, c->g_keys.ptr, 8 * (size_t)ng, cudaMemcpyDeviceToHost, st));
    if (nm > 0) CU(cudaMemcpyAsync(h + L.members, c->g_members.ptr, 4 * (size_t)nm, cudaMemcpyDeviceToHost, st));
    if (ng > 0) CU(cudaMemcpyAsync(h + L.fw, c->g_first_window.ptr, 4 * (size_t)ng, cudaMemcpyDeviceToHost, st));
    if (n > 0) CU(cudaMemcpyAsync(h + L.flags, c->flags.ptr, (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  if (n > 0) CU(cudaMemcpyAsync(h + L.last, c->last_single.ptr, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (ne > 0) {
    CU(cudaMemcpyAsync(h + L.eu, c->e_u.ptr, 4 * (size_t)ne, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h + L.ev, c->e_v.ptr, 4 * (size_t)ne, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  if (early) CU(cudaEventSynchronize(c->fetch_done));
  out->flags = h + L.flags;
  out->indptr = reinterpret_cast<const int64_t*>(h + L.indptr);
  out->members = reinterpret_cast<const uint32_t*>(h + L.members);
  out->first_window = reinterpret_cast<const uint32_t*>(h + L.fw);
  out->keys = want_keys ? reinterpret_cast<const uint64_t*>(h + L.keys) : nullptr;
  out->last_single = reinterpret_cast<const int32_t*>(h + L.last);
  out->edge_u = c->counts[NRP_C_EDGES] < 0 ? nullptr : reinterpret_cast<const uint32_t*>(h + L.eu);
  out->edge_v = c->counts[NRP_C_EDGES] < 0 ? nullptr : reinterpret_cast<const uint32_t*>(h + L.ev);
  out->n_parts = n; out->n_groups = ng; out->n_members = nm; out->n_edges = ne;
  return NRP_OK;
}

int nrp_result_last_single(nrp_ctx* c, int32_t* last_single) {
  NEED_RESULT();
  if (c->n_parts > 0) CU(cudaMemcpyAsync(last_single, c->last_single.ptr, 4 * (size_t)c->n_parts, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

int nrp_result_edges(nrp_ctx* c, uint32_t* u, uint32_t* v) {
  NEED_RESULT();
  const int64_t ne = c->counts[NRP_C_EDGES];
  if (ne < 0) return fail(NRP_E_STATE, "edges were not requested");
  if (ne > 0) {
    CU(cudaMemcpyAsync(u, c->e_u.ptr, 4 * (size_t)ne, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(v, c->e_v.ptr, 4 * (size_t)ne, cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return NRP_OK;
}

}  // extern "C"
