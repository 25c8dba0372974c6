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
#include "stages.cuh"

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
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

enum { CNT_CAND = 0, CNT_REPEAT = 1, CNT_FLAGGED = 2, CNT_WORK = 3, CNT_OVERSIZED = 4, CNT_N = 8 };

constexpr int kFilterThreads = 1024;
constexpr int kFilterItems = 8;
constexpr int kFilterCap = kFilterThreads * kFilterItems;   // 8192 records per leaf bucket in shared memory
constexpr int kMaxLeafBits = 15;                            // 2^15 u32 counters = 128 KB of shared memory

}  // namespace

struct nrp_ctx {
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;
  bool configured[2] = {false, false};

  // options
  int64_t opt_filter_cap = kFilterCap;
  int64_t opt_leaf_target = 5500;
  int64_t opt_max_level_bits = kMaxLevelBits;
  int64_t opt_max_pairs = (int64_t)1 << 31;

  // background
  DevBuf bg_table; int bg_K = 0; int bg_log_slots = 0; int bg_key_bytes = 0; int64_t bg_n = 0;

  // loaded parts
  DevBuf ascii, off, part_id, tile_first;
  std::vector<int64_t> h_off;
  int64_t n_parts = 0, total_len = 0, n_tiles = 0;
  int uniform_len = 0;
  bool have_parts = false, have_ids = false;
  uint32_t part_base = 0;

  // pipeline buffers
  DevBuf flags, hist, leaf_off, cursor1, cursor2, seg_off, tile_prefix, counters, scan_scratch, lsd_scratch;
  DevBuf keysA, valsA, keysB, valsB;
  DevBuf head, head_scan, group_start, valid, valid_scan, mem_cnt, mem_scan, pair_cnt, pair_scan;
  DevBuf g_keys, g_indptr, g_pair_off, g_members, g_first_part, g_first_window, last_single;
  DevBuf e_codesA, e_codesB, e_head, e_head_scan, e_u, e_v;
  unsigned long long* h_counters = nullptr;   // pinned

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
  return P;
}

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

// number of leaf-bucket bits for m records
int choose_leaf_bits(const nrp_ctx* c, int64_t m) {
  int bits = 1;
  while (bits < kMaxLeafBits && (m >> bits) > c->opt_leaf_target) ++bits;
  int cap_bits = (int)(2 * c->opt_max_level_bits);
  if (bits > cap_bits) bits = cap_bits;
  return bits;
}

template <typename KeyT>
int configure_kernels(nrp_ctx* c) {
  bool& done = c->configured[sizeof(KeyT) == 4 ? 0 : 1];
  if (done) return NRP_OK;
  const int split_smem = (int)(sizeof(SplitSmem<KeyT>) + sizeof(TileStage));
  CU(cudaFuncSetAttribute(k_split_from_ascii<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, split_smem));
  CU(cudaFuncSetAttribute(k_split_from_records<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem<KeyT>)));
  const int filter_smem = (int)((sizeof(KeyT) + 1) * 2 * kFilterCap);
  CU(cudaFuncSetAttribute((k_filter<KeyT, kFilterThreads, kFilterItems>), cudaFuncAttributeMaxDynamicSharedMemorySize, filter_smem));
  CU(cudaFuncSetAttribute(k_hist_from_ascii, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << kMaxLeafBits));
  CU(cudaFuncSetAttribute(k_hist_from_records<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << kMaxLeafBits));
  done = true;
  return NRP_OK;
}

// Sort the candidate records and turn key runs into groups.  On return the counters (host) hold the sizes.
// cand in (ck, cv) with alternates (ak, av); n_cand records.
template <typename KeyT>
int group_candidates(nrp_ctx* c, KeyT* ck, uint32_t* cv, KeyT* ak, uint32_t* av, int64_t n_cand, int key_bits, int val_bits,
                     int64_t n_parts_for_flags, bool mark_repeats, StageTimer& tm, KeyT** sorted_keys, uint32_t** sorted_vals,
                     int64_t* n_grouped, int64_t* n_repeat_records) {
  cudaStream_t st = c->stream;
  tm.begin(NRP_T_SORT);
  TRY(c->lsd_scratch.ensure(sizeof(uint32_t) * (size_t)lsd_scratch_elems(n_cand)));
  LsdPlan plan = lsd_plan(key_bits, val_bits);
  KeyT* sk; uint32_t* sv;
  lsd_sort<KeyT, true>(ck, cv, ak, av, n_cand, plan, c->lsd_scratch.as<uint32_t>(), st, &sk, &sv);
  *sorted_keys = sk; *sorted_vals = sv;

  tm.begin(NRP_T_GROUP);
  const size_t n1 = (size_t)n_cand + 2;
  TRY(c->head.ensure(4 * n1)); TRY(c->head_scan.ensure(4 * n1)); TRY(c->group_start.ensure(4 * n1));
  TRY(c->valid.ensure(4 * n1)); TRY(c->valid_scan.ensure(4 * n1));
  TRY(c->mem_cnt.ensure(4 * n1)); TRY(c->mem_scan.ensure(8 * n1));
  TRY(c->pair_cnt.ensure(8 * n1)); TRY(c->pair_scan.ensure(8 * n1));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems((int64_t)n1 + 1)));
  unsigned long long* cnt = c->counters.as<unsigned long long>();
  int64_t n = n_cand;
  *n_repeat_records = 0;
  if (n > 0) {
    NRP_LAUNCH((k_group_heads<KeyT><<<grid_for(n, 256), 256, 0, st>>>(sk, sv, n, c->part_base, n_parts_for_flags, c->head.as<uint32_t>(),
                                                                  mark_repeats ? c->flags.as<uint8_t>() : nullptr, cnt + CNT_REPEAT)));
    CU(cudaMemcpyAsync(c->h_counters + 18, cnt + CNT_REPEAT, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_repeat_records = (int64_t)c->h_counters[18];
    if (mark_repeats && *n_repeat_records > 0) {
      // drop every record of the flagged parts, keep the (key, part) order, and find the key runs again
      KeyT* ok = sk == ck ? ak : ck;
      uint32_t* ov = sv == cv ? av : cv;
      NRP_LAUNCH((k_keep_unflagged<<<grid_for(n, 256), 256, 0, st>>>(sv, n, c->flags.as<uint8_t>() - c->part_base, c->valid.as<uint32_t>())));
      exclusive_scan<uint32_t, uint32_t>(c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(), n, c->scan_scratch.as<uint32_t>(), st);
      NRP_LAUNCH((k_compact_records<KeyT><<<grid_for(n, 256), 256, 0, st>>>(sk, sv, c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(), n, ok, ov)));
      CU(cudaMemcpyAsync(c->h_counters + 19, c->valid_scan.as<uint32_t>() + n, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n = (int64_t)(uint32_t)c->h_counters[19];
      sk = ok; sv = ov;
      *sorted_keys = sk; *sorted_vals = sv;
      if (n > 0)
        NRP_LAUNCH((k_group_heads<KeyT><<<grid_for(n, 256), 256, 0, st>>>(sk, sv, n, c->part_base, n_parts_for_flags, c->head.as<uint32_t>(),
                                                                      nullptr, cnt + CNT_WORK)));
    }
  }
  *n_grouped = n;
  if (n > 0) {
    exclusive_scan<uint32_t, uint32_t>(c->head.as<uint32_t>(), c->head_scan.as<uint32_t>(), n, c->scan_scratch.as<uint32_t>(), st);
    NRP_LAUNCH((k_group_starts<<<grid_for(n + 1, 256), 256, 0, st>>>(c->head.as<uint32_t>(), c->head_scan.as<uint32_t>(), n,
                                                                 c->group_start.as<uint32_t>())));
  }
  CU(cudaGetLastError());
  return NRP_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int nrp_version(void) { return 100; }
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
  int r = c->counters.ensure(sizeof(unsigned long long) * CNT_N);
  if (r != NRP_OK) return r;
  *out = c;
  return NRP_OK;
}

int nrp_ctx_destroy(nrp_ctx* c) {
  if (!c) return NRP_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf* all[] = {&c->bg_table, &c->ascii, &c->off, &c->part_id, &c->tile_first, &c->flags, &c->hist, &c->leaf_off, &c->cursor1,
                   &c->cursor2, &c->seg_off, &c->tile_prefix, &c->counters, &c->scan_scratch, &c->lsd_scratch, &c->keysA, &c->valsA,
                   &c->keysB, &c->valsB, &c->head, &c->head_scan, &c->group_start, &c->valid, &c->valid_scan, &c->mem_cnt,
                   &c->mem_scan, &c->pair_cnt, &c->pair_scan, &c->g_keys, &c->g_indptr, &c->g_pair_off, &c->g_members,
                   &c->g_first_part, &c->g_first_window, &c->last_single, &c->e_codesA, &c->e_codesB, &c->e_head, &c->e_head_scan,
                   &c->e_u, &c->e_v};
  for (DevBuf* b : all) b->release();
  for (cudaEvent_t e : c->events) cudaEventDestroy(e);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  cudaStreamDestroy(c->stream);
  delete c;
  return NRP_OK;
}

int nrp_ctx_device(nrp_ctx* c) { return c ? c->device : -1; }

int nrp_ctx_set_option(nrp_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return fail(NRP_E_ARG, "null argument");
  std::string n(name);
  if (n == "filter_cap") c->opt_filter_cap = (value <= 0 || value > kFilterCap) ? kFilterCap : value;
  else if (n == "leaf_target") c->opt_leaf_target = value <= 0 ? 5500 : value;
  else if (n == "max_level_bits") c->opt_max_level_bits = (value <= 0 || value > kMaxLevelBits) ? kMaxLevelBits : value;
  else if (n == "max_pairs") c->opt_max_pairs = value <= 0 ? ((int64_t)1 << 31) : value;
  else return fail(NRP_E_ARG, "unknown option '%s'", name);
  return NRP_OK;
}

int nrp_bg_clear(nrp_ctx* c) {
  if (!c) return fail(NRP_E_ARG, "null context");
  c->bg_n = 0; c->bg_K = 0;
  return NRP_OK;
}

int nrp_bg_set(nrp_ctx* c, const uint64_t* canon_kmers, int64_t n, int K) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (n < 0 || (n > 0 && !canon_kmers)) return fail(NRP_E_ARG, "bad background array");
  if (K < 1 || K > 32) return fail(NRP_E_UNSUPPORTED, "background K=%d outside [1, 32]", K);
  CU(cudaSetDevice(c->device));
  if (n == 0) return nrp_bg_clear(c);
  int log_slots = 6;
  while (((int64_t)1 << log_slots) < 2 * n) ++log_slots;
  const int key_bytes = K <= 16 ? 4 : 8;
  const size_t table_bytes = (size_t)key_bytes << log_slots;
  TRY(c->bg_table.ensure(table_bytes));
  DevBuf staging;
  TRY(staging.ensure(sizeof(uint64_t) * (size_t)n));
  CU(cudaMemcpyAsync(staging.ptr, canon_kmers, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->bg_table.ptr, 0xff, table_bytes, c->stream));
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, c->n_sms * 8);
  if (key_bytes == 4) NRP_LAUNCH((k_bg_insert<uint32_t><<<grid, 256, 0, c->stream>>>(staging.as<uint64_t>(), n, c->bg_table.as<uint32_t>(), log_slots)));
  else NRP_LAUNCH((k_bg_insert<uint64_t><<<grid, 256, 0, c->stream>>>(staging.as<uint64_t>(), n, c->bg_table.as<uint64_t>(), log_slots)));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  staging.release();
  c->bg_n = n; c->bg_K = K; c->bg_log_slots = log_slots; c->bg_key_bytes = key_bytes;
  return NRP_OK;
}

int nrp_load_parts(nrp_ctx* c, const uint8_t* ascii, const int64_t* offsets, const uint32_t* part_id, int64_t n_parts, int ascii_on_device) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (n_parts < 0 || !offsets || (n_parts > 0 && !ascii && offsets[n_parts] > 0)) return fail(NRP_E_ARG, "bad part buffers");
  if (n_parts >= 0xfffffff0ll) return fail(NRP_E_UNSUPPORTED, "more than 2^32-16 parts");
  CU(cudaSetDevice(c->device));
  c->have_parts = false; c->have_result = false;
  if (offsets[0] != 0) return fail(NRP_E_ARG, "offsets[0] must be 0");
  const int64_t total = offsets[n_parts];
  int uniform = n_parts > 0 ? (int)std::min<int64_t>(offsets[1] - offsets[0], (int64_t)1 << 30) : 0;
  for (int64_t i = 0; i < n_parts; ++i) {
    int64_t len = offsets[i + 1] - offsets[i];
    if (len < 0) return fail(NRP_E_ARG, "offsets must be non-decreasing (part %lld)", (long long)i);
    if (len != uniform) uniform = 0;
  }
  if (uniform >= (1 << 30)) uniform = 0;
  c->h_off.assign(offsets, offsets + n_parts + 1);
  c->n_parts = n_parts; c->total_len = total; c->uniform_len = uniform;
  c->n_tiles = (total + kTileBytes - 1) / kTileBytes;
  const size_t padded = (size_t)c->n_tiles * kTileBytes + 256;
  TRY(c->ascii.ensure(padded));
  TRY(c->off.ensure(sizeof(int64_t) * (size_t)(n_parts + 1)));
  CU(cudaMemsetAsync(c->ascii.as<uint8_t>() + total, 0, padded - (size_t)total, c->stream));
  if (total > 0)
    CU(cudaMemcpyAsync(c->ascii.ptr, ascii, (size_t)total, ascii_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->off.ptr, offsets, sizeof(int64_t) * (size_t)(n_parts + 1), cudaMemcpyHostToDevice, c->stream));
  c->have_ids = part_id != nullptr;
  if (part_id && n_parts > 0) {
    TRY(c->part_id.ensure(sizeof(uint32_t) * (size_t)n_parts));
    CU(cudaMemcpyAsync(c->part_id.ptr, part_id, sizeof(uint32_t) * (size_t)n_parts, cudaMemcpyHostToDevice, c->stream));
  }
  TRY(c->tile_first.ensure(sizeof(uint32_t) * (size_t)(c->n_tiles + 2)));
  if (uniform == 0 && n_parts > 0)
    NRP_LAUNCH((k_tile_first_part<<<grid_for(c->n_tiles + 1, 256), 256, 0, c->stream>>>(c->off.as<int64_t>(), n_parts, total, c->n_tiles,
                                                                            c->tile_first.as<uint32_t>())));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->part_base = 0;
  c->have_parts = true;
  return NRP_OK;
}

}  // extern "C"

namespace {

template <typename KeyT>
int build_impl(nrp_ctx* c, int k, int internal_repeats, int want_edges) {
  TRY(configure_kernels<KeyT>(c));
  cudaStream_t st = c->stream;
  const int64_t n_parts = c->n_parts;
  const Parts P = parts_view(c);

  int64_t m_max = 0;
  if (c->uniform_len > 0) m_max = c->uniform_len >= k ? n_parts * (int64_t)(c->uniform_len - k + 1) : 0;
  else for (int64_t i = 0; i < n_parts; ++i) { int64_t w = c->h_off[i + 1] - c->h_off[i] - k + 1; if (w > 0) m_max += w; }
  if (m_max >= 0xffffe000ll) return fail(NRP_E_UNSUPPORTED, "%lld records exceed the 32-bit record index of one device", (long long)m_max);

  memset(c->counts, 0, sizeof(c->counts));
  launch_counter() = 0;
  memset(c->timings, 0, sizeof(c->timings));
  c->counts[NRP_C_WINDOWS_MAX] = m_max;
  c->counts[NRP_C_KEY_BYTES] = sizeof(KeyT);
  c->counts[NRP_C_EDGES] = want_edges ? 0 : -1;
  c->k = k; c->key_bytes = sizeof(KeyT);

  TRY(c->flags.ensure((size_t)n_parts + 16));
  TRY(c->last_single.ensure(sizeof(int32_t) * (size_t)(n_parts + 1)));
  const size_t rec = (size_t)m_max + 16;
  TRY(c->keysA.ensure(sizeof(KeyT) * rec)); TRY(c->valsA.ensure(4 * rec));
  TRY(c->keysB.ensure(sizeof(KeyT) * rec)); TRY(c->valsB.ensure(4 * rec));
  const size_t leaf_max = (size_t)1 << kMaxLeafBits;
  TRY(c->hist.ensure(4 * (leaf_max + 2))); TRY(c->leaf_off.ensure(4 * (leaf_max + 2)));
  TRY(c->cursor1.ensure(4 * 256)); TRY(c->cursor2.ensure(4 * (leaf_max + 2)));
  TRY(c->seg_off.ensure(4 * 260)); TRY(c->tile_prefix.ensure(4 * 260));
  TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems((int64_t)leaf_max + 2)));
  unsigned long long* cnt = c->counters.as<unsigned long long>();

  StageTimer tm(c);
  tm.begin(NRP_T_FLAGS);
  CU(cudaMemsetAsync(c->flags.ptr, 0, (size_t)n_parts + 16, st));
  if (c->n_tiles > 0 && m_max >= 0) {
    if (!internal_repeats && (k % 2) == 0)
      NRP_LAUNCH((k_flag_windows<uint32_t><<<(unsigned)c->n_tiles, kTileThreads, 0, st>>>(P, k, 0, nullptr, 0, c->flags.as<uint8_t>())));
    if (c->bg_n > 0) {
      if (c->bg_key_bytes == 4)
        NRP_LAUNCH((k_flag_windows<uint32_t><<<(unsigned)c->n_tiles, kTileThreads, 0, st>>>(P, c->bg_K, 1, c->bg_table.as<uint32_t>(), c->bg_log_slots,
                                                                              c->flags.as<uint8_t>())));
      else
        NRP_LAUNCH((k_flag_windows<uint64_t><<<(unsigned)c->n_tiles, kTileThreads, 0, st>>>(P, c->bg_K, 1, c->bg_table.as<uint64_t>(), c->bg_log_slots,
                                                                              c->flags.as<uint8_t>())));
    }
  }
  CU(cudaGetLastError());

  const int bits = choose_leaf_bits(c, m_max);
  const int r1 = std::min<int>(bits, (int)c->opt_max_level_bits);
  const int r2 = bits - r1;
  const uint32_t n_leaf = 1u << bits;
  c->counts[NRP_C_BUCKETS] = n_leaf;
  int part_bits = 1;
  while (((int64_t)1 << part_bits) < n_parts) ++part_bits;

  int64_t n_groups = 0, n_members = 0, n_pairs = 0, n_cand = 0, n_records = 0;
  KeyT* sk = nullptr; uint32_t* sv = nullptr;
  {
    c->counts[NRP_C_RUNS] = 1;
    tm.begin(NRP_T_HISTOGRAM);
    CU(cudaMemsetAsync(c->hist.ptr, 0, 4 * ((size_t)n_leaf + 2), st));
    CU(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * CNT_N, st));
    if (m_max > 0) {
      const unsigned hist_grid = (unsigned)std::min<int64_t>(c->n_tiles, (int64_t)c->n_sms * (bits >= 15 ? 1 : 2));
      NRP_LAUNCH((k_hist_from_ascii<<<hist_grid, kTileThreads, 4u << bits, st>>>(P, c->flags.as<uint8_t>(), k, bits, c->n_tiles, c->hist.as<uint32_t>())));
    }
    exclusive_scan<uint32_t, uint32_t>(c->hist.as<uint32_t>(), c->leaf_off.as<uint32_t>(), n_leaf, c->scan_scratch.as<uint32_t>(), st);
    NRP_LAUNCH((k_init_cursors<<<grid_for(n_leaf, 256), 256, 0, st>>>(c->leaf_off.as<uint32_t>(), r1, r2, c->cursor1.as<uint32_t>(), c->cursor2.as<uint32_t>())));

    tm.begin(NRP_T_SPLIT1);
    KeyT* leaf_keys = c->keysA.as<KeyT>(); uint32_t* leaf_vals = c->valsA.as<uint32_t>();
    KeyT* free_keys = c->keysB.as<KeyT>(); uint32_t* free_vals = c->valsB.as<uint32_t>();
    if (m_max > 0)
      NRP_LAUNCH((k_split_from_ascii<KeyT><<<(unsigned)c->n_tiles, kTileThreads, sizeof(SplitSmem<KeyT>) + sizeof(TileStage), st>>>(
          P, c->flags.as<uint8_t>(), k, r1, 0, c->part_base, r2 > 0 ? c->cursor1.as<uint32_t>() : c->cursor2.as<uint32_t>(), leaf_keys, leaf_vals)));
    if (r2 > 0 && m_max > 0) {
      tm.begin(NRP_T_SPLIT2);
      NRP_LAUNCH((k_segment_tiles<<<1, 256, 0, st>>>(c->leaf_off.as<uint32_t>(), r1, r2, c->seg_off.as<uint32_t>(), c->tile_prefix.as<uint32_t>())));
      const unsigned grid2 = (unsigned)std::min<int64_t>(m_max / kSplitTile + (1 << r1), (int64_t)c->n_sms * 4);
      NRP_LAUNCH((k_split_from_records<KeyT><<<grid2, kSplitThreads, sizeof(SplitSmem<KeyT>), st>>>(
          leaf_keys, leaf_vals, c->seg_off.as<uint32_t>(), c->tile_prefix.as<uint32_t>(), r1, r2, c->cursor2.as<uint32_t>(), free_keys, free_vals)));
      std::swap(leaf_keys, free_keys); std::swap(leaf_vals, free_vals);
    }
    tm.begin(NRP_T_FILTER);
    if (m_max > 0) {
      const size_t fsm = (sizeof(KeyT) + 1) * 2 * (size_t)kFilterCap;
      const unsigned fgrid = (unsigned)std::min<int64_t>(n_leaf, (int64_t)c->n_sms * (sizeof(KeyT) == 4 ? 2 : 1));
      NRP_LAUNCH((k_filter<KeyT, kFilterThreads, kFilterItems><<<fgrid, kFilterThreads, fsm, st>>>(
          leaf_keys, leaf_vals, c->leaf_off.as<uint32_t>(), n_leaf, bits, (uint32_t)c->opt_filter_cap,
          reinterpret_cast<uint32_t*>(cnt + CNT_WORK), free_keys, free_vals, cnt + CNT_CAND, reinterpret_cast<uint32_t*>(cnt + CNT_OVERSIZED))));
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_counters, cnt, sizeof(unsigned long long) * CNT_N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->h_counters + 16, c->leaf_off.as<uint32_t>() + n_leaf, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    n_cand = (int64_t)c->h_counters[CNT_CAND];
    n_records = (int64_t)(uint32_t)c->h_counters[16];
    c->counts[NRP_C_OVERSIZED] = (int64_t)(uint32_t)c->h_counters[CNT_OVERSIZED];

    // candidates live in (free_keys, free_vals); the leaf buffers are free again
    int64_t n_grp = 0, n_repeat = 0;
    TRY((group_candidates<KeyT>(c, free_keys, free_vals, leaf_keys, leaf_vals, n_cand, 2 * k, part_bits, n_parts, !internal_repeats, tm,
                                &sk, &sv, &n_grp, &n_repeat)));
    if (!internal_repeats && n_repeat > 0) {
      // windows of the parts just flagged were counted as records: take them out again
      CU(cudaMemsetAsync(cnt + CNT_WORK, 0, 8, st));
      NRP_LAUNCH((k_flagged_windows<<<grid_for(n_parts, 256), 256, 0, st>>>(c->off.as<int64_t>(), c->flags.as<uint8_t>(), n_parts, k, cnt + CNT_WORK)));
      CU(cudaMemcpyAsync(c->h_counters + 27, cnt + CNT_WORK, 8, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n_records -= (int64_t)c->h_counters[27];
    }
    int64_t n_runs_bound = 0;
    if (n_grp > 0) {
      CU(cudaMemcpyAsync(c->h_counters + 17, c->head_scan.as<uint32_t>() + n_grp, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n_runs_bound = (int64_t)(uint32_t)c->h_counters[17];
      NRP_LAUNCH((k_group_sizes<<<grid_for(n_runs_bound, 256), 256, 0, st>>>(c->group_start.as<uint32_t>(), n_runs_bound, c->valid.as<uint32_t>(),
                                                                c->mem_cnt.as<uint32_t>(), c->pair_cnt.as<unsigned long long>())));
      exclusive_scan<uint32_t, uint32_t>(c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(), n_runs_bound, c->scan_scratch.as<uint32_t>(), st);
      exclusive_scan<uint32_t, unsigned long long>(c->mem_cnt.as<uint32_t>(), c->mem_scan.as<unsigned long long>(), n_runs_bound,
                                                   c->scan_scratch.as<unsigned long long>(), st);
      exclusive_scan<unsigned long long, unsigned long long>(c->pair_cnt.as<unsigned long long>(), c->pair_scan.as<unsigned long long>(),
                                                             n_runs_bound, c->scan_scratch.as<unsigned long long>(), st);
      CU(cudaMemcpyAsync(c->h_counters + 20, c->valid_scan.as<uint32_t>() + n_runs_bound, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(c->h_counters + 21, c->mem_scan.as<unsigned long long>() + n_runs_bound, 8, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(c->h_counters + 22, c->pair_scan.as<unsigned long long>() + n_runs_bound, 8, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n_groups = (int64_t)(uint32_t)c->h_counters[20];
      n_members = (int64_t)c->h_counters[21];
      n_pairs = (int64_t)c->h_counters[22];
      TRY(c->g_keys.ensure(8 * (size_t)(n_groups + 1))); TRY(c->g_indptr.ensure(8 * (size_t)(n_groups + 2)));
      TRY(c->g_pair_off.ensure(8 * (size_t)(n_groups + 2))); TRY(c->g_members.ensure(4 * (size_t)(n_members + 1)));
      NRP_LAUNCH((k_group_emit<KeyT><<<grid_for(n_runs_bound + 1, 256), 256, 0, st>>>(
          sk, c->group_start.as<uint32_t>(), c->valid.as<uint32_t>(), c->valid_scan.as<uint32_t>(), c->mem_scan.as<unsigned long long>(),
          c->pair_scan.as<unsigned long long>(), n_runs_bound, c->g_keys.as<uint64_t>(), c->g_indptr.as<int64_t>(),
          c->g_pair_off.as<unsigned long long>())));
      NRP_LAUNCH((k_group_members<<<grid_for(n_grp, 256), 256, 0, st>>>(sv, c->head_scan.as<uint32_t>(), c->head.as<uint32_t>(), c->group_start.as<uint32_t>(),
                                                             c->valid.as<uint32_t>(), c->mem_scan.as<unsigned long long>(), n_grp,
                                                             c->g_members.as<uint32_t>())));
      CU(cudaGetLastError());
    } else {
      n_groups = n_members = n_pairs = 0;
      TRY(c->g_keys.ensure(8)); TRY(c->g_indptr.ensure(16)); TRY(c->g_pair_off.ensure(16)); TRY(c->g_members.ensure(4));
      CU(cudaMemsetAsync(c->g_indptr.ptr, 0, 16, st));
      CU(cudaMemsetAsync(c->g_pair_off.ptr, 0, 16, st));
    }
  }
  c->counts[NRP_C_RECORDS] = n_records;
  c->counts[NRP_C_CANDIDATES] = n_cand;
  c->counts[NRP_C_GROUPS] = n_groups;
  c->counts[NRP_C_MEMBERS] = n_members;
  c->counts[NRP_C_PAIRS] = n_pairs;

  // ---- order information for the host replay
  tm.begin(NRP_T_ORDER);
  TRY(c->g_first_part.ensure(4 * (size_t)(n_groups + 1))); TRY(c->g_first_window.ensure(4 * (size_t)(n_groups + 1)));
  if (n_groups > 0) {
    NRP_LAUNCH((k_gather_first_part<<<grid_for(n_groups, 256), 256, 0, st>>>(c->g_indptr.as<int64_t>(), c->g_members.as<uint32_t>(), n_groups,
                                                                 c->g_first_part.as<uint32_t>())));
    NRP_LAUNCH((k_group_first_window<<<grid_for(n_groups, 128), 128, 0, st>>>(c->ascii.as<uint8_t>(), c->off.as<int64_t>(), c->g_keys.as<uint64_t>(),
                                                                  c->g_first_part.as<uint32_t>(), c->part_base, n_parts, n_groups, k,
                                                                  c->g_first_window.as<uint32_t>())));
  }
  if (n_parts > 0)
    NRP_LAUNCH((k_last_single<<<grid_for(n_parts, 128), 128, 0, st>>>(c->ascii.as<uint8_t>(), c->off.as<int64_t>(), c->flags.as<uint8_t>(), n_parts, k,
                                                          c->g_keys.as<uint64_t>(), n_groups, c->last_single.as<int32_t>())));
  CU(cudaGetLastError());

  // ---- edges
  if (want_edges) {
    tm.begin(NRP_T_EDGES);
    if (!c->have_ids) return fail(NRP_E_ARG, "want_edges needs part ids");
    if (n_pairs > c->opt_max_pairs) return fail(NRP_E_EDGES, "%lld member pairs exceed the limit of %lld", (long long)n_pairs, (long long)c->opt_max_pairs);
    int64_t n_edges = 0;
    if (n_pairs > 0) {
      TRY(c->e_codesA.ensure(8 * (size_t)(n_pairs + 1))); TRY(c->e_codesB.ensure(8 * (size_t)(n_pairs + 1)));
      TRY(c->e_head.ensure(4 * (size_t)(n_pairs + 2))); TRY(c->e_head_scan.ensure(4 * (size_t)(n_pairs + 2)));
      TRY(c->lsd_scratch.ensure(sizeof(uint32_t) * (size_t)lsd_scratch_elems(n_pairs)));
      TRY(c->scan_scratch.ensure(8 * (size_t)scan_scratch_elems(n_pairs + 2)));
      NRP_LAUNCH((k_edge_pairs<<<grid_for(n_members, 128), 128, 0, st>>>(c->g_indptr.as<int64_t>(), c->g_members.as<uint32_t>(),
                                                             c->g_pair_off.as<unsigned long long>(), n_groups, n_members,
                                                             c->part_id.as<uint32_t>(), c->part_base, c->e_codesA.as<uint64_t>())));
      LsdPlan plan; plan.n_passes = 0;
      lsd_plan_add(plan, 0, 0, 64);
      uint64_t* sorted_codes;
      lsd_sort<uint64_t, false>(c->e_codesA.as<uint64_t>(), nullptr, c->e_codesB.as<uint64_t>(), nullptr, n_pairs, plan,
                                c->lsd_scratch.as<uint32_t>(), st, &sorted_codes, nullptr);
      NRP_LAUNCH((k_edge_heads<<<grid_for(n_pairs, 256), 256, 0, st>>>(sorted_codes, n_pairs, c->e_head.as<uint32_t>())));
      exclusive_scan<uint32_t, uint32_t>(c->e_head.as<uint32_t>(), c->e_head_scan.as<uint32_t>(), n_pairs, c->scan_scratch.as<uint32_t>(), st);
      CU(cudaMemcpyAsync(c->h_counters + 24, c->e_head_scan.as<uint32_t>() + n_pairs, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      n_edges = (int64_t)(uint32_t)c->h_counters[24];
      TRY(c->e_u.ensure(4 * (size_t)(n_edges + 1))); TRY(c->e_v.ensure(4 * (size_t)(n_edges + 1)));
      NRP_LAUNCH((k_edge_compact<<<grid_for(n_pairs, 256), 256, 0, st>>>(sorted_codes, c->e_head.as<uint32_t>(), c->e_head_scan.as<uint32_t>(), n_pairs,
                                                             c->e_u.as<uint32_t>(), c->e_v.as<uint32_t>())));
      CU(cudaGetLastError());
    }
    c->counts[NRP_C_EDGES] = n_edges;
  }
  tm.begin(NRP_T_GROUP);
  CU(cudaMemsetAsync(cnt + CNT_FLAGGED, 0, 8, st));
  if (n_parts > 0) NRP_LAUNCH((k_count_flagged<<<grid_for(n_parts, 256), 256, 0, st>>>(c->flags.as<uint8_t>(), n_parts, cnt + CNT_FLAGGED)));
  CU(cudaMemcpyAsync(c->h_counters + 26, cnt + CNT_FLAGGED, 8, cudaMemcpyDeviceToHost, st));
  tm.finish(c->timings);
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  c->counts[NRP_C_EXCLUDED] = (int64_t)c->h_counters[26];
  c->counts[NRP_C_LAUNCHES] = launch_counter();
  double total = 0;
  for (int i = NRP_T_FLAGS; i < NRP_N_TIMINGS; ++i) total += c->timings[i];
  c->timings[NRP_T_TOTAL] = total;
  c->have_result = true;
  return NRP_OK;
}

}  // namespace

extern "C" {

int nrp_build(nrp_ctx* c, int k, int internal_repeats, int want_edges) {
  if (!c) return fail(NRP_E_ARG, "null context");
  if (!c->have_parts) return fail(NRP_E_STATE, "nrp_load_parts has not been called");
  if (k < 1 || k > 32) return fail(NRP_E_UNSUPPORTED, "k=%d outside [1, 32]", k);
  CU(cudaSetDevice(c->device));
  c->have_result = false;
  return k <= 16 ? build_impl<uint32_t>(c, k, internal_repeats, want_edges) : build_impl<uint64_t>(c, k, internal_repeats, want_edges);
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
