/*
 * nrp_b200.h -- C ABI of the B200-native Finder-mode repeat-graph builder.
 *
 * This library replaces the graph-building internals of ayaanhossain/nrpcalc's Finder mode:
 *
 *   reference nrpcalc/base/hgraph.py:23-94    get_repeat_dict      (k-mer -> part-id table)      -> nrp_build
 *   reference nrpcalc/base/hgraph.py:117-135  get_repeat_cliques   (distinct id tuples)           -> nrp_result_groups / nrp_result_last_single
 *   reference nrpcalc/base/hgraph.py:157-200  build_homology_graph (part-pair conflict edges)     -> nrp_result_edges
 *   reference nrpcalc/base/utils.py:37-50     stream_kmers / get_revcomp / stream_min_kmers       -> device code of nrp_build
 *   reference nrpcalc/base/kmerSetDB.py:177-207  __contains__ / multicheck (background membership) -> nrp_bg_set + nrp_build
 *
 * The reference has no FFI; the seam a maintainer would bind is the Python function
 * hgraph.get_homology_graph (hgraph.py:202-249).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success or a negative NRP_E_* code, and
 *     nrp_last_error() returns a thread-local, library-owned message for the last failure;
 *   - parts are passed in the reference's PROCESSING ORDER (descending last occurrence of the unique,
 *     upper-cased strings, hgraph.py:11-21 and 218-223) as one ASCII buffer + int64 offsets[n_parts+1];
 *     a part is identified on the device by its index in that order;
 *   - k = Lmax + 1 (reference nrpcalc/nrpcalc.py:385), 1 <= k <= 32; characters must be A/C/G/T in
 *     either case (2-bit alphabet A=0 C=1 G=2 T=3, most significant base first, so that integer
 *     order equals the reference's string order); anything else must be rejected by the caller;
 *   - only windows of parts with length >= k are formed here; the host layer submits the parts shorter
 *     than k (whole-string keys, utils.py:37-40) as separate builds with k = their length;
 *   - inputs are borrowed for the duration of the call; results stay owned by the context until the
 *     next nrp_build/nrp_ctx_destroy and are copied into caller-allocated buffers by the getters;
 *   - one context per device and per host thread; calls on one context must not overlap.
 */
#ifndef NRP_B200_H
#define NRP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRP_OK              0
#define NRP_E_ARG          -1   /* invalid argument */
#define NRP_E_CUDA         -2   /* CUDA runtime error */
#define NRP_E_OOM          -3   /* device allocation failed */
#define NRP_E_UNSUPPORTED  -4   /* k > 32, more than 2^32-2 records, ... */
#define NRP_E_STATE        -5   /* getter called without a finished build */
#define NRP_E_EDGES        -6   /* part-pair expansion exceeds the configured limit */

/* per-part exclusion bits (reference hgraph.py:50-79) */
#define NRP_FLAG_REPEAT      1  /* two windows share a canonical k-mer (direct or inverted internal repeat) */
#define NRP_FLAG_PALINDROME  2  /* a window equals its own reverse complement */
#define NRP_FLAG_BACKGROUND  4  /* a canonical K-window is a background key */

/* indices into nrp_result_counts() */
#define NRP_C_RECORDS        0  /* (window, part) records of surviving parts that were grouped */
#define NRP_C_EXCLUDED       1  /* parts with a non-zero flag */
#define NRP_C_GROUPS         2  /* canonical keys with >= 2 members */
#define NRP_C_MEMBERS        3  /* total members of those groups */
#define NRP_C_EDGES          4  /* unique part-pair edges (-1 when edges were not requested) */
#define NRP_C_PAIRS          5  /* member pairs expanded before dedupe */
#define NRP_C_CANDIDATES     6  /* records that survived the in-bucket duplicate filter */
#define NRP_C_RUNS           7  /* passes over slices of the hash space (1 unless the batch exceeds pass_records) */
#define NRP_C_BUCKETS        8  /* leaf buckets of the radix partition */
#define NRP_C_OVERSIZED      9  /* leaf buckets larger than the shared-memory filter capacity */
#define NRP_C_WINDOWS_MAX   10  /* windows of all parts before exclusion */
#define NRP_C_KEY_BYTES     11  /* 4 or 8 */
#define NRP_C_LAUNCHES      12  /* kernels launched by the build */
#define NRP_C_DIRECT        13  /* 1: buckets were placed in fixed regions (no histogram pass); 0: exact histogram path */
#define NRP_C_FALLBACKS     14  /* a fixed region overflowed and the batch was repeated on the exact path */
#define NRP_C_PACKED        15  /* 1: records travelled as single 64-bit words (k-mer << value bits) | value */
#define NRP_C_UNIFORM       16  /* 1: the part-aligned extraction kernel ran (parts of one length, packed records); 2: it also computed the palindrome / background flags */
#define NRP_C_FILTER_WARP   17  /* 1 / 2 / 3: the warp-private duplicate filter ran (leaves of <= 800 / <= 400 records on average / room for 192 suspects) */
#define NRP_N_COUNTS        20

/* indices into nrp_result_timings() (milliseconds, CUDA events on the context's stream) */
#define NRP_T_TOTAL          0  /* device-resident inputs -> device-resident results (T_build) */
#define NRP_T_H2D            1
#define NRP_T_FLAGS          2
#define NRP_T_HISTOGRAM      3
#define NRP_T_SPLIT1         4
#define NRP_T_SPLIT2         5
#define NRP_T_FILTER         6
#define NRP_T_SORT           7
#define NRP_T_GROUP          8
#define NRP_T_EDGES          9
#define NRP_T_ORDER         10  /* first-window ranks + last single window */
#define NRP_N_TIMINGS       16

typedef struct nrp_ctx nrp_ctx;

int         nrp_version(void);
const char* nrp_last_error(void);
/* Kernels this library has launched from the calling thread since it was loaded (never reset; NRP_C_LAUNCHES counts one build). */
int64_t     nrp_launches_total(void);

/* Bind a context to a CUDA device (creates a stream; device buffers grow on demand and are reused). */
int nrp_ctx_create(int device, nrp_ctx** out);
int nrp_ctx_destroy(nrp_ctx* ctx);
int nrp_ctx_device(nrp_ctx* ctx);

/* Tuning knobs used by the tests to force rarely taken paths: name in
 * {"filter_cap", "leaf_target", "max_level_bits", "max_pairs", "no_window_carry", "no_bloom", "pass_records", "direct", "packed", "persistent_split", "r1_bias", "nvlink_digits", "coop_group", "filter_tiny"};
 * value <= 0 restores the default ("direct": 0 forces the exact histogram path, negative restores the default 1;
 * "coop_group": 0 = grouping as a kernel sequence, 1 = one cooperative kernel up to 2^20 candidates (default), 2 = always). */
int nrp_ctx_set_option(nrp_ctx* ctx, const char* name, int64_t value);

/* Background (replaces the kmerSetDB membership test, kmerSetDB.py:177-207): n canonical K-mers in the
 * 2-bit encoding, host memory, any order, duplicates allowed.  A part is excluded iff K <= len(part) and any
 * canonical K-window of it is in the set (hgraph.py:72-79).  n = 0 or nrp_bg_clear() disables the filter. */
int nrp_bg_set(nrp_ctx* ctx, const uint64_t* canon_kmers, int64_t n, int K);
int nrp_bg_clear(nrp_ctx* ctx);

/* Background BUILD on the device (replaces the one-put-per-k-mer loop of kmerSetDB.add / multiadd, kmerSetDB.py:129-175, and is
 * what the persistent store mirrors): the canonical K-mers of every K-window of n_seqs sequences over A/C/G/T (either case;
 * sequences shorter than K contribute nothing here -- the host layer keeps their whole-string keys), sorted ascending, duplicates
 * removed.  *n_unique receives the number of distinct keys; nrp_bg_build_fetch copies them to the host.  install != 0 also makes
 * them this context's probe set (as nrp_bg_set would). */
int nrp_bg_build(nrp_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int64_t n_seqs, int K, int install, int64_t* n_unique);
int nrp_bg_build_fetch(nrp_ctx* ctx, uint64_t* keys_out);

/* Device-side uniquify of a packed part list (replaces uniquify_seq_list, hgraph.py:11-21, and the processing order of
 * hgraph.py:218-223): n_parts parts as one ASCII buffer (host) + int64 offsets[n_parts+1], or offsets == NULL and every part
 * part_len bytes long.  The parts are upper-cased, exact duplicates are dropped, the survivors are ordered by DESCENDING index of
 * their last occurrence and gathered into a device buffer.  out[0] = number of unique parts, out[1] = their total bytes,
 * out[2] = status bits (1: a character outside A/C/G/T occurs; 2: two different parts collided in the 64-bit hash -- the result
 * must not be used, take the host path), out[3] = longest part.  nrp_uniquify_fetch copies, for processing rank q,
 * ids[q] = index of the FIRST occurrence (the part's id), source[q] = index of the last occurrence, offsets[0..n_unique] of the
 * gathered buffer (any of them may be NULL), and returns the gathered buffer's device address (valid until the next nrp_uniquify /
 * nrp_ctx_destroy) -- pass it to nrp_load_parts with ascii_on_device = 1. */
int nrp_uniquify(nrp_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int part_len, int64_t n_parts, int64_t out[4]);
int nrp_uniquify_fetch(nrp_ctx* ctx, uint32_t* ids, uint32_t* source, int64_t* offsets, void** dev_ascii);

/* Stage one batch of parts on the device (H2D copy unless ascii_on_device != 0; offsets and part_id are
 * always host arrays).
 *   ascii, offsets   concatenated parts and their n_parts+1 byte offsets, offsets[0] == 0
 *   part_id          id of each part (first index in the caller's list); may be NULL when edges are never requested */
int nrp_load_parts(nrp_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, const uint32_t* part_id, int64_t n_parts,
                   int ascii_on_device);

/* Streamed staging for page-locked host buffers: like nrp_load_parts, but the ASCII copy is only ENQUEUED, in n_chunks pieces
 * (contiguous part ranges) on a second stream, and the next nrp_build extracts every piece as soon as it has landed, so the
 * host-to-device copy overlaps the extraction.  `ascii` must stay valid and unchanged until that nrp_build has returned. */
int nrp_load_parts_streamed(nrp_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, const uint32_t* part_id, int64_t n_parts,
                            int n_chunks);

/* Parts that all have the same length (the BASELINE shapes; the host layer knows this from joining the strings): no offsets
 * array -- part i is ascii[i * part_len, (i + 1) * part_len) and the offsets are generated on the device -- so nothing is swept
 * or staged on the host and the call returns as soon as the copies are enqueued (n_chunks > 1, page-locked `ascii`; `part_id`
 * is copied from where it lies when it is page-locked too, else through the library's staging buffer).  The next nrp_build
 * extracts each piece as it lands: the host-to-device copy overlaps the first split.  n_chunks == 1 is a blocking load.
 * Lifetime of `ascii` (and of a page-locked `part_id`) as for nrp_load_parts_streamed. */
int nrp_load_parts_uniform(nrp_ctx* ctx, const uint8_t* ascii, int64_t n_parts, int part_len, const uint32_t* part_id, int n_chunks);

/* Build the repeat structure of the staged parts.
 *   internal_repeats 0: exclude parts with internal direct/inverted repeats or palindromic windows
 *   want_edges       != 0: also expand groups into the distinct (id_u < id_v) edges
 * Synchronous: returns when the results are resident on the device.  May be called repeatedly. */
int nrp_build(nrp_ctx* ctx, int k, int internal_repeats, int want_edges);

/* The same build for parts over ANY byte alphabet and for ANY k >= 1 (the reference has neither limit: utils.py:10 maps 'ATGCU'
 * to 'TACGA' and leaves every other character alone -- U's complement is A but A's is T --, keys are compared as strings,
 * hgraph.py:23-94).  Windows are keyed by a seeded 64-bit hash of the canonical string min(w, rc(w)); every equality that changes
 * a result is verified on the bytes, so the result is exact: *status != 0 reports a hash collision between different strings
 * (no result was produced -- call again with another seed).  Internal repeats follow hgraph.py:50-69 literally (two equal forward
 * windows, or the reverse complement of a window among the forward windows).  The background probe applies U -> T first
 * (kmerSetDB.py:183-184) and takes windows over A/C/G/T; preset_flags (host, n_parts bytes, or NULL) carries exclusions the
 * host layer decided itself (background keys it holds as strings, K > 32) -- parts with a non-zero preset flag produce no record.
 * The parts must be upper-case already.  Limits: at most 4096 windows per part, (part, window) must fit 32 bits.  seed: the low
 * 58 bits seed the hash; the top 6 bits must be 0 (a non-zero value b keeps only b hash bits -- for tests of the collision path).
 * Results: the same getters as after nrp_build; group keys are hashes (meaningful only for equality). */
int nrp_build_generic(nrp_ctx* ctx, int k, int internal_repeats, int want_edges, uint64_t seed, const uint8_t* preset_flags, int* status);

int nrp_result_counts(nrp_ctx* ctx, int64_t out[NRP_N_COUNTS]);
int nrp_result_timings(nrp_ctx* ctx, double out[NRP_N_TIMINGS]);
/* flags[n_parts]: NRP_FLAG_* bits per part */
int nrp_result_flags(nrp_ctx* ctx, uint8_t* flags);
/* groups with >= 2 members as CSR: indptr[n_groups+1], members[n_members] = part indices in (part, window)
 * order with multiplicity; first_window[g] = smallest window index of members[indptr[g]] whose canonical
 * k-mer is the group's key, i.e. the group's rank is (members[indptr[g]], first_window[g]);
 * keys[g] = the canonical k-mer (may be NULL). */
int nrp_result_groups(nrp_ctx* ctx, int64_t* indptr, uint32_t* members, uint32_t* first_window, uint64_t* keys);
/* last_single[n_parts]: largest window index whose canonical k-mer no other record shares, -1 if none
 * (and -1 for excluded parts) */
int nrp_result_last_single(nrp_ctx* ctx, int32_t* last_single);
/* u[n_edges] < v[n_edges], every conflicting id pair exactly once, order unspecified */
int nrp_result_edges(nrp_ctx* ctx, uint32_t* u, uint32_t* v);

/* All results of the last build in ONE call: the library copies them into its own page-locked host buffers (full-speed DMA,
 * one synchronisation) and returns pointers that stay valid until the next nrp_load_parts / nrp_build / nrp_ctx_destroy on this
 * context.  keys is NULL unless want_keys != 0; edge_u / edge_v are NULL when edges were not requested.
 * With the option early_fetch set (nrp_ctx_set_option; for callers that go on to fetch), a single-GPU nrp_build already starts
 * the copies of everything that is final after its grouping stage (members, group offsets, first windows, keys, flags) beside the
 * edge expansion; this call then moves only the rest. */
typedef struct nrp_host_result {
  const uint8_t*  flags;         /* n_parts */
  const int64_t*  indptr;        /* n_groups + 1 */
  const uint32_t* members;       /* n_members */
  const uint32_t* first_window;  /* n_groups */
  const uint64_t* keys;          /* n_groups */
  const int32_t*  last_single;   /* n_parts */
  const uint32_t* edge_u;        /* n_edges */
  const uint32_t* edge_v;        /* n_edges */
  int64_t n_parts, n_groups, n_members, n_edges;
} nrp_host_result;
int nrp_result_host(nrp_ctx* ctx, int want_keys, nrp_host_result* out);

/* ---- sharded path: one process per GPU, k-mer space hash-partitioned across n_owners ranks ---------------------
 * Every rank stages ALL parts with nrp_load_parts (record values are then global part indices and any rank can
 * read any part), extracts the records of its own contiguous shard of parts, and owns the k-mers with
 * owner_hash(k-mer) % n_owners == rank.  Between the calls the host layer moves data with torch.distributed:
 *
 *   nrp_shard_extract      flags + records of parts [part_lo, part_hi), partitioned by owner; counts per owner
 *   (all-to-all of the counts and of the records: NCCL over NVLink)
 *   nrp_shard_group        partition + filter + sort of the records this rank owns; marks internal repeats in flags
 *   (all-reduce MAX of the per-part flags, nrp_flags_export / nrp_flags_import)
 *   nrp_shard_finish       drops records of flagged parts, builds this rank's groups (+ first windows, + local edges)
 *   (all-gather of the shared k-mers)
 *   nrp_shard_last_single  last unshared window of the parts of this rank's shard
 */
/* Fused exchange (preferred when the ranks' GPUs have peer access): instead of an NCCL all-to-all, the extraction kernel
 * stores every record straight into its owner's receive buffer over NVLink.
 *   nrp_peer_export          allocate this rank's receive buffers (capacity in records) and return their CUDA IPC handles
 *   nrp_peer_open            open all ranks' handles (n_peers x 128 bytes: keys handle, values handle); own rank = local memory
 *   nrp_shard_count          flags + number of records this rank will send to every owner
 *   (host: all-gather of the counts -> write offsets; barrier)
 *   nrp_shard_scatter_peers  extraction + scatter into the owners' buffers at write_offsets[owner]
 *   (host: barrier)
 *   nrp_shard_group_received = nrp_shard_group on the first n_records of this rank's receive buffers */
int nrp_peer_export(nrp_ctx* ctx, int64_t capacity_records, int k, unsigned char* handle_keys /*64*/, unsigned char* handle_vals /*64*/);
int nrp_peer_open(nrp_ctx* ctx, int n_peers, int my_rank, const unsigned char* handles);
int nrp_shard_count(nrp_ctx* ctx, int64_t part_lo, int64_t part_hi, int k, int internal_repeats, int n_owners, int64_t* owner_counts);
int nrp_shard_scatter_peers(nrp_ctx* ctx, int k, const int64_t* write_offsets);
int nrp_shard_group_received(nrp_ctx* ctx, int64_t n_records, int k, int internal_repeats);

/* Direct fused exchange (default): no count pass and no count exchange.  Every (level-1 bucket, sender) pair owns a FIXED
 * region of the owner's receive buffer; the extraction kernel stores each record into region (bucket * n_owners + sender) of
 * its owner over NVLink and finally delivers the regions' end positions.  The owner then continues with its level-2 split
 * straight from the received regions.  All ranks must call with the same windows_total (k-windows of ALL parts).
 *   nrp_peer_export_direct    size + allocate the receive buffers for (windows_total, n_owners); three IPC handles
 *   nrp_peer_open_direct      open all ranks' handles (n_peers x 192 bytes: keys, values, ends)
 *   nrp_shard_scatter_direct  flags + extraction + scatter into the owners' regions; *overflow != 0: some region was too
 *                             small (heavily duplicated input) -- every rank must then fall back to the counted exchange
 *   (host: all-reduce MAX of the overflow flags = barrier)
 *   nrp_shard_group_direct    level-2 split + filter + ordering of the records this rank received */
int nrp_peer_export_direct(nrp_ctx* ctx, int64_t windows_total, int n_owners, int k, unsigned char* handles /*192*/);
int nrp_peer_open_direct(nrp_ctx* ctx, int n_peers, int my_rank, const unsigned char* handles /* n_peers x 192 */);
int nrp_shard_scatter_direct(nrp_ctx* ctx, int64_t part_lo, int64_t part_hi, int k, int internal_repeats, int n_owners, int* overflow);
int nrp_shard_group_direct(nrp_ctx* ctx, int k, int internal_repeats);
/* out[0] = records this rank sent in the last nrp_shard_scatter_direct, out[1] = records it received and partitioned */
int nrp_shard_moved(nrp_ctx* ctx, int64_t out[2]);
/* Shard-only staging (multi-GPU end to end): every rank copies only the bytes of ITS shard [stage_lo, stage_hi) to the device
 * -- 1/N of the host-to-device traffic -- while the device buffer, the offsets, the ids and the record values stay GLOBAL (a
 * part keeps its global index and byte offset on every rank).  offsets == NULL: parts of one length part_len.  The one place
 * that needs a foreign part's bytes (the first window of a shared k-mer when records do not carry the window index) reads them
 * from the home rank's buffer over NVLink: nrp_peer_export_parts / nrp_peer_open_parts exchange the CUDA IPC handles of the
 * part buffers and the shard bounds (part_bounds[r] .. part_bounds[r + 1] live on rank r).  nrp_build refuses a context
 * that holds only a shard.  Replaces nothing in the reference (it has no distributed mode); host side: sharded.py. */
int nrp_load_parts_shard(nrp_ctx* ctx, const uint8_t* ascii, const int64_t* offsets, int part_len, const uint32_t* part_id, int64_t n_parts,
                         int64_t stage_lo, int64_t stage_hi, int n_chunks);
/* bytes the part buffer can hold without being re-allocated (a batch that needs more must not be loaded while peers map the
 * buffer: nrp_peer_close on every rank first) */
int nrp_parts_capacity(nrp_ctx* ctx, int64_t* bytes);
int nrp_peer_export_parts(nrp_ctx* ctx, unsigned char* handle /*64*/);
int nrp_peer_open_parts(nrp_ctx* ctx, int n_peers, int my_rank, const unsigned char* handles /* n_peers x 64 */, const int64_t* part_bounds /* n_peers + 1 */);
/* Close every peer mapping this context opened (cudaIpcCloseMemHandle) and forget the handles.  COLLECTIVE by contract: before
 * any rank re-exports its receive buffers (nrp_peer_export / nrp_peer_export_direct may free and re-allocate them), ALL ranks
 * call nrp_peer_close and synchronise (barrier) -- CUDA leaves freeing an exported allocation that an importer still maps
 * undefined. */
int nrp_peer_close(nrp_ctx* ctx);

/* run all work of the context on a caller-owned CUDA stream (cudaStream_t), e.g. torch's current stream; NULL
 * restores a private stream */
int nrp_ctx_set_stream(nrp_ctx* ctx, void* cuda_stream);
int nrp_shard_extract(nrp_ctx* ctx, int64_t part_lo, int64_t part_hi, int k, int internal_repeats, int n_owners,
                      int64_t* owner_counts /* n_owners, host */);
/* device pointers of the owner-partitioned records (owner o's run starts at the prefix sum of owner_counts) */
int nrp_shard_send_buffers(nrp_ctx* ctx, void** keys, void** vals, int* key_bytes);
/* keys_dev/vals_dev: n_records received records in device memory, borrowed until the call returns */
int nrp_shard_group(nrp_ctx* ctx, const void* keys_dev, const uint32_t* vals_dev, int64_t n_records, int k, int internal_repeats);
/* copy the n_parts per-part flag bytes to / from caller-owned device memory (for the all-reduce) */
int nrp_flags_export(nrp_ctx* ctx, uint8_t* dst_dev);
int nrp_flags_import(nrp_ctx* ctx, const uint8_t* src_dev);
/* Returns NRP_RETRY_DENSE_FLAGS (> 0) when a list handed to nrp_flags_repeat_apply was truncated: the caller combines the flag
 * bytes densely (nrp_flags_export / all-reduce / nrp_flags_import) and calls nrp_shard_finish again. */
int nrp_shard_finish(nrp_ctx* ctx, int k, int internal_repeats, int want_edges);
#define NRP_RETRY_DENSE_FLAGS 1
/* Sparse form of the flag exchange.  Only the repeat bit has to travel between ranks during the device phase (it is set by the
 * OWNER of a repeated k-mer, every rank must drop the part's records; palindrome / background bits are set and used by the
 * part's home rank).  nrp_flags_repeat_list writes {count, part indices...} of the parts this rank flagged into list_dev
 * (1 + cap words, device memory); after an all-gather, nrp_flags_repeat_apply ORs the lists of all ranks (world rows of
 * 1 + cap words) into this rank's flags. */
int nrp_flags_repeat_list(nrp_ctx* ctx, uint32_t* list_dev, int cap);
int nrp_flags_repeat_apply(nrp_ctx* ctx, const uint32_t* lists_dev, int world, int cap);
/* shared_keys: canonical k-mers of ALL ranks' groups, any order (host memory, or device memory when keys_on_device != 0);
 * afterwards nrp_result_last_single holds valid entries for the parts of this rank's shard */
int nrp_shard_last_single(nrp_ctx* ctx, const uint64_t* shared_keys, int64_t n_shared, int k, int keys_on_device);
/* replace this context's edge list by the distinct union of n gathered (u < v) edges in device memory;
 * with v_dev == NULL, u_dev points to n 64-bit codes (u << 32) | v */
int nrp_edges_merge(nrp_ctx* ctx, const uint32_t* u_dev, const uint32_t* v_dev, int64_t n);
/* device pointer + element count of a result array, for zero-copy use by the host layer's collectives */
#define NRP_D_GROUP_KEYS   0   /* uint64 */
#define NRP_D_EDGE_U       1   /* uint32 */
#define NRP_D_EDGE_V       2   /* uint32 */
#define NRP_D_FLAGS        3   /* uint8  */
#define NRP_D_LAST_SINGLE  4   /* int32  */
#define NRP_D_INDPTR       5   /* int64  */
#define NRP_D_MEMBERS      6   /* uint32 */
#define NRP_D_FIRST_WINDOW 7   /* uint32 */
int nrp_result_device(nrp_ctx* ctx, int what, void** ptr, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* NRP_B200_H */
