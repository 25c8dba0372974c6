/*
 * ORACLE (test infrastructure, NOT product code): plain-C CPU restatement of the reference's Finder-mode
 * repeat-table construction, used (a) as a second, independent checker of the CUDA path at sizes where the
 * Python oracles are too slow and (b) as the CPU baseline that bench.py times next to the GPU.
 *
 * It follows the reference's algorithm, not the GPU's: parts are visited in processing order, each part is
 * tested for internal repeats and background hits first, and the windows of surviving parts are appended to
 * a hash table keyed by the canonical k-mer whose values are id lists in insertion order:
 *
 *   windows / reverse complement / canonical min   <- nrpcalc/base/utils.py:37-50
 *   internal-repeat exclusion                      <- nrpcalc/base/hgraph.py:50-69
 *   background exclusion (any canonical K-window)  <- nrpcalc/base/hgraph.py:72-79, kmerSetDB.py:177-207
 *   repeat_dict[min(kmer, rmer)].append(id)        <- nrpcalc/base/hgraph.py:82-89
 *
 * Encoding: A=0 C=1 G=2 T=3, most significant base first (integer order == the reference's string order).
 * Only parts with length >= k are windowed here (the Python side submits shorter parts per length class).
 *
 * Threads: n_threads > 1 shards the TABLE by hash(key) % n_threads; every thread walks all parts and keeps the
 * keys it owns, which keeps per-key insertion order identical to the single-threaded run.
 *
 * Parity status: PINNED by tests/test_oracle.py (equality with oracle/np_oracle.py, which is pinned to the
 * reference's golden fixtures).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct {
  uint64_t key;
  uint32_t count;      /* 0 = empty slot */
  uint32_t first_p;    /* rank of the first insertion: (part, window) */
  uint32_t first_j;
  uint32_t head, tail; /* member list in the pool */
} slot_t;

typedef struct {
  slot_t* slots;
  uint64_t mask;
  uint32_t* pool_part;
  uint32_t* pool_next;
  uint64_t pool_used, pool_cap;
} table_t;

typedef struct {
  int64_t n_parts, n_records, n_groups, n_members;
  int n_tables;
  table_t* tables;
} oracle_t;

static inline uint32_t code_of(uint8_t c) { uint32_t t = (c >> 1) & 3u; return t ^ (t >> 1); }

static inline uint64_t revcomp(uint64_t fwd, int k) {
  uint64_t x = ~fwd, r = 0;
  for (int i = 0; i < k; ++i) { r = (r << 2) | (x & 3u); x >>= 2; }
  return r;
}

static inline uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

/* canonical k-mers of one part into out[] (and palindrome detection); returns the number of windows.
 * Forward and reverse-complement values roll by one base per window. */
static int64_t part_windows(const uint8_t* seq, int64_t len, int k, uint64_t* out, int* has_palindrome) {
  if (len < k) return 0;
  const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
  const int top = 2 * (k - 1);
  uint64_t fwd = 0, rc = 0;
  for (int t = 0; t < k - 1; ++t) {
    uint64_t c = code_of(seq[t]);
    fwd = (fwd << 2) | c;
    rc = (rc >> 2) | ((3u - c) << top);
  }
  int64_t w = 0;
  for (int64_t j = 0; j + k <= len; ++j) {
    uint64_t c = code_of(seq[j + k - 1]);
    fwd = ((fwd << 2) | c) & mask;
    rc = (rc >> 2) | ((3u - c) << top);
    if (fwd == rc && has_palindrome) *has_palindrome = 1;
    out[w++] = fwd < rc ? fwd : rc;
  }
  return w;
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

static int bg_has(const uint64_t* bg, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (bg[mid] == key) return 1;
    if (bg[mid] < key) lo = mid + 1; else hi = mid;
  }
  return 0;
}

typedef struct {
  const uint8_t* ascii; const int64_t* off; int64_t n_parts; int k; int internal_repeats;
  const uint64_t* bg; int64_t n_bg; int bg_K; int n_threads; int tid;
  int n_slices, slice;     /* only keys with (mix(key) >> 52) % n_slices == slice enter the table (bounds the memory of one pass) */
  int64_t max_len, m_max;
  uint8_t* flags; int32_t* last_single;
  table_t* table; int64_t n_records;
} job_t;

/* pass 1: exclusion flags; thread t takes blocks of 256 parts round-robin */
static void* flags_worker(void* arg) {
  job_t* jb = (job_t*)arg;
  uint64_t* win = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(jb->max_len + 1));
  uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(jb->max_len + 1));
  for (int64_t blk = jb->tid; blk * 256 < jb->n_parts; blk += jb->n_threads) {
    int64_t hi = (blk + 1) * 256 < jb->n_parts ? (blk + 1) * 256 : jb->n_parts;
    for (int64_t p = blk * 256; p < hi; ++p) {
      const uint8_t* seq = jb->ascii + jb->off[p];
      const int64_t len = jb->off[p + 1] - jb->off[p];
      uint8_t f = 0;
      jb->last_single[p] = -1;
      if (!jb->internal_repeats) {
        int pal = 0;
        int64_t w = part_windows(seq, len, jb->k, win, &pal);
        if (pal) f |= 2;
        if (w <= 64) {                                   /* short parts: insertion sort beats qsort's callbacks */
          for (int64_t i = 0; i < w; ++i) {
            uint64_t v = win[i]; int64_t q = i;
            while (q > 0 && tmp[q - 1] > v) { tmp[q] = tmp[q - 1]; --q; }
            tmp[q] = v;
          }
        } else {
          memcpy(tmp, win, sizeof(uint64_t) * (size_t)w);
          qsort(tmp, (size_t)w, sizeof(uint64_t), cmp_u64);
        }
        for (int64_t i = 1; i < w; ++i) if (tmp[i] == tmp[i - 1]) { f |= 1; break; }
      }
      if (jb->n_bg > 0 && jb->bg_K <= len) {
        int64_t w = part_windows(seq, len, jb->bg_K, win, NULL);
        for (int64_t i = 0; i < w; ++i) if (bg_has(jb->bg, jb->n_bg, win[i])) { f |= 4; break; }
      }
      jb->flags[p] = f;
    }
  }
  free(win); free(tmp);
  return NULL;
}

/* pass 2: the table shard owned by thread t */
static void* table_worker(void* arg) {
  job_t* jb = (job_t*)arg;
  table_t* tb = jb->table;
  const int n_threads = jb->n_threads, t = jb->tid;
  const int64_t m_mine = jb->m_max / (jb->n_slices > 1 ? jb->n_slices : 1);
  uint64_t want = (uint64_t)(2 * m_mine / n_threads + m_mine / 8) + 64, cap = 64;
  while (cap < want) cap <<= 1;
  tb->slots = (slot_t*)calloc(cap, sizeof(slot_t));
  tb->mask = cap - 1;
  tb->pool_cap = (uint64_t)(m_mine / n_threads + m_mine / 8) + 64;
  tb->pool_part = (uint32_t*)malloc(sizeof(uint32_t) * tb->pool_cap);
  tb->pool_next = (uint32_t*)malloc(sizeof(uint32_t) * tb->pool_cap);
  uint64_t* win = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(jb->max_len + 1));
  int64_t n_records = 0;
  for (int64_t p = 0; p < jb->n_parts; ++p) {
    if (jb->flags[p]) continue;
    const int64_t w = part_windows(jb->ascii + jb->off[p], jb->off[p + 1] - jb->off[p], jb->k, win, NULL);
    for (int64_t j = 0; j < w; ++j) {
      const uint64_t key = win[j], h = mix(key);
      if (jb->n_slices > 1 && (int)((h >> 52) % (uint64_t)jb->n_slices) != jb->slice) continue;    /* top bits: the slot index uses the low ones */
      if (n_threads > 1 && (int)(((h >> 40) & 0xfffu) % (uint64_t)n_threads) != t) continue;
      ++n_records;
      if (tb->pool_used == tb->pool_cap) {
        tb->pool_cap *= 2;
        tb->pool_part = (uint32_t*)realloc(tb->pool_part, sizeof(uint32_t) * tb->pool_cap);
        tb->pool_next = (uint32_t*)realloc(tb->pool_next, sizeof(uint32_t) * tb->pool_cap);
      }
      const uint32_t node = (uint32_t)tb->pool_used++;
      tb->pool_part[node] = (uint32_t)p;
      tb->pool_next[node] = 0xffffffffu;
      uint64_t s = h & tb->mask;
      for (;;) {
        slot_t* sl = &tb->slots[s];
        if (sl->count == 0) {
          sl->key = key; sl->count = 1; sl->first_p = (uint32_t)p; sl->first_j = (uint32_t)j; sl->head = sl->tail = node;
          break;
        }
        if (sl->key == key) {
          tb->pool_next[sl->tail] = node; sl->tail = node; sl->count++;
          break;
        }
        s = (s + 1) & tb->mask;
      }
    }
  }
  free(win);
  jb->n_records = n_records;
  return NULL;
}

static void run_threads(void* (*fn)(void*), job_t* jobs, int n_threads) {
  if (n_threads == 1) { fn(&jobs[0]); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(th);
}

/* One pass over slice `slice` of `n_slices` slices of the key space.  flags_given != 0: `flags` and `last_single` come from
 * an earlier pass over another slice (flags are complete after the first pass; last_single is max-accumulated). */
void* nrp_oracle_build_slice(const uint8_t* ascii, const int64_t* off, int64_t n_parts, int k, int internal_repeats,
                             const uint64_t* bg_sorted, int64_t n_bg, int bg_K, int n_threads,
                             uint8_t* flags, int32_t* last_single, int n_slices, int slice, int flags_given) {
  if (n_threads < 1) n_threads = 1;
  if (n_slices < 1) n_slices = 1;
  int64_t max_len = 0, m_max = 0;
  for (int64_t p = 0; p < n_parts; ++p) {
    int64_t len = off[p + 1] - off[p];
    if (len > max_len) max_len = len;
    if (len >= k) m_max += len - k + 1;
  }
  oracle_t* o = (oracle_t*)calloc(1, sizeof(oracle_t));
  o->n_parts = n_parts;
  o->n_tables = n_threads;
  o->tables = (table_t*)calloc((size_t)n_threads, sizeof(table_t));
  job_t* jobs = (job_t*)calloc((size_t)n_threads, sizeof(job_t));
  for (int t = 0; t < n_threads; ++t) {
    job_t jb = {ascii, off, n_parts, k, internal_repeats, bg_sorted, n_bg, bg_K, n_threads, t, n_slices, slice, max_len, m_max,
                flags, last_single, &o->tables[t], 0};
    jobs[t] = jb;
  }
  if (!flags_given) run_threads(flags_worker, jobs, n_threads);
  run_threads(table_worker, jobs, n_threads);
  for (int t = 0; t < n_threads; ++t) o->n_records += jobs[t].n_records;
  free(jobs);
  /* singletons -> last_single (max window per part); groups counted */
  for (int t = 0; t < o->n_tables; ++t) {
    table_t* tb = &o->tables[t];
    for (uint64_t s = 0; s <= tb->mask; ++s) {
      slot_t* sl = &tb->slots[s];
      if (sl->count == 1) {
        if ((int32_t)sl->first_j > last_single[sl->first_p]) last_single[sl->first_p] = (int32_t)sl->first_j;
      } else if (sl->count >= 2) {
        o->n_groups++;
        o->n_members += sl->count;
      }
    }
  }
  return o;
}

void* nrp_oracle_build(const uint8_t* ascii, const int64_t* off, int64_t n_parts, int k, int internal_repeats,
                       const uint64_t* bg_sorted, int64_t n_bg, int bg_K, int n_threads,
                       uint8_t* flags, int32_t* last_single) {
  return nrp_oracle_build_slice(ascii, off, n_parts, k, internal_repeats, bg_sorted, n_bg, bg_K, n_threads, flags, last_single, 1, 0, 0);
}

void nrp_oracle_counts(void* handle, int64_t out[4]) {
  oracle_t* o = (oracle_t*)handle;
  out[0] = o->n_records; out[1] = o->n_groups; out[2] = o->n_members; out[3] = o->n_parts;
}

/* groups with >= 2 members (table order): CSR members as part indices, rank = (rank_p, rank_j), and the key */
void nrp_oracle_groups(void* handle, int64_t* indptr, uint32_t* members, uint32_t* rank_p, uint32_t* rank_j, uint64_t* keys) {
  oracle_t* o = (oracle_t*)handle;
  int64_t g = 0, m = 0;
  for (int t = 0; t < o->n_tables; ++t) {
    table_t* tb = &o->tables[t];
    for (uint64_t s = 0; s <= tb->mask; ++s) {
      slot_t* sl = &tb->slots[s];
      if (sl->count < 2) continue;
      indptr[g] = m;
      rank_p[g] = sl->first_p; rank_j[g] = sl->first_j;
      if (keys) keys[g] = sl->key;
      for (uint32_t node = sl->head; node != 0xffffffffu; node = tb->pool_next[node]) members[m++] = tb->pool_part[node];
      ++g;
    }
  }
  indptr[g] = m;
}

void nrp_oracle_free(void* handle) {
  oracle_t* o = (oracle_t*)handle;
  if (!o) return;
  for (int t = 0; t < o->n_tables; ++t) {
    free(o->tables[t].slots); free(o->tables[t].pool_part); free(o->tables[t].pool_next);
  }
  free(o->tables);
  free(o);
}
