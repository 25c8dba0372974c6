/* C ABI of libnrptail.so: the order-exact HOST tail of the Finder graph build on compact arrays (plain C++, no CUDA).
 *
 * Replaces, with identical results, the clique tail of reference nrpcalc/base/hgraph.py:96-200 (get_maximal_cliques,
 * get_repeat_cliques, build_homology_graph) and the dump / reload of nrpcalc/base/recovery.py:26-30.  Their results depend on
 * the iteration order of CPython sets; the library runs them on a layout-exact model of those sets (csrc/cpyset.hpp) instead
 * of on Python objects.  Bound with ctypes in nrpcalc_b200/native_tail.py, which self-checks the model against the running
 * interpreter before first use.
 */
#ifndef NRP_TAIL_H
#define NRP_TAIL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* The compressed stream of id tuples in the order the reference meets them (descending first-insertion rank, as
 * nrpcalc_b200.hgraph.ordered_stream_arrays builds it from the device result): tuple t = members[start[t] .. start[t] + size[t]).
 * Returns a handle to the result: the graph's node ids in insertion order, each node's neighbours in insertion order (CSR),
 * and the clique stream that was inserted. */
void* nrp_tail_run(const int64_t* start, const int64_t* size, int64_t n_tuples, const int64_t* members);

/* The graph that networkx's read_adjlist(write_adjlist(G)) returns, from the arrays of G (node order, CSR adjacency). */
void* nrp_tail_reload(const int64_t* nodes, const int64_t* indptr, const int64_t* adj, int64_t n_nodes);

/* out = {nodes, adjacency entries, cliques, clique members} of a result */
void nrp_tail_sizes(void* handle, int64_t out[4]);
/* copy a result into caller-allocated arrays (nodes[n], indptr[n + 1], adj[...]; clique_off / clique_nodes may be NULL) */
void nrp_tail_fetch(void* handle, int64_t* nodes, int64_t* indptr, int64_t* adj, int64_t* clique_off, int64_t* clique_nodes);
void nrp_tail_free(void* handle);

/* Vertex cover 'nrp2' (mode 0, reference vercov.py:119-138) or 'nrpG' (mode 1, vercov.py:140-154) with the pendant elimination
 * of vercov.py:47-89, on the CSR of a graph restricted to the nodes with keep[i] != 0 (n_nodes nodes in iteration order,
 * adj_index = neighbour INDICES in adjacency order).  mt_state = the 624 words + index of the interpreter's Mersenne Twister
 * (random.getstate()[1]); random.shuffle is replayed on it and the state is updated in place.  cover_out[i] = 1 for cover nodes.
 * Returns 0; -1 out of memory; -2 the adjacency is not symmetric. */
int nrp_tail_cover(const int64_t* indptr, const int32_t* adj_index, int64_t n_nodes, const uint8_t* keep, int mode, uint32_t* mt_state,
                   uint8_t* cover_out);
/* The whole independent-set recovery of reference recovery.py:32-96 around 'nrp2' / 'nrpG' (verbose off): graph 0 = the homology
 * graph as assembled (ids0, CSR over indices into ids0), graph 1 = its dump-and-reload (nrp_tail_reload; ids1, CSR over indices
 * into ids1).  independent_out[i] = 1 when ids1[i] is in the recovered independent set.  Returns the number of rounds (>= 1 for a
 * non-empty graph); -1 out of memory; -2 malformed adjacency.  mt_state as in nrp_tail_cover. */
int64_t nrp_tail_recover(const int64_t* ids0, const int64_t* indptr0, const int32_t* adj0, const int64_t* ids1, const int64_t* indptr1,
                         const int32_t* adj1, int64_t n_nodes, int mode, uint32_t* mt_state, uint8_t* independent_out);
/* The lines networkx.write_adjlist emits for a graph (recovery.py:26-30), without the comment header.  Returns the bytes written,
 * or -(bytes needed) when cap is too small. */
int64_t nrp_tail_adjlist(const int64_t* nodes, const int64_t* indptr, const int32_t* adj_index, int64_t n_nodes, char* out, int64_t cap);

/* Test hooks (tests/test_cpyset.py): hash(tuple of small ints) as CPython computes it, and a small interpreter over a pool of
 * modelled sets (rows of opcode, a, b, c -- see csrc/host_replay.cpp) whose iteration orders are compared with real sets. */
int64_t nrp_tail_tuple_hash(const int64_t* items, int64_t len);
int64_t nrp_tail_set_script(const int64_t* ops, int64_t n_ops, int64_t n_sets, int64_t* out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
