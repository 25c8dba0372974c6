// Vertex-cover heuristics of Finder mode on compact arrays: 'nrp2' and 'nrpG' (reference nrpcalc/base/vercov.py:12-154),
// the adjacency-list body that recovery dumps (recovery.py:26-30) and one round of the independent-set expansion
// (recovery.py:80-86).  Plain C++ (g++), part of libnrptail.so, C ABI in include/nrp_tail.h, bound in native_tail.py.
//
// The reference's routines are order sensitive: ties are broken by node-iteration order, adjacency order and the global
// Mersenne Twister (random.shuffle).  Everything observable is reproduced:
//   * a working graph is the CSR of an nx.Graph (node order, per-node adjacency order) restricted to a node subset; deleting a
//     node or an edge leaves the order of what remains unchanged, exactly like deleting keys from the dicts networkx uses;
//   * random.shuffle / Random._randbelow_with_getrandbits are re-implemented on the interpreter's own MT19937 state
//     (random.getstate() in, random.setstate() out), so the stream of random numbers continues where Python left it;
//   * sorted(..., key=degree) is a stable sort; deque.extendleft reverses.
// Printing is not reproduced: the Python routines run when verbose is set.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>
#include "cpyset.hpp"
#include "../../include/nrp_tail.h"

namespace {

// ---- MT19937 exactly as CPython's _randommodule.c drives it ------------------------------------------------------------
struct Twister {
  uint32_t mt[624];
  int index;
  uint32_t next() {
    if (index >= 624) {
      static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
      int kk;
      uint32_t y;
      for (kk = 0; kk < 624 - 397; ++kk) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u]; }
      for (; kk < 623; ++kk) { y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu); mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u]; }
      y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
      index = 0;
    }
    uint32_t y = mt[index++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
  // Random._randbelow_with_getrandbits(n), n >= 1: k = n.bit_length(); r = getrandbits(k) until r < n
  uint32_t below(uint32_t n) {
    int k = 32 - __builtin_clz(n);
    uint32_t r = next() >> (32 - k);
    while (r >= n) r = next() >> (32 - k);
    return r;
  }
  // random.shuffle(x): for i in reversed(range(1, len(x))): j = randbelow(i + 1); x[i], x[j] = x[j], x[i]
  void shuffle(std::vector<int32_t>& x) {
    for (size_t i = x.size(); i-- > 1;) {
      const uint32_t j = below((uint32_t)i + 1u);
      std::swap(x[i], x[j]);
    }
  }
};

// ---- working graph: CSR slots with tombstones ---------------------------------------------------------------------------
struct Work {
  int32_t n = 0;
  const int64_t* indptr = nullptr;
  std::vector<int32_t> nbr;        // neighbour index per slot
  std::vector<int64_t> rev;        // slot of the reverse edge
  std::vector<uint8_t> slot_alive, node_alive;
  std::vector<int32_t> deg;
  int64_t n_edges = 0;

  bool has(int32_t u) const { return node_alive[u] != 0; }
  void remove_edge_slot(int64_t s, int32_t u) {            // slot s of node u
    const int32_t v = nbr[s];
    slot_alive[s] = 0; slot_alive[rev[s]] = 0;
    --deg[u]; --deg[v]; --n_edges;
  }
  void remove_node(int32_t u) {
    if (!node_alive[u]) return;
    for (int64_t s = indptr[u]; s < indptr[u + 1]; ++s)
      if (slot_alive[s]) remove_edge_slot(s, u);
    node_alive[u] = 0;
  }
  template <typename F> void for_neighbours(int32_t u, F f) const {
    for (int64_t s = indptr[u]; s < indptr[u + 1]; ++s)
      if (slot_alive[s]) f(nbr[s], s);
  }
};

// [nb for nb in graph[node] if len(graph[nb]) <= 2], sorted (stable) by degree (vercov.py:28-34, 57-60)
void low_degree_neighbours(const Work& g, int32_t node, std::vector<int32_t>& out) {
  out.clear();
  g.for_neighbours(node, [&](int32_t nb, int64_t) { if (g.deg[nb] <= 2) out.push_back(nb); });
  std::stable_sort(out.begin(), out.end(), [&](int32_t a, int32_t b) { return g.deg[a] < g.deg[b]; });
}

// get_max_degree_nodes over (node, degree) pairs in iteration order, then shuffled (vercov.py:12-20): the running maximum starts
// at 0, so isolated nodes tie with it
struct TopDegree {
  int32_t best = 0;
  std::vector<int32_t> nodes;
  void reset() { best = 0; nodes.clear(); }
  void see(int32_t node, int32_t degree) {
    if (degree == best) nodes.push_back(node);
    else if (degree > best) { best = degree; nodes.clear(); nodes.push_back(node); }
  }
};

// Build the working graph from a CSR over indices 0..n-1 (every node and slot alive).  Returns false if the adjacency is not symmetric.
bool make_work(Work& g, const int64_t* indptr, const int32_t* adj_index, int64_t n_nodes) {
  g.n = (int32_t)n_nodes;
  g.indptr = indptr;
  const int64_t n_slots = indptr[n_nodes];
  g.nbr.assign(adj_index, adj_index + n_slots);
  g.rev.assign((size_t)n_slots, -1);
  g.slot_alive.assign((size_t)n_slots, 1);
  g.node_alive.assign((size_t)n_nodes, 1);
  g.deg.assign((size_t)n_nodes, 0);
  std::unordered_map<uint64_t, int64_t> where;
  where.reserve((size_t)n_slots);
  for (int64_t u = 0; u < n_nodes; ++u) {
    g.deg[u] = (int32_t)(indptr[u + 1] - indptr[u]);
    for (int64_t s = indptr[u]; s < indptr[u + 1]; ++s) where.emplace(((uint64_t)(uint32_t)u << 32) | (uint32_t)adj_index[s], s);
  }
  for (int64_t u = 0; u < n_nodes; ++u)
    for (int64_t s = indptr[u]; s < indptr[u + 1]; ++s) {
      auto it = where.find(((uint64_t)(uint32_t)adj_index[s] << 32) | (uint32_t)u);
      if (it == where.end()) return false;
      g.rev[s] = it->second;
    }
  g.n_edges = n_slots / 2;
  return true;
}

// The cover routine proper.  `sequence` receives the nodes in the order the reference adds them to its cover SET (with
// repetitions: adding a present element changes nothing) -- the set's later iteration order depends on that order.
void run_cover(Work& g, int mode, Twister& rng, std::vector<int32_t>& sequence) {
  std::vector<int32_t> scratch, exposed;
  sequence.clear();
  // eliminate_pendants (vercov.py:47-68) with verbose == False: isolated nodes stay in the graph
  auto drop = [&](std::deque<int32_t>& queue) {
    while (!queue.empty()) {
      const int32_t node = queue.front();
      queue.pop_front();
      if (!g.has(node) || g.deg[node] != 1) continue;
      int32_t anchor = -1; int64_t slot = -1;
      for (int64_t s = g.indptr[node]; s < g.indptr[node + 1]; ++s)
        if (g.slot_alive[s]) { anchor = g.nbr[s]; slot = s; break; }
      g.remove_edge_slot(slot, node);
      sequence.push_back(anchor);                                  // cover.add(anchor)
      low_degree_neighbours(g, anchor, exposed);
      for (int32_t nb : exposed) queue.push_front(nb);             // deque.extendleft: one by one, so the order is reversed
      g.remove_node(node);
      g.remove_node(anchor);
    }
  };
  auto remove_node_with_pendants = [&](int32_t node) {           // remove_node_with_pendants (vercov.py:36-45)
    if (!g.has(node)) return;
    low_degree_neighbours(g, node, scratch);
    std::deque<int32_t> queue(scratch.begin(), scratch.end());
    g.remove_node(node);
    if (!queue.empty()) drop(queue);
  };
  {  // init_vercov (vercov.py:70-89): nodes of degree <= 1, stably sorted by degree
    std::vector<int32_t> low;
    for (int32_t u = 0; u < g.n; ++u)
      if (g.node_alive[u] && g.deg[u] <= 1) low.push_back(u);
    std::stable_sort(low.begin(), low.end(), [&](int32_t a, int32_t b) { return g.deg[a] < g.deg[b]; });
    if (!low.empty()) {
      std::deque<int32_t> queue(low.begin(), low.end());
      drop(queue);
    }
  }
  TopDegree top, partner;
  std::vector<int32_t> round;
  while (g.n_edges > 0) {
    top.reset();
    for (int32_t u = 0; u < g.n; ++u)
      if (g.node_alive[u]) top.see(u, g.deg[u]);
    rng.shuffle(top.nodes);
    round = top.nodes;                                             // the list is fixed before the graph changes
    for (int32_t node : round) {
      if (!g.has(node)) continue;
      if (mode == 1) {                                             // 'nrpG': the node alone
        sequence.push_back(node);
        remove_node_with_pendants(node);
        continue;
      }
      partner.reset();
      g.for_neighbours(node, [&](int32_t nb, int64_t) { partner.see(nb, g.deg[nb]); });
      rng.shuffle(partner.nodes);
      if (partner.nodes.empty()) { g.remove_node(node); continue; }
      const int32_t mate = partner.nodes.back();                   // list.pop()
      sequence.push_back(node); sequence.push_back(mate);          // cover.update([node, mate])
      remove_node_with_pendants(node);
      remove_node_with_pendants(mate);
    }
  }
}

}  // namespace

extern "C" {

// mode 0 = 'nrp2' (vercov.py:119-138), 1 = 'nrpG' (vercov.py:140-154).  Working graph = the nodes with keep[i] != 0 of the graph
// (n nodes, CSR over node INDICES 0..n-1 in adj_index).  mt_state: 624 words + index of the interpreter's generator, updated.
// cover_out[i] = 1 for the nodes of the cover.  Returns 0; -1 out of memory; -2 adjacency not symmetric.
int nrp_tail_cover(const int64_t* indptr, const int32_t* adj_index, int64_t n_nodes, const uint8_t* keep, int mode, uint32_t* mt_state,
                   uint8_t* cover_out) {
  try {
    Work g;
    if (!make_work(g, indptr, adj_index, n_nodes)) return -2;
    g.n_edges = 0;                                                 // restrict to the kept nodes: the rest is dead from the start
    for (int64_t u = 0; u < n_nodes; ++u) {
      g.deg[u] = 0;
      g.node_alive[u] = keep[u] ? 1 : 0;
      for (int64_t s = indptr[u]; s < indptr[u + 1]; ++s) {
        g.slot_alive[s] = (keep[u] && keep[adj_index[s]]) ? 1 : 0;
        if (g.slot_alive[s]) { ++g.deg[u]; ++g.n_edges; }
      }
    }
    g.n_edges /= 2;
    Twister rng;
    memcpy(rng.mt, mt_state, sizeof(rng.mt));
    rng.index = (int)mt_state[624];
    std::vector<int32_t> sequence;
    run_cover(g, mode, rng, sequence);
    memcpy(mt_state, rng.mt, sizeof(rng.mt));
    mt_state[624] = (uint32_t)rng.index;
    memset(cover_out, 0, (size_t)n_nodes);
    for (int32_t x : sequence) cover_out[x] = 1;
    return 0;
  } catch (...) {
    return -1;
  }
}

// The whole independent-set recovery (reference recovery.py:32-96) around 'nrp2' / 'nrpG', verbose == False.
//   graph 0   the homology graph as assembled (round 0 runs on it): ids0[n], CSR indptr0 / adj0 over indices into ids0
//   graph 1   what read_adjlist(write_adjlist(graph 0)) returns (every later round runs on a subgraph of it): ids1[n], CSR over
//             indices into ids1 (nrp_tail_reload gives it)
// What makes later rounds order sensitive, and is reproduced: the working graph of round r > 0 is nx.Graph(G1.subgraph(C)), C the
// candidate SET; networkx iterates the filter set instead of G1's nodes when 2 * |C| < |G1| (coreviews.FilterAdjacency.__iter__),
// and the filter is set(nbunch_iter(C)) -- so the node order is CPython's iteration order of a set built from C in C's own
// iteration order, and C is the cover set as the routine built it, minus what difference_update removed (with its
// quarter-dummies resize).  The adjacency of the copy is re-derived by from_dict_of_dicts (an edge is appended to both ends
// when its first end point is visited).  All sets are modelled slot for slot (cpyset.hpp).
// independent_out[i] = 1 for node ids1[i]; returns the number of rounds, or a negative error code.
int64_t nrp_tail_recover(const int64_t* ids0, const int64_t* indptr0, const int32_t* adj0, const int64_t* ids1, const int64_t* indptr1,
                         const int32_t* adj1, int64_t n_nodes, int mode, uint32_t* mt_state, uint8_t* independent_out) {
  try {
    using cpyset::IntSet;
    const int32_t n = (int32_t)n_nodes;
    memset(independent_out, 0, (size_t)n_nodes);
    if (n == 0) return 0;
    int64_t max_id = 0;
    for (int32_t i = 0; i < n; ++i) max_id = std::max(max_id, ids1[i]);
    std::unordered_map<int64_t, int32_t> index1;                 // id -> index in graph 1
    index1.reserve((size_t)n);
    for (int32_t i = 0; i < n; ++i) index1.emplace(ids1[i], i);
    Twister rng;
    memcpy(rng.mt, mt_state, sizeof(rng.mt));
    rng.index = (int)mt_state[624];

    IntSet candidates;                                           // set(homology_graph.nodes())
    for (int32_t i = 0; i < n; ++i) candidates.add(ids0[i]);
    std::vector<uint8_t> member((size_t)n, 1);                   // candidates as a mask over graph-1 indices
    std::vector<int32_t> order, sequence, w_index((size_t)n), w_deg;
    std::vector<int64_t> w_indptr;
    std::vector<int32_t> w_adj, w_fill;
    int64_t rounds = 0;
    for (;; ++rounds) {
      // ---- the working graph of this round and its cover (ids of the cover in the order they were added to the set)
      std::vector<int64_t> cover_ids;
      if (rounds == 0) {
        Work g;
        if (!make_work(g, indptr0, adj0, n_nodes)) return -2;
        run_cover(g, mode, rng, sequence);
        for (int32_t x : sequence) cover_ids.push_back(ids0[x]);
      } else {
        IntSet filter;                                           // show_nodes.nodes = set(G.nbunch_iter(candidates))
        candidates.for_each([&](int64_t id) { filter.add(id); });
        order.clear();
        if (2 * filter.used < (int64_t)n) filter.for_each([&](int64_t id) { order.push_back(index1.find(id)->second); });
        else for (int32_t i = 0; i < n; ++i) if (member[i]) order.push_back(i);
        const int32_t nw = (int32_t)order.size();
        for (int32_t p = 0; p < nw; ++p) w_index[order[p]] = p;
        // from_dict_of_dicts: for u in order, for v in G1[u] (members only): the edge is new iff v has not been visited as u
        w_deg.assign((size_t)nw, 0);
        for (int32_t p = 0; p < nw; ++p) {
          const int32_t u = order[p];
          for (int64_t s = indptr1[u]; s < indptr1[u + 1]; ++s) if (member[adj1[s]]) ++w_deg[p];
        }
        w_indptr.assign((size_t)nw + 1, 0);
        for (int32_t p = 0; p < nw; ++p) w_indptr[p + 1] = w_indptr[p] + w_deg[p];
        w_adj.assign((size_t)w_indptr[nw], 0);
        w_fill.assign((size_t)nw, 0);
        for (int32_t p = 0; p < nw; ++p) {
          const int32_t u = order[p];
          for (int64_t s = indptr1[u]; s < indptr1[u + 1]; ++s) {
            const int32_t v = adj1[s];
            if (!member[v]) continue;
            const int32_t q = w_index[v];
            if (q < p) continue;                                 // added when v was visited
            w_adj[(size_t)(w_indptr[p] + w_fill[p]++)] = q;
            w_adj[(size_t)(w_indptr[q] + w_fill[q]++)] = p;
          }
        }
        Work g;
        if (!make_work(g, w_indptr.data(), w_adj.data(), nw)) return -2;
        run_cover(g, mode, rng, sequence);
        for (int32_t x : sequence) cover_ids.push_back(ids1[order[x]]);
      }
      // ---- recovery.py:80-95
      IntSet cover;
      for (int64_t id : cover_ids) cover.add(id);
      IntSet fresh;
      IntSet::difference(candidates, cover, fresh);              // new_indi = possible - cover
      fresh.for_each([&](int64_t id) { independent_out[index1.find(id)->second] = 1; });
      candidates = cover;                                        // possible = cover (the very set the routine built)
      const int64_t before = candidates.used;
      fresh.for_each([&](int64_t id) {                           // possible.difference_update(graph[node]) per new node
        const int32_t u = index1.find(id)->second;
        for (int64_t s = indptr1[u]; s < indptr1[u + 1]; ++s) candidates.discard(ids1[adj1[s]]);
        candidates.shed_dummies();
      });
      const int64_t after = candidates.used;
      std::fill(member.begin(), member.end(), 0);
      candidates.for_each([&](int64_t id) { member[index1.find(id)->second] = 1; });
      if (after == 0 || before == after) { ++rounds; break; }
    }
    memcpy(mt_state, rng.mt, sizeof(rng.mt));
    mt_state[624] = (uint32_t)rng.index;
    return rounds;
  } catch (...) {
    return -1;
  }
}

// Body of networkx.write_adjlist (generate_adjlist): one line per node in node order, "u v1 v2 ..." listing the neighbours not
// seen as an earlier line's node, in adjacency order.  Returns the number of bytes written, or -(bytes needed) if cap is too small.
int64_t nrp_tail_adjlist(const int64_t* nodes, const int64_t* indptr, const int32_t* adj_index, int64_t n_nodes, char* out, int64_t cap) {
  try {
    std::string text;
    text.reserve((size_t)(indptr[n_nodes] * 4 + n_nodes * 8 + 16));
    std::vector<uint8_t> seen((size_t)n_nodes, 0);
    char buf[24];
    auto put = [&](int64_t v) {
      int len = 0;
      if (v == 0) buf[len++] = '0';
      bool neg = v < 0;
      uint64_t x = neg ? (uint64_t)(-v) : (uint64_t)v;
      while (x) { buf[len++] = (char)('0' + x % 10); x /= 10; }
      if (neg) buf[len++] = '-';
      while (len) text.push_back(buf[--len]);
    };
    for (int64_t u = 0; u < n_nodes; ++u) {
      put(nodes[u]);
      for (int64_t s = indptr[u]; s < indptr[u + 1]; ++s) {
        const int32_t v = adj_index[s];
        if (!seen[v]) { text.push_back(' '); put(nodes[v]); }
      }
      text.push_back('\n');
      seen[u] = 1;
    }
    if ((int64_t)text.size() > cap) return -(int64_t)text.size();
    memcpy(out, text.data(), text.size());
    return (int64_t)text.size();
  } catch (...) {
    return INT64_MIN;
  }
}

}  // extern "C"
