// Order-exact host tail of the Finder graph build on compact arrays (no Python objects): the clique replay and the graph
// assembly of reference nrpcalc/base/hgraph.py:96-200, run on layout-exact models of CPython's sets (cpyset.hpp).
// Plain C++ (g++), C ABI for ctypes: nrpcalc_b200/native_tail.py.  Built by `make -C nrpcalc_b200/csrc tail`.
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <unordered_map>
#include <vector>
#include "cpyset.hpp"
#include "../../include/nrp_tail.h"

using cpyset::IntSet;
using cpyset::TupleSet;

extern "C" {

// ---- test hooks (tests/test_cpyset.py drives these against the interpreter's own sets) --------------------------------
int64_t nrp_tail_tuple_hash(const int64_t* items, int64_t len) { return cpyset::tuple_hash(items, len); }

// A tiny interpreter over a pool of sets.  ops = rows of (opcode, a, b, c):
//   0 a = set()            1 a.add(b)            2 a.discard(b)         3 a |= b / a.update(b)      4 a &= b
//   5 a = b - c            6 a = set(); a |= b   7 output: len(a), then the elements of a in iteration order
//   8 a = b.difference(dict with the keys of set c)
// Returns the number of int64 written to out (or -1 if out is too small).
int64_t nrp_tail_set_script(const int64_t* ops, int64_t n_ops, int64_t n_sets, int64_t* out, int64_t cap) {
  std::vector<IntSet> pool((size_t)n_sets);
  int64_t w = 0;
  for (int64_t i = 0; i < n_ops; ++i) {
    const int64_t op = ops[4 * i], a = ops[4 * i + 1], b = ops[4 * i + 2], c = ops[4 * i + 3];
    switch (op) {
      case 0: pool[a].reset(); break;
      case 1: pool[a].add(b); break;
      case 2: pool[a].discard(b); break;
      case 3: pool[a].merge(pool[b]); break;
      case 4: { IntSet tmp; IntSet::intersection(pool[a], pool[b], tmp); pool[a] = tmp; break; }
      case 5: { IntSet tmp; IntSet::difference(pool[b], pool[c], tmp); pool[a] = tmp; break; }
      case 6: pool[a].reset(); pool[a].merge(pool[b]); break;
      case 7:
        if (w + 1 + pool[a].used > cap) return -1;
        out[w++] = pool[a].used;
        pool[a].for_each([&](int64_t k) { out[w++] = k; });
        break;
      case 8: {
        std::vector<int64_t> keys;
        pool[c].for_each([&](int64_t k) { keys.push_back(k); });
        IntSet tmp;
        const IntSet& other = pool[c];
        IntSet::difference_dict(pool[b], (int64_t)keys.size(), keys.data(), [&](int64_t k) { return other.contains(k); }, tmp);
        pool[a] = tmp;
        break;
      }
      default: return -2;
    }
  }
  return w;
}

// ---- the tail itself ------------------------------------------------------------------------------------------------
// id -> small index: a flat table when the ids are dense relative to the data (they are first indices in the caller's list),
// a hash map otherwise
struct IdIndex {
  std::vector<int32_t> flat;
  std::unordered_map<int64_t, int32_t> sparse;
  bool dense = false;
  void init(int64_t max_id, int64_t n_values) {
    dense = max_id < 4 * n_values + 1024;
    if (dense) flat.assign((size_t)max_id + 1, -1);
    else sparse.reserve((size_t)n_values);
  }
  int32_t get(int64_t id) const {
    if (dense) return flat[(size_t)id];
    auto it = sparse.find(id);
    return it == sparse.end() ? -1 : it->second;
  }
  void set(int64_t id, int32_t idx) { if (dense) flat[(size_t)id] = idx; else sparse.emplace(id, idx); }
};

// open-addressing set of 64-bit pair codes (never 0: a node is not its own neighbour and index pairs are offset by one)
struct PairSet {
  std::vector<uint64_t> slots;
  size_t mask, used = 0;
  PairSet() : slots(1 << 16, 0), mask((1 << 16) - 1) {}
  static size_t mix(uint64_t k) { k *= 0x9E3779B97F4A7C15ull; return (size_t)(k ^ (k >> 29)); }
  bool contains(uint64_t k) const {
    for (size_t i = mix(k) & mask;; i = (i + 1) & mask) {
      if (slots[i] == k) return true;
      if (slots[i] == 0) return false;
    }
  }
  void insert(uint64_t k) {
    if ((used + 1) * 2 > mask) {
      std::vector<uint64_t> old;
      old.swap(slots);
      slots.assign((mask + 1) * 2, 0);
      mask = slots.size() - 1;
      for (uint64_t v : old)
        if (v) { size_t i = mix(v) & mask; while (slots[i]) i = (i + 1) & mask; slots[i] = v; }
    }
    size_t i = mix(k) & mask;
    while (slots[i]) { if (slots[i] == k) return; i = (i + 1) & mask; }
    slots[i] = k; ++used;
  }
};
inline uint64_t pair_code(int32_t a, int32_t b) { return ((uint64_t)((uint32_t)a + 1u) << 32) | ((uint32_t)b + 1u); }

struct TailResult {
  std::vector<int64_t> nodes, indptr, adj;          // graph: node ids in insertion order, CSR of neighbours in insertion order
  std::vector<int64_t> clique_off, clique_nodes;    // the clique stream that was inserted (tests compare it with the Python replay)
};

// Input: the compressed stream of id tuples in the order the reference meets them (descending first-insertion rank):
// tuple t = members[start[t] .. start[t] + size[t]).
static void* tail_run(const int64_t* start, const int64_t* size, int64_t n_tuples, const int64_t* members) {
  std::unique_ptr<TailResult> owner(new TailResult());
  TailResult* R = owner.get();
  const bool timing = std::getenv("NRP_TAIL_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "  tail %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  // 1. distinct = set(); distinct.add(tuple) in stream order (hgraph.py:131-135).  Items are stream indices.
  std::vector<int64_t> order1;
  {
    TupleSet distinct;
    auto equal = [&](int64_t a, int64_t b) {
      return size[a] == size[b] && std::memcmp(members + start[a], members + start[b], sizeof(int64_t) * (size_t)size[a]) == 0;
    };
    for (int64_t t = 0; t < n_tuples; ++t) distinct.add(t, cpyset::tuple_hash(members + start[t], size[t]), equal);
    order1.reserve((size_t)distinct.used);
    for (size_t s = 0; s <= distinct.mask; ++s)                  // distinct.pop() until empty: slot order
      if (distinct.slot_item[s] >= 0) order1.push_back(distinct.slot_item[s]);
  }
  lap("distinct tuples");
  // 2. as_sets.append(set(distinct.pop())): only the iteration order of each set is ever observed
  const int64_t n_sets = (int64_t)order1.size();
  std::vector<int64_t> cl_off((size_t)n_sets + 1, 0), cl_nodes;
  {
    IntSet tmp;
    for (int64_t p = 0; p < n_sets; ++p) {
      const int64_t t = order1[(size_t)p];
      if (size[t] == 1) cl_nodes.push_back(members[start[t]]);
      else {
        tmp.reset();
        for (int64_t i = 0; i < size[t]; ++i) tmp.add(members[start[t] + i]);
        tmp.for_each([&](int64_t k) { cl_nodes.push_back(k); });
      }
      cl_off[(size_t)p + 1] = (int64_t)cl_nodes.size();
    }
  }
  lap("clique sets");
  // 3. maximal cliques (hgraph.py:96-115): member_of[node] = positions of the cliques that hold the node
  std::vector<int64_t> kept;
  {
    int64_t max_id = 0;
    for (const int64_t id : cl_nodes) max_id = id > max_id ? id : max_id;
    IdIndex slot_of;
    slot_of.init(max_id, (int64_t)cl_nodes.size());
    std::vector<IntSet> member_of;
    std::vector<int32_t> node_slot(cl_nodes.size());
    for (int64_t p = 0; p < n_sets; ++p)
      for (int64_t i = cl_off[(size_t)p]; i < cl_off[(size_t)p + 1]; ++i) {
        int32_t slot = slot_of.get(cl_nodes[(size_t)i]);
        if (slot < 0) { slot = (int32_t)member_of.size(); slot_of.set(cl_nodes[(size_t)i], slot); member_of.emplace_back(); }
        node_slot[(size_t)i] = slot;
        member_of[(size_t)slot].add(p);
      }
    lap("member_of");
    IntSet keep, common, tmp, self, rest;
    for (int64_t p = 0; p < n_sets; ++p) {
      // `common` ends as the intersection of the members' position sets and always holds p.  If some member occurs in no other
      // clique the intersection is {p} whatever the order of the operations, and only its size is used then: keep.add(p).
      bool alone = false;
      for (int64_t i = cl_off[(size_t)p]; i < cl_off[(size_t)p + 1] && !alone; ++i) alone = member_of[(size_t)node_slot[(size_t)i]].used == 1;
      if (alone) { keep.add(p); continue; }
      common.reset();
      for (int64_t i = cl_off[(size_t)p]; i < cl_off[(size_t)p + 1]; ++i) {
        const IntSet& where = member_of[(size_t)node_slot[(size_t)i]];
        if (common.used == 0) common.merge(where);                                  // common |= member_of[node]
        else { IntSet::intersection(common, where, tmp); std::swap(common, tmp); }   // common &= member_of[node]
      }
      if (common.used == 1) keep.add(p);
      else {                                                                        // keep.update(common - set([pos]))
        self.reset(); self.add(p);
        IntSet::difference(common, self, rest);
        keep.merge(rest);
      }
    }
    keep.for_each([&](int64_t p) { kept.push_back(p); });                           // [clique_sets[pos] for pos in keep]
  }
  lap("maximal filter");
  // 4. final.add(tuple(as_sets.pop())) from the END of the kept list, then final.pop() until empty (hgraph.py:150-155)
  std::vector<int64_t> order3;
  {
    TupleSet final_set;
    auto equal = [&](int64_t a, int64_t b) {
      const int64_t la = cl_off[(size_t)a + 1] - cl_off[(size_t)a], lb = cl_off[(size_t)b + 1] - cl_off[(size_t)b];
      return la == lb && std::memcmp(cl_nodes.data() + cl_off[(size_t)a], cl_nodes.data() + cl_off[(size_t)b], sizeof(int64_t) * (size_t)la) == 0;
    };
    for (size_t i = kept.size(); i-- > 0;) {
      const int64_t p = kept[i];
      final_set.add(p, cpyset::tuple_hash(cl_nodes.data() + cl_off[(size_t)p], cl_off[(size_t)p + 1] - cl_off[(size_t)p]), equal);
    }
    for (size_t s = 0; s <= final_set.mask; ++s)
      if (final_set.slot_item[s] >= 0) order3.push_back(final_set.slot_item[s]);
  }
  lap("final tuples");
  // 5. graph assembly (hgraph.py:157-200): node and adjacency insertion order
  {
    int64_t max_id = 0;
    for (const int64_t id : cl_nodes) max_id = id > max_id ? id : max_id;
    IdIndex index_of;
    index_of.init(max_id, (int64_t)cl_nodes.size());
    std::vector<std::vector<int64_t>> nbrs;                      // by node index: neighbour ids in insertion order
    PairSet linked;                                              // (index_u, index_v) for both directions
    auto ensure = [&](int64_t id) -> int32_t {
      int32_t idx = index_of.get(id);
      if (idx >= 0) return idx;
      idx = (int32_t)R->nodes.size();
      index_of.set(id, idx);
      R->nodes.push_back(id);
      nbrs.emplace_back();
      return idx;
    };
    IntSet pending, fresh;
    std::vector<int64_t> targets;
    R->clique_off.push_back(0);
    for (const int64_t p : order3) {
      const int64_t lo = cl_off[(size_t)p], hi = cl_off[(size_t)p + 1];
      for (int64_t i = lo; i < hi; ++i) R->clique_nodes.push_back(cl_nodes[(size_t)i]);
      R->clique_off.push_back((int64_t)R->clique_nodes.size());
      if (hi - lo == 1) { ensure(cl_nodes[(size_t)lo]); continue; }
      pending.reset();
      for (int64_t i = lo; i < hi; ++i) pending.add(cl_nodes[(size_t)i]);              // pending = set(clique)
      for (int64_t i = lo; i < hi; ++i) {                                               // u = clique.popleft()
        const int64_t u = cl_nodes[(size_t)i];
        pending.discard(u);
        const int32_t ui = ensure(u);
        const std::vector<int64_t>& mine = nbrs[(size_t)ui];
        IntSet::difference_dict(pending, (int64_t)mine.size(), mine.data(),
                                [&](int64_t v) {
                                  const int32_t vi = index_of.get(v);
                                  return vi >= 0 && linked.contains(pair_code(ui, vi));
                                },
                                fresh);                                                   // pending - set(graph[u])
        targets.clear();
        fresh.for_each([&](int64_t v) { targets.push_back(v); });
        for (const int64_t v : targets) {                                                // graph.add_edge(u, v)
          const int32_t vi = ensure(v);
          nbrs[(size_t)ui].push_back(v);
          nbrs[(size_t)vi].push_back(u);
          linked.insert(pair_code(ui, vi));
          linked.insert(pair_code(vi, ui));
        }
      }
    }
    R->indptr.push_back(0);
    for (const auto& list : nbrs) {
      R->adj.insert(R->adj.end(), list.begin(), list.end());
      R->indptr.push_back((int64_t)R->adj.size());
    }
  }
  lap("graph assembly");
  return owner.release();
}

// The graph that nx.read_adjlist(nx.write_adjlist(G)) gives (reference recovery.py:26-30): the dump lists every edge once,
// under its endpoint that comes first in G's node order (networkx generate_adjlist skips neighbours already "seen"), and the
// parser adds, line by line, the line's node and then its edges (parse_adjlist: add_node, add_edges_from) -- which fixes the
// node order and every adjacency order of the reloaded graph.  Same simulation on arrays, no text.
static void* tail_reload(const int64_t* nodes, const int64_t* indptr, const int64_t* adj, int64_t n_nodes) {
  std::unique_ptr<TailResult> owner(new TailResult());
  TailResult* R = owner.get();
  int64_t max_id = 0;
  for (int64_t i = 0; i < n_nodes; ++i) max_id = nodes[i] > max_id ? nodes[i] : max_id;
  IdIndex old_index, new_index;
  old_index.init(max_id, n_nodes);
  new_index.init(max_id, n_nodes);
  for (int64_t i = 0; i < n_nodes; ++i) old_index.set(nodes[i], (int32_t)i);
  std::vector<char> seen((size_t)n_nodes, 0);
  std::vector<std::vector<int64_t>> nbrs;
  auto ensure = [&](int64_t id) -> int32_t {
    int32_t idx = new_index.get(id);
    if (idx >= 0) return idx;
    idx = (int32_t)R->nodes.size();
    new_index.set(id, idx);
    R->nodes.push_back(id);
    nbrs.emplace_back();
    return idx;
  };
  for (int64_t i = 0; i < n_nodes; ++i) {
    const int64_t s = nodes[i];
    const int32_t si = ensure(s);                                   // G.add_node(u)
    for (int64_t e = indptr[i]; e < indptr[i + 1]; ++e) {
      const int64_t t = adj[e];
      if (seen[(size_t)old_index.get(t)]) continue;                 // written under the earlier endpoint
      const int32_t ti = ensure(t);                                 // add_edges_from: the target joins the graph if new
      nbrs[(size_t)si].push_back(t);
      nbrs[(size_t)ti].push_back(s);
    }
    seen[(size_t)i] = 1;
  }
  R->indptr.push_back(0);
  for (const auto& list : nbrs) {
    R->adj.insert(R->adj.end(), list.begin(), list.end());
    R->indptr.push_back((int64_t)R->adj.size());
  }
  R->clique_off.push_back(0);
  return owner.release();
}

// no C++ exception crosses the C ABI: NULL = out of memory (the caller falls back to the Python loops)
void* nrp_tail_run(const int64_t* start, const int64_t* size, int64_t n_tuples, const int64_t* members) {
  try { return tail_run(start, size, n_tuples, members); } catch (...) { return nullptr; }
}
void* nrp_tail_reload(const int64_t* nodes, const int64_t* indptr, const int64_t* adj, int64_t n_nodes) {
  try { return tail_reload(nodes, indptr, adj, n_nodes); } catch (...) { return nullptr; }
}

void nrp_tail_sizes(void* handle, int64_t out[4]) {
  const TailResult* R = static_cast<const TailResult*>(handle);
  out[0] = (int64_t)R->nodes.size(); out[1] = (int64_t)R->adj.size();
  out[2] = (int64_t)R->clique_off.size() - 1; out[3] = (int64_t)R->clique_nodes.size();
}

void nrp_tail_fetch(void* handle, int64_t* nodes, int64_t* indptr, int64_t* adj, int64_t* clique_off, int64_t* clique_nodes) {
  const TailResult* R = static_cast<const TailResult*>(handle);
  auto copy = [](const std::vector<int64_t>& v, int64_t* dst) { if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(int64_t) * v.size()); };
  copy(R->nodes, nodes); copy(R->indptr, indptr); copy(R->adj, adj); copy(R->clique_off, clique_off); copy(R->clique_nodes, clique_nodes);
}

void nrp_tail_free(void* handle) { delete static_cast<TailResult*>(handle); }

}  // extern "C"
