// Layout-exact model of CPython's set (Objects/setobject.c, CPython 3.8 - 3.13) for keys that are small non-negative ints
// or tuples of them: same probing (LINEAR_PROBES = 9, PERTURB_SHIFT = 5), same resize rule (fill*5 >= mask*3 -> 4x / 2x of
// `used`), same dummy handling, same bulk operations (merge, intersection, difference and their size heuristics), so that the
// ITERATION ORDER of every set equals the interpreter's.  Nothing here creates Python objects: the host tail of the Finder
// graph build (reference nrpcalc/base/hgraph.py:96-200) runs on these models instead of on millions of small objects.
// Checked against the running interpreter by tests/test_cpyset.py (randomised operation sequences) and by the self-check in
// nrpcalc_b200/native_tail.py, which disables this path if the interpreter's layout rules ever differ.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace cpyset {

constexpr int kLinearProbes = 9;
constexpr int kPerturbShift = 5;
constexpr int64_t kEmpty = -1, kDummy = -2;      // slot states; active slots hold the key itself (hash(int) == int below 2^61 - 1)

struct IntSet {
  std::vector<int64_t> table;                      // mask + 1 slots
  size_t mask = 7;
  int64_t fill = 0, used = 0;                      // fill = active + dummy
  IntSet() : table(8, kEmpty) {}

  void reset() { table.assign(8, kEmpty); mask = 7; fill = used = 0; }

  // set_insert_clean: no dummies, key known to be absent
  static void insert_clean(int64_t* t, size_t mask, int64_t key) {
    size_t perturb = (size_t)key, i = (size_t)key & mask;
    for (;;) {
      if (t[i] == kEmpty) { t[i] = key; return; }
      if (i + kLinearProbes <= mask)
        for (int j = 1; j <= kLinearProbes; ++j)
          if (t[i + j] == kEmpty) { t[i + j] = key; return; }
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }

  // set_table_resize: smallest power of two > minused, active entries re-inserted in slot order, dummies dropped
  void resize(int64_t minused) {
    size_t newsize = 8;
    while (newsize <= (size_t)minused) newsize <<= 1;
    std::vector<int64_t> fresh(newsize, kEmpty);
    for (size_t s = 0; s <= mask; ++s)
      if (table[s] >= 0) insert_clean(fresh.data(), newsize - 1, table[s]);
    table.swap(fresh);
    mask = newsize - 1;
    fill = used;
  }

  // set_lookkey: slot of the key, or of the unused slot that ends its probe sequence
  size_t lookup(int64_t key) const {
    size_t perturb = (size_t)key, i = (size_t)key & mask;
    for (;;) {
      int probes = (i + kLinearProbes <= mask) ? kLinearProbes : 0;
      size_t e = i;
      do {
        if (table[e] == kEmpty || table[e] == key) return e;
        ++e;
      } while (probes--);
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
  bool contains(int64_t key) const { return table[lookup(key)] == key; }

  // set_add_entry
  void add(int64_t key) {
    size_t perturb = (size_t)key, i = (size_t)key & mask;
    int64_t freeslot = -1;
    for (;;) {
      int probes = (i + kLinearProbes <= mask) ? kLinearProbes : 0;
      size_t e = i;
      do {
        const int64_t k = table[e];
        if (k == kEmpty) {
          if (freeslot >= 0) { table[(size_t)freeslot] = key; ++used; return; }
          table[e] = key; ++fill; ++used;
          if ((size_t)fill * 5 < mask * 3) return;
          resize(used > 50000 ? used * 2 : used * 4);
          return;
        }
        if (k == key) return;
        if (k == kDummy) freeslot = (int64_t)e;
        ++e;
      } while (probes--);
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }

  // set_discard_entry: the slot becomes a dummy (fill unchanged)
  bool discard(int64_t key) {
    const size_t e = lookup(key);
    if (table[e] != key) return false;
    table[e] = kDummy;
    --used;
    return true;
  }

  // set_merge: self.update(other) / self |= other for a set operand
  void merge(const IntSet& other) {
    if (&other == this || other.used == 0) return;
    if ((size_t)(fill + other.used) * 5 >= mask * 3) resize((used + other.used) * 2);
    if (fill == 0 && mask == other.mask && other.fill == other.used) {       // same geometry, no dummies: slot-for-slot copy
      table = other.table;
      fill = other.fill; used = other.used;
      return;
    }
    if (fill == 0) {                                                         // empty target: clean inserts in the source's slot order
      fill = used = other.used;
      for (size_t s = 0; s <= other.mask; ++s)
        if (other.table[s] >= 0) insert_clean(table.data(), mask, other.table[s]);
      return;
    }
    for (size_t s = 0; s <= other.mask; ++s)
      if (other.table[s] >= 0) add(other.table[s]);
  }

  // set_intersection (result is a fresh set): walks the SMALLER operand (the right one on ties) in slot order
  static void intersection(const IntSet& so, const IntSet& other, IntSet& result) {
    result.reset();
    if (&so == &other) { result.merge(so); return; }
    const IntSet* big = &so; const IntSet* small = &other;
    if (other.used > so.used) { big = &other; small = &so; }
    for (size_t s = 0; s <= small->mask; ++s)
      if (small->table[s] >= 0 && big->contains(small->table[s])) result.add(small->table[s]);
  }

  // after bulk discards (set_difference_update_internal): "if more than 1/4th are dummies, then resize them away"
  void shed_dummies() {
    if ((size_t)(fill - used) <= mask / 4) return;
    resize(used > 50000 ? used * 2 : used * 4);
  }

  // set_difference with a set operand (so - other)
  static void difference(const IntSet& so, const IntSet& other, IntSet& result) {
    result.reset();
    if ((so.used >> 2) > other.used) {                                       // set_copy_and_difference
      result.merge(so);
      if (&so == &other) { result.reset(); return; }
      for (size_t s = 0; s <= other.mask; ++s)
        if (other.table[s] >= 0) result.discard(other.table[s]);
      result.shed_dummies();
      return;
    }
    for (size_t s = 0; s <= so.mask; ++s)
      if (so.table[s] >= 0 && !other.contains(so.table[s])) result.add(so.table[s]);
  }

  // set_difference with a dict operand (so.difference(d)): `in_other(key)` = key in d, `other_keys` = its keys (any order)
  template <typename Contains>
  static void difference_dict(const IntSet& so, int64_t other_size, const int64_t* other_keys, Contains in_other, IntSet& result) {
    result.reset();
    if ((so.used >> 2) > other_size) {
      result.merge(so);
      for (int64_t i = 0; i < other_size; ++i) result.discard(other_keys[i]);
      result.shed_dummies();
      return;
    }
    for (size_t s = 0; s <= so.mask; ++s)
      if (so.table[s] >= 0 && !in_other(so.table[s])) result.add(so.table[s]);
  }

  template <typename F> void for_each(F f) const {
    for (size_t s = 0; s <= mask; ++s)
      if (table[s] >= 0) f(table[s]);
  }
};

// hash of a tuple of small non-negative ints (Objects/tupleobject.c, xxHash-style, CPython >= 3.8)
inline int64_t tuple_hash(const int64_t* items, int64_t len) {
  const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P5 = 2870177450012600261ull;
  uint64_t acc = P5;
  for (int64_t i = 0; i < len; ++i) {
    acc += (uint64_t)items[i] * P2;
    acc = (acc << 31) | (acc >> 33);
    acc *= P1;
  }
  acc += (uint64_t)len ^ (P5 ^ 3527539ull);
  if (acc == (uint64_t)-1) return 1546275796;
  return (int64_t)acc;
}

// A set of tuples that only ever grows and is then emptied by pop(): pop order = slot order.  Entries are indices into the
// caller's tuple store; `equal(a, b)` compares two stored tuples.
struct TupleSet {
  std::vector<int64_t> slot_item;     // -1 = unused
  std::vector<int64_t> slot_hash;
  size_t mask = 7;
  int64_t used = 0;
  TupleSet() : slot_item(8, -1), slot_hash(8, 0) {}

  void insert_clean(std::vector<int64_t>& items, std::vector<int64_t>& hashes, size_t m, int64_t item, int64_t hash) const {
    size_t perturb = (size_t)hash, i = (size_t)hash & m;
    for (;;) {
      if (items[i] < 0) { items[i] = item; hashes[i] = hash; return; }
      if (i + kLinearProbes <= m)
        for (int j = 1; j <= kLinearProbes; ++j)
          if (items[i + j] < 0) { items[i + j] = item; hashes[i + j] = hash; return; }
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & m;
    }
  }

  template <typename Equal> void add(int64_t item, int64_t hash, Equal equal) {
    size_t perturb = (size_t)hash, i = (size_t)hash & mask;
    for (;;) {
      int probes = (i + kLinearProbes <= mask) ? kLinearProbes : 0;
      size_t e = i;
      do {
        if (slot_item[e] < 0) {
          slot_item[e] = item; slot_hash[e] = hash; ++used;
          if ((size_t)used * 5 < mask * 3) return;
          const int64_t minused = used > 50000 ? used * 2 : used * 4;
          size_t newsize = 8;
          while (newsize <= (size_t)minused) newsize <<= 1;
          std::vector<int64_t> items(newsize, -1), hashes(newsize, 0);
          for (size_t s = 0; s <= mask; ++s)
            if (slot_item[s] >= 0) insert_clean(items, hashes, newsize - 1, slot_item[s], slot_hash[s]);
          slot_item.swap(items); slot_hash.swap(hashes);
          mask = newsize - 1;
          return;
        }
        if (slot_hash[e] == hash && equal(slot_item[e], item)) return;
        ++e;
      } while (probes--);
      perturb >>= kPerturbShift;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
};

}  // namespace cpyset
