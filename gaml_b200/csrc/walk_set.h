// Host-side walk sets and the GetChanges replica (graph.cc:1745-1764). Header only, no CUDA: also compiled into the CPU
// unit test tests/cpp/test_walk_changes.cc, which checks the fast path below against the reference's own container.
//
// The reference finds the walks an evaluation erased and added with
//     unordered_multiset<vector<int>> idx(old_paths); for (p : new_paths) { find -> erase, or added.push_back(p); }
//     erased = what is left in idx, IN THE CONTAINER'S ITERATION ORDER
// and that order decides in which sequence the erased walks' terms are subtracted from a read's running probability
// (bit-exact replay). get_changes_reference() runs exactly that, with the same hash values, on
// (pointer, length, hash) elements. An annealing step, though, re-submits all walks but one to four unchanged, so
// get_changes_fast() first diffs the new walk list against the previous one (common prefix / suffix), confines the
// multiset difference to the changed region, and orders two or more erased walks by the container's rule instead of
// building it: libstdc++ chains all nodes in one list; a bucket's nodes sit together, a new node goes to the FRONT of
// its bucket, and a bucket that receives its first node goes to the front of the whole list. With no equal elements
// involved (checked; otherwise the reference path runs) the iteration order therefore is: buckets by the index of
// their first element, descending; inside a bucket by index, descending.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory_resource>
#include <unordered_set>
#include <vector>

namespace gaml {

using Walk = std::vector<int>;

// graph.h:21-45: the reference's hash for vector<int> (hash_combine over std::hash<int>)
inline size_t hash_nodes(const int* p, int n) {
  size_t seed = 0;
  for (int i = 0; i < n; i++) seed ^= std::hash<int>()(p[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
  return seed;
}

struct WalkView {
  const int* p;
  int n;
  size_t h;
  int index;   // position in its walk set
};
inline bool same_walk(const WalkView& a, const WalkView& b) {
  return a.n == b.n && a.h == b.h && (a.n == 0 || memcmp(a.p, b.p, sizeof(int) * (size_t)a.n) == 0);
}

struct WalkSet {
  std::vector<int32_t> nodes;
  std::vector<int64_t> offs;   // n + 1
  std::vector<size_t> hash;    // n
  int n = 0;
  WalkView view(int i) const { return WalkView{nodes.data() + offs[i], (int)(offs[i + 1] - offs[i]), hash[i], i}; }
  Walk walk(int i) const { return Walk(nodes.begin() + offs[i], nodes.begin() + offs[i + 1]); }
};

// cur = old[0, prefix) + NEW REGION + old[n_old - suffix, n_old); the old region is old[prefix, n_old - suffix)
struct WalkDiff {
  bool valid = false;
  int prefix = 0, suffix = 0;
};

// Copies the caller's walk arrays into `cur`; hashes only the walks that differ from `old` (null: all of them).
inline WalkDiff load_walks(WalkSet& cur, const WalkSet* old, const int32_t* nodes, const int64_t* offs, int n_walks) {
  const int64_t base = n_walks > 0 ? offs[0] : 0, total = n_walks > 0 ? offs[n_walks] - base : 0;
  cur.n = n_walks;
  cur.nodes.assign(nodes + base, nodes + base + total);
  cur.offs.resize((size_t)n_walks + 1);
  for (int i = 0; i <= n_walks; i++) cur.offs[i] = (n_walks > 0 ? offs[i] : 0) - base;
  cur.hash.resize((size_t)n_walks);
  WalkDiff d;
  int lo = 0, hi_new = n_walks;
  if (old) {
    auto equal_at = [&](int i_new, int i_old) {
      const int64_t ln = cur.offs[i_new + 1] - cur.offs[i_new], lo_ = old->offs[i_old + 1] - old->offs[i_old];
      return ln == lo_ && (ln == 0 || memcmp(cur.nodes.data() + cur.offs[i_new], old->nodes.data() + old->offs[i_old],
                                            sizeof(int32_t) * (size_t)ln) == 0);
    };
    const int m = std::min(n_walks, old->n);
    int p = 0;
    while (p < m && equal_at(p, p)) p++;
    int q = 0;
    while (q < m - p && equal_at(n_walks - 1 - q, old->n - 1 - q)) q++;
    for (int i = 0; i < p; i++) cur.hash[i] = old->hash[i];
    for (int i = 0; i < q; i++) cur.hash[n_walks - 1 - i] = old->hash[old->n - 1 - i];
    d.valid = true;
    d.prefix = p;
    d.suffix = q;
    lo = p;
    hi_new = n_walks - q;
  }
  for (int i = lo; i < hi_new; i++) cur.hash[i] = hash_nodes(cur.nodes.data() + cur.offs[i], (int)(cur.offs[i + 1] - cur.offs[i]));
  return d;
}

// How many walks of a set carry a given hash (open addressing; an entry whose count drops to zero stays as a tombstone
// until the next rebuild).
struct HashCounts {
  std::vector<uint64_t> key;
  std::vector<int32_t> cnt;
  size_t mask = 0, used = 0;
  bool valid = false;
  void rebuild(const WalkSet& s) {
    size_t cap = 64;
    while (cap < (size_t)s.n * 4 + 16) cap <<= 1;
    key.assign(cap, 0);
    cnt.assign(cap, -1);   // -1 = never used
    mask = cap - 1;
    used = 0;
    for (int i = 0; i < s.n; i++) add(s.hash[i], +1);
    valid = true;
  }
  size_t slot(uint64_t h) const {
    size_t i = (size_t)(h * 0x9e3779b97f4a7c15ull >> 17) & mask;
    while (cnt[i] >= 0 && key[i] != h) i = (i + 1) & mask;
    return i;
  }
  void add(uint64_t h, int delta) {
    const size_t i = slot(h);
    if (cnt[i] < 0) {
      key[i] = h;
      cnt[i] = 0;
      used++;
    }
    cnt[i] += delta;
  }
  int get(uint64_t h) const {
    const size_t i = slot(h);
    return cnt[i] < 0 ? 0 : cnt[i];
  }
  bool crowded() const { return used * 2 > mask; }
};

struct Changes {
  std::vector<WalkView> erased, added;   // erased: views into the OLD set, in the reference's order; added: into the new set
};

// The reference's algorithm on its own container (bump-allocated nodes).
struct WalkViewHash { size_t operator()(const WalkView& r) const { return r.h; } };
struct WalkViewEq { bool operator()(const WalkView& a, const WalkView& b) const { return same_walk(a, b); } };
inline void get_changes_reference(const WalkSet& old, const WalkSet& cur, Changes& out, std::vector<WalkView>& scratch,
                                  std::pmr::memory_resource* pool) {
  out.erased.clear();
  out.added.clear();
  scratch.clear();
  for (int i = 0; i < old.n; i++) scratch.push_back(old.view(i));
  std::pmr::unordered_multiset<WalkView, WalkViewHash, WalkViewEq> idx(scratch.begin(), scratch.end(), 0, WalkViewHash(),
                                                                       WalkViewEq(), pool);
  for (int i = 0; i < cur.n; i++) {
    const WalkView v = cur.view(i);
    auto f = idx.find(v);
    if (f == idx.end()) out.added.push_back(v);
    else idx.erase(f);
  }
  for (const WalkView& r : idx) out.erased.push_back(r);
}

// x mod d for a fixed d (Lemire, Kaser, Kurz: "Faster remainder by direct computation", 64-bit version)
struct FastMod {
  unsigned __int128 M;
  uint64_t d;
  explicit FastMod(uint64_t d_) : M(~(unsigned __int128)0 / d_ + 1), d(d_) {}
  uint64_t operator()(uint64_t a) const {
    const unsigned __int128 low = M * a;
    const unsigned __int128 bottom = (unsigned __int128)(uint64_t)low * d;
    const unsigned __int128 top = (unsigned __int128)(uint64_t)(low >> 64) * d;
    return (uint64_t)((top + (bottom >> 64)) >> 64);
  }
};

// Bucket count libstdc++ gives unordered_multiset(first, last) for n elements.
inline size_t bucket_count_for(size_t n) {
  std::__detail::_Prime_rehash_policy pol;
  return pol._M_next_bkt(pol._M_bkt_for_elements(n));
}

// Same result as get_changes_reference, from the diff. false = not applicable (equal walks are involved): use the
// reference path. old_counts must describe `old`.
inline bool get_changes_fast(const WalkSet& old, const WalkSet& cur, const WalkDiff& d, const HashCounts& old_counts, Changes& out) {
  if (!d.valid || !old_counts.valid) return false;
  out.erased.clear();
  out.added.clear();
  const int a0 = d.prefix, a1 = old.n - d.suffix, b0 = d.prefix, b1 = cur.n - d.suffix;
  if (a1 - a0 > 64 || b1 - b0 > 64) return false;   // a wholesale change: the quadratic matching below is not meant for it
  // no walk outside the changed region may share a hash with one inside it (then the multiset difference is confined
  // to the region, and no erased walk has an equal one elsewhere in the container)
  auto outside = [&](uint64_t h) {
    int c = old_counts.get(h);
    for (int i = a0; i < a1; i++) c -= old.hash[i] == h;
    return c;
  };
  for (int i = a0; i < a1; i++)
    if (outside(old.hash[i]) != 0) return false;
  for (int i = b0; i < b1; i++)
    if (outside(cur.hash[i]) != 0) return false;
  // inside the region: every new walk, in order, takes an unmatched equal old walk or is added
  bool taken[64] = {false};
  for (int i = b0; i < b1; i++) {
    const WalkView v = cur.view(i);
    int hit = -1;   // (find() returns the most recently inserted of several equal elements)
    for (int j = a1 - 1; j >= a0 && hit < 0; j--)
      if (!taken[j - a0] && same_walk(old.view(j), v)) hit = j;
    if (hit >= 0) taken[hit - a0] = true;
    else out.added.push_back(v);
  }
  for (int j = a0; j < a1; j++)
    if (!taken[j - a0]) out.erased.push_back(old.view(j));
  if (out.erased.size() < 2) return true;
  // an element with an equal one anywhere in the container sits next to it, out of index order: the rule below holds
  // for elements that are unique in the old set
  for (const WalkView& e : out.erased)
    if (old_counts.get(e.h) != 1) return false;
  // the container's iteration order: buckets by the index of their first element (descending), then index (descending)
  const FastMod mod(bucket_count_for((size_t)old.n));
  const size_t k = out.erased.size();
  uint64_t bkt[64];
  int first[64];
  int last_needed = 0;
  for (size_t x = 0; x < k; x++) {
    bkt[x] = mod(out.erased[x].h);
    first[x] = out.erased[x].index;   // an element is in its own bucket: the bucket's first element is at or before it
    last_needed = std::max(last_needed, out.erased[x].index);
  }
  for (int j = 0; j < last_needed; j++) {
    const uint64_t b = mod(old.hash[j]);
    for (size_t x = 0; x < k; x++)
      if (b == bkt[x] && j < first[x]) first[x] = j;
  }
  int order[64];
  for (size_t x = 0; x < k; x++) order[x] = (int)x;
  std::sort(order, order + k, [&](int x, int y) {
    if (first[x] != first[y]) return first[x] > first[y];
    return out.erased[x].index > out.erased[y].index;
  });
  std::vector<WalkView> sorted(k);
  for (size_t x = 0; x < k; x++) sorted[x] = out.erased[order[x]];
  out.erased.swap(sorted);
  return true;
}

}  // namespace gaml
