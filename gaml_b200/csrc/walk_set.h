// Host-side walk sets and the GetChanges replica (graph.cc:1745-1764). Header only, no CUDA: also compiled into the CPU
// unit test tests/cpp/test_walk_changes.cc, which checks the fast path below against the reference's own container.
//
// The reference finds the walks an evaluation erased and added with
//     unordered_multiset<vector<int>> idx(old_paths); for (p : new_paths) { find -> erase, or added.push_back(p); }
//     erased = what is left in idx, IN THE CONTAINER'S ITERATION ORDER
// and that order decides in which sequence the erased walks' terms are subtracted from a read's running probability
// (bit-exact replay). get_changes_reference() runs exactly that, with the same hash values, on
// (pointer, length, hash) elements. An annealing step, though, re-submits all walks but one to four unchanged, so
// load_walks() first diffs the new walk list against the previous one (runs of equal walks, resynchronised after each
// edit), get_changes_fast() confines the multiset difference to the few walks that changed and orders two or more
// erased walks by the container's rule instead of
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
  std::vector<uint32_t> bkt;   // n: hash % bkt_count — the walk's bucket when THIS set is the reference's old_paths
  size_t bkt_count = 0;        // bucket_count_for(n)
  int n = 0;
  WalkView view(int i) const { return WalkView{nodes.data() + offs[i], (int)(offs[i + 1] - offs[i]), hash[i], i}; }
  Walk walk(int i) const { return Walk(nodes.begin() + offs[i], nodes.begin() + offs[i + 1]); }
};

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

// The walks of the new list that have no equal partner at the aligned position of the old list, and vice versa
// (ascending indices). valid = the alignment succeeded with few enough changes for the fast path.
struct WalkDiff {
  bool valid = false;
  std::vector<int> old_changed, new_changed;
};
constexpr int kMaxChanged = 64;

// Length of the common prefix of a[0..limit) and b[0..limit) + shift, compared in blocks (a vectorised pass over the
// long agreeing runs instead of a branch per element); common_suffix likewise from the back.
template <class T>
inline int64_t common_prefix(const T* a, const T* b, int64_t limit, T shift) {
  int64_t k = 0;
  constexpr int64_t kBlockLen = 256;
  while (k + kBlockLen <= limit) {
    T diff = 0;
    for (int64_t i = 0; i < kBlockLen; i++) diff |= (T)(a[k + i] ^ (T)(b[k + i] + shift));
    if (diff != 0) break;
    k += kBlockLen;
  }
  while (k < limit && a[k] == (T)(b[k] + shift)) k++;
  return k;
}

// Copies the caller's walk arrays into `cur` and aligns them with `old` (null: no previous list): equal walks keep
// their hashes, the others are hashed and reported in the diff. Two cursors walk both lists; a run of equal walks is
// found on the flat arrays (boundaries and nodes compared in bulk); at a mismatch the cursors resynchronise on the
// nearest equal pair within a few walks (an edited, removed or inserted walk), else the rest counts as changed.
inline void load_walks(WalkSet& cur, const WalkSet* old, const int32_t* nodes, const int64_t* offs, int n_walks, WalkDiff& d) {
  const int64_t base = n_walks > 0 ? offs[0] : 0, total = n_walks > 0 ? offs[n_walks] - base : 0;
  cur.n = n_walks;
  cur.nodes.assign(nodes + base, nodes + base + total);
  cur.offs.resize((size_t)n_walks + 1);
  if (n_walks > 0)
    for (int i = 0; i <= n_walks; i++) cur.offs[i] = offs[i] - base;
  else
    cur.offs[0] = 0;
  cur.hash.resize((size_t)n_walks);
  cur.bkt.resize((size_t)n_walks);
  cur.bkt_count = bucket_count_for((size_t)std::max(n_walks, 1));
  const FastMod mod(cur.bkt_count);
  const bool same_buckets = old && old->bkt_count == cur.bkt_count;   // the prime changes only when n crosses a table entry
  d.valid = false;
  d.old_changed.clear();
  d.new_changed.clear();
  auto hash_new = [&](int i) {
    cur.hash[i] = hash_nodes(cur.nodes.data() + cur.offs[i], (int)(cur.offs[i + 1] - cur.offs[i]));
    cur.bkt[i] = (uint32_t)mod(cur.hash[i]);
  };
  if (!old) {
    for (int i = 0; i < n_walks; i++) hash_new(i);
    return;
  }
  auto equal_at = [&](int x, int y) {   // old[x] == new[y]
    const int64_t lo_ = old->offs[x + 1] - old->offs[x], ln = cur.offs[y + 1] - cur.offs[y];
    return lo_ == ln && (ln == 0 || memcmp(cur.nodes.data() + cur.offs[y], old->nodes.data() + old->offs[x], sizeof(int32_t) * (size_t)ln) == 0);
  };
  int x = 0, y = 0;
  bool overflow = false;
  while (x < old->n && y < n_walks) {
    // run of equal walks from (x, y): boundaries (shifted), then nodes
    const int64_t lim = std::min(old->n - x, n_walks - y);
    const int64_t r1 = common_prefix<int64_t>(cur.offs.data() + y + 1, old->offs.data() + x + 1, lim, cur.offs[y] - old->offs[x]);
    const int64_t span = cur.offs[y + r1] - cur.offs[y];
    const int64_t same = common_prefix<int32_t>(cur.nodes.data() + cur.offs[y], old->nodes.data() + old->offs[x], span, 0);
    int64_t r = r1;
    if (same < span) r = (std::upper_bound(cur.offs.begin() + y, cur.offs.begin() + y + r1 + 1, cur.offs[y] + same) - (cur.offs.begin() + y)) - 1;
    if (r > 0) {
      memcpy(cur.hash.data() + y, old->hash.data() + x, sizeof(size_t) * (size_t)r);
      if (same_buckets) memcpy(cur.bkt.data() + y, old->bkt.data() + x, sizeof(uint32_t) * (size_t)r);
      x += (int)r;
      y += (int)r;
      continue;
    }
    // mismatch at (x, y): nearest equal pair within a few walks
    int best_dx = -1, best_dy = -1;
    for (int sdist = 1; sdist <= 6 && best_dx < 0; sdist++)
      for (int dx = 0; dx <= sdist && best_dx < 0; dx++) {
        const int dy = sdist - dx;
        if (x + dx < old->n && y + dy < n_walks && equal_at(x + dx, y + dy)) { best_dx = dx; best_dy = dy; }
      }
    if (best_dx < 0) break;   // no resynchronisation point: everything from here on counts as changed
    for (int k = 0; k < best_dx; k++) d.old_changed.push_back(x + k);
    for (int k = 0; k < best_dy; k++) { hash_new(y + k); d.new_changed.push_back(y + k); }
    x += best_dx;
    y += best_dy;
    if ((int)d.old_changed.size() > kMaxChanged || (int)d.new_changed.size() > kMaxChanged) { overflow = true; break; }
  }
  if ((old->n - x) > kMaxChanged || (n_walks - y) > kMaxChanged) overflow = true;
  for (; y < n_walks; y++) {
    hash_new(y);
    if (!overflow) d.new_changed.push_back(y);
  }
  if (!overflow)
    for (; x < old->n; x++) d.old_changed.push_back(x);
  if (!same_buckets)
    for (int i = 0; i < n_walks; i++) cur.bkt[i] = (uint32_t)mod(cur.hash[i]);
  d.valid = !overflow && (int)d.old_changed.size() <= kMaxChanged && (int)d.new_changed.size() <= kMaxChanged;
  if (!d.valid) {
    d.old_changed.clear();
    d.new_changed.clear();
  }
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

// Same result as get_changes_reference, from the diff. false = not applicable (equal walks are involved): use the
// reference path. old_counts must describe `old`.
inline bool get_changes_fast(const WalkSet& old, const WalkSet& cur, const WalkDiff& d, const HashCounts& old_counts, Changes& out) {
  if (!d.valid || !old_counts.valid) return false;
  out.erased.clear();
  out.added.clear();
  const std::vector<int>& A = d.old_changed;
  const std::vector<int>& B = d.new_changed;
  // no walk outside the changed ones may share a hash with one of them (then the multiset difference is confined to
  // them: every other walk has an equal partner in the other list)
  auto outside = [&](uint64_t h) {
    int c = old_counts.get(h);
    for (int i : A) c -= old.hash[i] == h;
    return c;
  };
  for (int i : A)
    if (outside(old.hash[i]) != 0) return false;
  for (int i : B)
    if (outside(cur.hash[i]) != 0) return false;
  // among the changed walks: every new one, in order, takes an unmatched equal old walk or is added
  bool taken[kMaxChanged] = {false};
  for (int i : B) {
    const WalkView v = cur.view(i);
    int hit = -1;   // (find() returns the most recently inserted of several equal elements)
    for (int j = (int)A.size() - 1; j >= 0 && hit < 0; j--)
      if (!taken[j] && same_walk(old.view(A[j]), v)) hit = j;
    if (hit >= 0) taken[hit] = true;
    else out.added.push_back(v);
  }
  for (size_t j = 0; j < A.size(); j++)
    if (!taken[j]) out.erased.push_back(old.view(A[j]));
  if (out.erased.size() < 2) return true;
  // an element with an equal one anywhere in the container sits next to it, out of index order: the rule below holds
  // for elements that are unique in the old set
  for (const WalkView& e : out.erased)
    if (old_counts.get(e.h) != 1) return false;
  // the container's iteration order: buckets by the index of their first element (descending), then index (descending)
  if (old.bkt_count != bucket_count_for((size_t)std::max(old.n, 1))) return false;   // (cannot happen: load_walks keeps it)
  const size_t k = out.erased.size();
  uint32_t bkt[kMaxChanged];
  int first[kMaxChanged];
  int last_needed = 0;
  for (size_t x = 0; x < k; x++) {
    bkt[x] = old.bkt[out.erased[x].index];
    first[x] = out.erased[x].index;   // an element is in its own bucket: the bucket's first element is at or before it
    last_needed = std::max(last_needed, out.erased[x].index);
  }
  if (k == 2) {   // the common case (a join, a tail swap), kept tight
    const uint32_t b0 = bkt[0], b1 = bkt[1];
    int f0 = first[0], f1 = first[1];
    const uint32_t* bp = old.bkt.data();
    for (int j = 0; j < last_needed; j++) {
      const uint32_t b = bp[j];
      if (b == b0 && j < f0) f0 = j;
      if (b == b1 && j < f1) f1 = j;
    }
    first[0] = f0;
    first[1] = f1;
  } else {
    for (int j = 0; j < last_needed; j++) {
      const uint32_t b = old.bkt[j];
      for (size_t x = 0; x < k; x++)
        if (b == bkt[x] && j < first[x]) first[x] = j;
    }
  }
  int order[kMaxChanged];
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

// ---- walk labels for patched full evaluations (engine.cu "patched full evaluation") ---------------------------------
// The slot tables of ONE base walk list stay on the device. Every walk of a later list carries a label that orders it
// among the base walks without renumbering them: base walk b has label (b + 1) << g; a walk that is not in the base list
// gets a label strictly between its neighbours' (appended walks count upwards), never a multiple of 2^g. Labels are
// followed from list to list along load_walks' alignment.
struct ListTrack {
  bool valid = false;
  uint64_t base_gen = 0;
  std::vector<uint32_t> label;   // per walk of the list, strictly increasing
  std::vector<int> removed;      // base walks that are not in the list (ascending base index)
  int n_added = 0;               // walks of the list that are not base walks
};

// Labels of the new list `ws` from the previous list's (`old`, labels in tp) along the diff d: equal walks keep theirs,
// run by run. A walk that left the list is noted when it was a base walk; a walk that entered it gets its base label back
// when it equals a missing base walk and that label fits between its neighbours', else a label in the gap between them.
// false (reason in *why) = the next full evaluation renumbers: a gap is used up, the lists do not align, or the list has
// drifted more than max_walks walks away from the base. g / s: label spacing and the enumeration bits below a label
// (labels stay below 2^(32 - s)).
inline bool derive_track(const WalkSet& base, int g, int s, int max_walks, const WalkSet& old, const WalkSet& ws, const WalkDiff& d,
                         const ListTrack& tp, ListTrack& tc, const char** why = nullptr) {
  auto no = [&](const char* reason) {
    if (why) *why = reason;
    return false;
  };
  const uint32_t gmask = (1u << g) - 1u;
  if (!d.valid) return no("no alignment with the previous list");
  if ((int)tp.label.size() != old.n) return no("label array out of step");
  tc.label.resize((size_t)ws.n);
  tc.removed = tp.removed;
  tc.n_added = tp.n_added;
  size_t io = 0, in = 0;
  int x = 0, y = 0;
  while (x < old.n || y < ws.n) {
    const int nx = io < d.old_changed.size() ? d.old_changed[io] : old.n;
    const int ny = in < d.new_changed.size() ? d.new_changed[in] : ws.n;
    const int r = std::min(nx - x, ny - y);
    if (r < 0) return no("negative run");
    if (r > 0) {
      memcpy(tc.label.data() + y, tp.label.data() + x, sizeof(uint32_t) * (size_t)r);
      x += r;
      y += r;
    }
    bool moved = r > 0;
    if (x == nx && x < old.n) {
      const uint32_t lab = tp.label[x];
      if ((lab & gmask) == 0) tc.removed.insert(std::upper_bound(tc.removed.begin(), tc.removed.end(), (int)(lab >> g) - 1), (int)(lab >> g) - 1);
      else tc.n_added--;
      x++;
      io++;
      moved = true;
    }
    if (y == ny && y < ws.n) {
      tc.label[y] = 0xffffffffu;   // assigned below, once its right neighbour's label is known
      y++;
      in++;
      moved = true;
    }
    if (!moved) return no("lists of different length with nothing left to skip");
  }
  const uint32_t label_max = s >= 32 ? 0u : (uint32_t)((1ull << (32 - s)) - 1ull);
  for (size_t k = 0; k < d.new_changed.size(); k++) {
    const int yy = d.new_changed[k];
    const uint32_t left = yy > 0 ? tc.label[yy - 1] : 0u;
    int y2 = yy + 1;
    while (y2 < ws.n && tc.label[y2] == 0xffffffffu) y2++;
    const bool at_end = y2 >= ws.n;   // nothing labelled to the right: walks appended to the list count upwards
    const uint32_t right = at_end ? label_max : tc.label[y2];
    if (right <= left + 1) return no("no label left in the gap");
    uint32_t lab = 0;
    for (size_t q = 0; q < tc.removed.size(); q++) {
      const int b = tc.removed[q];
      const uint32_t bl = (uint32_t)(b + 1) << g;
      if (bl <= left || bl >= right) continue;
      if (base.hash[b] != ws.hash[yy] || !same_walk(base.view(b), ws.view(yy))) continue;
      lab = bl;
      tc.removed.erase(tc.removed.begin() + (long)q);
      break;
    }
    if (!lab) {
      lab = at_end ? left + 1 : left + (right - left) / 2;
      if ((lab & gmask) == 0) {   // multiples of 2^g are the base walks' labels
        if (lab + 1 < right) lab++;
        else if (lab - 1 > left) lab--;
        else return no("the gap holds only a base label");
      }
      tc.n_added++;
    }
    tc.label[yy] = lab;
  }
  if ((int)tc.removed.size() > max_walks || tc.n_added > max_walks) return no("too far from the base list");
  tc.base_gen = tp.base_gen;
  return true;
}

}  // namespace gaml
