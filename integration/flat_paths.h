// The C ABI's walk layout and the list bookkeeping of the drop-in ProbCalculator (integration/prob_calculator.h), free of
// any reference type so that it can be unit-tested on its own (tests/cpp/test_flat_paths.cc): flattening a
// vector<vector<int>> walk list, aligning it with the previous call's list, and the multiset difference of two lists
// (what gaml_calc_prob_batch takes for a candidate: indices of the current walks it drops + the walks it adds).
#ifndef GAML_B200_FLAT_PATHS_H__
#define GAML_B200_FLAT_PATHS_H__

#include <cstdint>
#include <cstring>
#include <vector>

namespace gaml_flat {
using std::vector;

// The walk list as the C ABI takes it (and as the adapter remembers the previous call).
struct FlatPaths {
  vector<int32_t> nodes;
  vector<int64_t> offs;   // n + 1
  size_t n() const { return offs.empty() ? 0 : offs.size() - 1; }
  void swap(FlatPaths& o) { nodes.swap(o.nodes); offs.swap(o.offs); }
};
inline void Flatten(const vector<vector<int>>& paths, FlatPaths& f) {
  size_t total = 0;
  for (size_t p = 0; p < paths.size(); p++) total += paths[p].size();
  f.nodes.resize(total ? total : 1);
  f.offs.resize(paths.size() + 1);
  int32_t* out = f.nodes.data();
  size_t at = 0;
  f.offs[0] = 0;
  for (size_t p = 0; p < paths.size(); p++) {
    const vector<int>& w = paths[p];
    if (!w.empty()) memcpy(out + at, w.data(), w.size() * sizeof(int32_t));
    at += w.size();
    f.offs[p + 1] = (int64_t)at;
  }
}
inline bool SameWalk(const FlatPaths& a, size_t x, const FlatPaths& b, size_t y) {
  const int64_t la = a.offs[x + 1] - a.offs[x], lb = b.offs[y + 1] - b.offs[y];
  return la == lb && (la == 0 || memcmp(a.nodes.data() + a.offs[x], b.nodes.data() + b.offs[y], (size_t)la * sizeof(int32_t)) == 0);
}
// match[y] = index of the equal walk of `old` that new walk y is aligned with, or -1: two cursors, runs of equal walks
// compared in blocks on the flat arrays, resynchronisation within a few walks after an edited / removed / inserted
// walk (indices shift when a move erases a walk: comparing position by position would miss everything behind it).
inline void Align(const FlatPaths& old, const FlatPaths& cur, vector<int>& match) {
  const size_t no = old.n(), nc = cur.n();
  match.assign(nc, -1);
  size_t x = 0, y = 0;
  const size_t kBlock = 128;
  while (x < no && y < nc) {
    if (x + kBlock <= no && y + kBlock <= nc) {   // a whole block of equal walks: same boundaries (shifted), same nodes
      const int64_t shift = cur.offs[y] - old.offs[x];
      int64_t diff = 0;
      for (size_t i = 1; i <= kBlock; i++) diff |= (cur.offs[y + i] - old.offs[x + i]) ^ shift;
      if (diff == 0 && memcmp(cur.nodes.data() + cur.offs[y], old.nodes.data() + old.offs[x],
                              (size_t)(cur.offs[y + kBlock] - cur.offs[y]) * sizeof(int32_t)) == 0) {
        for (size_t i = 0; i < kBlock; i++) match[y + i] = (int)(x + i);
        x += kBlock;
        y += kBlock;
        continue;
      }
    }
    if (SameWalk(old, x, cur, y)) {
      match[y++] = (int)x++;
      continue;
    }
    size_t bdx = 0, bdy = 0;
    bool found = false;
    for (size_t dist = 1; dist <= 6 && !found; dist++)
      for (size_t dx = 0; dx <= dist && !found; dx++) {
        const size_t dy = dist - dx;
        if (x + dx < no && y + dy < nc && SameWalk(old, x + dx, cur, y + dy)) { bdx = dx; bdy = dy; found = true; }
      }
    if (!found) break;   // everything from here on counts as unmatched
    x += bdx;
    y += bdy;
  }
}

// Multiset difference `cand` minus/plus `base` from their alignment (match[y] = base index aligned with candidate walk y,
// -1 = none): base walks without a partner are appended to erased_idx unless an unaligned candidate walk equals them
// (a walk that only moved), the remaining unaligned candidate walks are appended to (added_nodes, added_walk_off).
// un_base / un_cand: scratch.
inline void Difference(const FlatPaths& base, const FlatPaths& cand, const vector<int>& match, vector<int>& un_base, vector<int>& un_cand,
                       vector<int32_t>& erased_idx, vector<int32_t>& added_nodes, vector<int64_t>& added_walk_off) {
  un_cand.clear();
  un_base.clear();
  size_t x = 0;   // base walks between two matched ones are unmatched
  for (size_t y = 0; y < match.size(); y++) {
    if (match[y] < 0) { un_cand.push_back((int)y); continue; }
    for (; x < (size_t)match[y]; x++) un_base.push_back((int)x);
    x = (size_t)match[y] + 1;
  }
  for (; x < base.n(); x++) un_base.push_back((int)x);
  vector<char> cand_taken(un_cand.size(), 0);
  for (size_t i = 0; i < un_base.size(); i++) {
    bool kept = false;
    for (size_t j = 0; j < un_cand.size() && !kept; j++)
      if (!cand_taken[j] && SameWalk(base, (size_t)un_base[i], cand, (size_t)un_cand[j])) { cand_taken[j] = 1; kept = true; }
    if (!kept) erased_idx.push_back((int32_t)un_base[i]);
  }
  for (size_t j = 0; j < un_cand.size(); j++) {
    if (cand_taken[j]) continue;
    const size_t y = (size_t)un_cand[j];
    added_nodes.insert(added_nodes.end(), cand.nodes.begin() + cand.offs[y], cand.nodes.begin() + cand.offs[y + 1]);
    added_walk_off.push_back((int64_t)added_nodes.size());
  }
}

}  // namespace gaml_flat
#endif
