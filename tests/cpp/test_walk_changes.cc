// CPU unit test: get_changes_fast (diff + container-order replica, gaml_b200/csrc/walk_set.h) must return exactly what
// the reference's algorithm on the real std::unordered_multiset returns (get_changes_reference), over random annealing
// style trajectories: joins, splits, tail swaps, flips, duplicates, insertions and deletions at any position.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../gaml_b200/csrc/walk_set.h"

using namespace gaml;

static std::vector<int32_t> g_nodes;
static std::vector<int64_t> g_offs;
static void flatten(const std::vector<Walk>& ws) {
  g_nodes.clear();
  g_offs.assign(1, 0);
  for (const Walk& w : ws) {
    g_nodes.insert(g_nodes.end(), w.begin(), w.end());
    g_offs.push_back((int64_t)g_nodes.size());
  }
  if (g_nodes.empty()) g_nodes.push_back(0);
}

int main() {
  std::mt19937_64 rng(12345);
  long checked = 0, fast_used = 0, multi_erased = 0;
  for (int trial = 0; trial < 300; trial++) {
    const int alphabet = 4 + (int)(rng() % 2000);
    int n0 = 1 + (int)(rng() % 400);
    std::vector<Walk> cur;
    for (int i = 0; i < n0; i++) {
      Walk w(1 + rng() % 4);
      for (int& x : w) x = (int)(rng() % alphabet);
      cur.push_back(w);
    }
    WalkSet sets[2];
    int which = 0;
    flatten(cur);
    WalkDiff d0;
    load_walks(sets[which], nullptr, g_nodes.data(), g_offs.data(), (int)cur.size(), d0);
    HashCounts counts;
    counts.rebuild(sets[which]);
    std::pmr::monotonic_buffer_resource pool(1 << 16);
    std::vector<WalkView> scratch;
    for (int step = 0; step < 60; step++) {
      std::vector<Walk> nw = cur;
      const int kind = (int)(rng() % 8);
      auto pick = [&]() { return (size_t)(rng() % nw.size()); };
      if (kind == 0 && nw.size() >= 2) {           // join two walks
        size_t i = pick(), j = pick();
        if (i != j) {
          nw[i].insert(nw[i].end(), nw[j].begin(), nw[j].end());
          nw.erase(nw.begin() + (long)j);
        }
      } else if (kind == 1) {                      // split
        size_t i = pick();
        if (nw[i].size() >= 2) {
          size_t c = 1 + rng() % (nw[i].size() - 1);
          Walk tail(nw[i].begin() + (long)c, nw[i].end());
          nw[i].resize(c);
          nw.insert(nw.begin() + (long)(rng() % (nw.size() + 1)), tail);
        }
      } else if (kind == 2 && nw.size() >= 2) {    // swap tails
        size_t i = pick(), j = pick();
        if (i != j) {
          size_t ci = rng() % (nw[i].size() + 1), cj = rng() % (nw[j].size() + 1);
          Walk a(nw[i].begin(), nw[i].begin() + (long)ci), b(nw[j].begin(), nw[j].begin() + (long)cj);
          a.insert(a.end(), nw[j].begin() + (long)cj, nw[j].end());
          b.insert(b.end(), nw[i].begin() + (long)ci, nw[i].end());
          if (!a.empty() && !b.empty()) { nw[i] = a; nw[j] = b; }
        }
      } else if (kind == 3) {                      // flip
        size_t i = pick();
        std::reverse(nw[i].begin(), nw[i].end());
        for (int& x : nw[i]) x ^= 1;
      } else if (kind == 4) {                      // duplicate a walk somewhere
        nw.insert(nw.begin() + (long)(rng() % (nw.size() + 1)), nw[pick()]);
      } else if (kind == 5 && nw.size() >= 2) {    // delete
        nw.erase(nw.begin() + (long)pick());
      } else if (kind == 6) {                      // several independent edits
        for (int t = 0; t < 3; t++) nw[pick()].push_back((int)(rng() % alphabet));
      } else {                                     // nothing
      }
      flatten(nw);
      WalkSet& old = sets[which];
      WalkSet& nxt = sets[which ^ 1];
      WalkDiff d;
      load_walks(nxt, &old, g_nodes.data(), g_offs.data(), (int)nw.size(), d);
      for (int i = 0; i < nxt.n; i++)
        if (nxt.hash[i] != hash_nodes(nxt.nodes.data() + nxt.offs[i], (int)(nxt.offs[i + 1] - nxt.offs[i]))) {
          printf("FAIL: hash carried over wrongly (trial %d step %d walk %d)\n", trial, step, i);
          return 1;
        }
      Changes ref, fast;
      pool.release();
      get_changes_reference(old, nxt, ref, scratch, &pool);
      const bool ok = get_changes_fast(old, nxt, d, counts, fast);
      checked++;
      if (ref.erased.size() >= 2) multi_erased++;
      if (ok) {
        fast_used++;
        bool same = ref.erased.size() == fast.erased.size() && ref.added.size() == fast.added.size();
        for (size_t i = 0; same && i < ref.erased.size(); i++) same = same_walk(ref.erased[i], fast.erased[i]);
        for (size_t i = 0; same && i < ref.added.size(); i++) same = same_walk(ref.added[i], fast.added[i]) && ref.added[i].index == fast.added[i].index;
        if (!same) {
          printf("FAIL: trial %d step %d kind %d: erased %zu/%zu added %zu/%zu (n_old %d)\n", trial, step, kind, ref.erased.size(),
                 fast.erased.size(), ref.added.size(), fast.added.size(), old.n);
          for (auto& e : ref.erased) printf("  ref erased index %d\n", e.index);
          for (auto& e : fast.erased) printf("  fast erased index %d\n", e.index);
          return 1;
        }
      }
      // the evaluated set becomes the old one: update the counts by the diff, like the engine does
      if (d.valid && !counts.crowded()) {
        for (int i : d.old_changed) counts.add(old.hash[i], -1);
        for (int i : d.new_changed) counts.add(nxt.hash[i], +1);
      } else {
        counts.rebuild(nxt);
      }
      for (int i = 0; i < nxt.n; i++)
        if (counts.get(nxt.hash[i]) < 1) { printf("FAIL: counts lost a walk\n"); return 1; }
      which ^= 1;
      cur.swap(nw);
    }
  }
  printf("OK: %ld evaluations, fast path %ld (%.0f%%), %ld with two or more erased walks\n", checked, fast_used,
         100.0 * (double)fast_used / (double)checked, multi_erased);
  return fast_used * 2 > checked ? 0 : 1;
}
