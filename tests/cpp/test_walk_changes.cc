// CPU unit test: get_changes_fast (diff + container-order replica, gaml_b200/csrc/walk_set.h) must return exactly what
// the reference's algorithm on the real std::unordered_multiset returns (get_changes_reference), over random annealing
// style trajectories: joins, splits, tail swaps, flips, duplicates, insertions and deletions at any position.
// Second part: the walk labels of patched full evaluations (derive_track).
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

// Walk labels of patched full evaluations (derive_track): over random annealing-style trajectories the labels must stay
// strictly increasing, a walk carrying a base label must BE that base walk, `removed` must list exactly the base walks
// that carry no label in the list, and going back to the base list must restore the base labels.
static int test_labels() {
  std::mt19937_64 rng(777);
  long tracked = 0, renumbered = 0, back_home = 0;
  for (int trial = 0; trial < 200; trial++) {
    const int alphabet = 50 + (int)(rng() % 5000);
    const int n0 = 2 + (int)(rng() % 600);
    std::vector<Walk> base;
    for (int i = 0; i < n0; i++) {
      Walk w(1 + rng() % 3);
      for (int& x : w) x = (int)(rng() % alphabet);
      base.push_back(w);
    }
    const int g = 3;
    int bits = 0;
    while (((uint64_t)(n0 + 2) << g) >> bits) bits++;
    if (32 - (bits + 1) >= 12) bits++;
    const int s = std::min(32 - bits, 20);
    WalkSet sets[2], base_ws;
    ListTrack tracks[2];
    int which = 0;
    WalkDiff d;
    flatten(base);
    load_walks(sets[which], nullptr, g_nodes.data(), g_offs.data(), (int)base.size(), d);
    base_ws = sets[which];
    auto rebase = [&](const WalkSet& ws, ListTrack& t) {
      base_ws = ws;
      t.valid = true;
      t.base_gen++;
      t.removed.clear();
      t.n_added = 0;
      t.label.resize((size_t)ws.n);
      for (int i = 0; i < ws.n; i++) t.label[(size_t)i] = (uint32_t)(i + 1) << g;
    };
    rebase(sets[which], tracks[which]);
    std::vector<Walk> cur = base, home = base;
    for (int step = 0; step < 80; step++) {
      std::vector<Walk> nw = cur;
      const int kind = (int)(rng() % 7);
      auto pick = [&]() { return (size_t)(rng() % nw.size()); };
      if (kind == 0 && nw.size() >= 2) {            // join (replace one, erase the other)
        size_t i = pick(), j = pick();
        if (i != j) { nw[i].insert(nw[i].end(), nw[j].begin(), nw[j].end()); nw.erase(nw.begin() + (long)j); }
      } else if (kind == 1) {                       // split: second part appended
        size_t i = pick();
        if (nw[i].size() >= 2) { Walk tail(nw[i].begin() + 1, nw[i].end()); nw[i].resize(1); nw.push_back(tail); }
      } else if (kind == 2) {                       // edit in place
        size_t i = pick();
        nw[i].push_back((int)(rng() % alphabet));
      } else if (kind == 3) {                       // insert in the middle
        nw.insert(nw.begin() + (long)(rng() % (nw.size() + 1)), Walk(1, (int)(rng() % alphabet)));
      } else if (kind == 4 && nw.size() >= 2) {     // erase
        nw.erase(nw.begin() + (long)pick());
      } else if (kind == 5) {                       // back to the base list
        nw = home;
      }                                             // kind 6: the same list again
      if (nw.empty()) nw.push_back(Walk(1, 0));
      flatten(nw);
      WalkSet& nxt = sets[which ^ 1];
      load_walks(nxt, &sets[which], g_nodes.data(), g_offs.data(), (int)nw.size(), d);
      ListTrack& tc = tracks[which ^ 1];
      const ListTrack& tp = tracks[which];
      tc.valid = tp.valid && d.valid && derive_track(base_ws, g, s, 64, sets[which], nxt, d, tp, tc);
      if (tc.valid) {
        tracked++;
        const uint32_t gmask = (1u << g) - 1u, label_max = (uint32_t)((1ull << (32 - s)) - 1ull);
        std::vector<char> present((size_t)base_ws.n, 0);
        int added = 0;
        for (int i = 0; i < nxt.n; i++) {
          const uint32_t lab = tc.label[(size_t)i];
          if (i > 0 && tc.label[(size_t)i - 1] >= lab) { printf("FAIL: labels not increasing (trial %d step %d)\n", trial, step); return 1; }
          if (lab == 0 || lab > label_max) { printf("FAIL: label out of range\n"); return 1; }
          if ((lab & gmask) == 0) {
            const int b = (int)(lab >> g) - 1;
            if (b < 0 || b >= base_ws.n || !same_walk(base_ws.view(b), nxt.view(i))) { printf("FAIL: base label on another walk\n"); return 1; }
            present[(size_t)b] = 1;
          } else {
            added++;
          }
        }
        if (added != tc.n_added) { printf("FAIL: n_added %d != %d\n", tc.n_added, added); return 1; }
        std::vector<int> gone;
        for (int b = 0; b < base_ws.n; b++)
          if (!present[(size_t)b]) gone.push_back(b);
        if (gone != tc.removed) { printf("FAIL: removed list (trial %d step %d)\n", trial, step); return 1; }
        if (nw == home && tc.n_added == 0 && tc.removed.empty()) back_home++;
      } else {
        renumbered++;
        rebase(nxt, tc);   // what a full evaluation does when the labels cannot be kept
        home = nw;
      }
      which ^= 1;
      cur.swap(nw);
    }
  }
  printf("OK labels: %ld lists tracked, %ld renumbered, %ld returns to the base list recognised\n", tracked, renumbered, back_home);
  return tracked > 5 * renumbered && back_home > 100 ? 0 : 1;
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
  if (fast_used * 2 <= checked) return 1;
  return test_labels();
}
