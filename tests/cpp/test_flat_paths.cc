// CPU unit test of the drop-in ProbCalculator's list bookkeeping (integration/flat_paths.h): Align must only pair equal
// walks, monotonically; Difference must yield exactly the multiset difference of the two lists (what
// gaml_calc_prob_batch takes for a candidate) over random annealing-style edits, including index shifts after an erase,
// moved walks and duplicates.
#include <algorithm>
#include <cstdio>
#include <map>
#include <random>

#include "../../integration/flat_paths.h"

using namespace gaml_flat;
typedef std::vector<int> Walk;

int main() {
  std::mt19937_64 rng(4242);
  long checked = 0, aligned_walks = 0, total_walks = 0;
  for (int trial = 0; trial < 400; trial++) {
    const int alphabet = 3 + (int)(rng() % 3000);
    const int n0 = 1 + (int)(rng() % 700);
    std::vector<Walk> base;
    for (int i = 0; i < n0; i++) {
      Walk w(1 + rng() % 3);
      for (int& x : w) x = (int)(rng() % alphabet);
      base.push_back(w);
    }
    for (int step = 0; step < 30; step++) {
      std::vector<Walk> cand = base;
      const int n_edits = 1 + (int)(rng() % 3);
      for (int e = 0; e < n_edits; e++) {
        const int kind = (int)(rng() % 7);
        auto pick = [&]() { return (size_t)(rng() % cand.size()); };
        if (cand.empty()) { cand.push_back(Walk(1, 1)); continue; }
        if (kind == 0 && cand.size() >= 2) {            // join: replace one, erase the other (indices behind it shift)
          size_t i = pick(), j = pick();
          if (i != j) { cand[i].insert(cand[i].end(), cand[j].begin(), cand[j].end()); cand.erase(cand.begin() + (long)j); }
        } else if (kind == 1) {                         // split, tail appended
          size_t i = pick();
          if (cand[i].size() >= 2) { Walk t(cand[i].begin() + 1, cand[i].end()); cand[i].resize(1); cand.push_back(t); }
        } else if (kind == 2) {                         // insert anywhere
          cand.insert(cand.begin() + (long)(rng() % (cand.size() + 1)), Walk(1, (int)(rng() % alphabet)));
        } else if (kind == 3) {                         // erase
          cand.erase(cand.begin() + (long)pick());
        } else if (kind == 4) {                         // move a walk to the end (equal walk, other position)
          size_t i = pick();
          Walk w = cand[i];
          cand.erase(cand.begin() + (long)i);
          cand.push_back(w);
        } else if (kind == 5) {                         // duplicate a walk
          cand.push_back(cand[pick()]);
        } else {                                        // edit in place
          cand[pick()].push_back((int)(rng() % alphabet));
        }
      }
      FlatPaths fb, fc;
      Flatten(base, fb);
      Flatten(cand, fc);
      if (fb.n() != base.size() || fc.n() != cand.size()) { printf("FAIL: Flatten sizes\n"); return 1; }
      std::vector<int> match, ub, uc;
      Align(fb, fc, match);
      int last = -1;
      for (size_t y = 0; y < match.size(); y++) {
        if (match[y] < 0) continue;
        if (match[y] <= last || (size_t)match[y] >= base.size() || base[(size_t)match[y]] != cand[y]) {
          printf("FAIL: Align paired unequal walks or went backwards (trial %d step %d)\n", trial, step);
          return 1;
        }
        last = match[y];
        aligned_walks++;
      }
      total_walks += (long)cand.size();
      std::vector<int32_t> erased, added_nodes;
      std::vector<int64_t> added_off(1, 0);
      Difference(fb, fc, match, ub, uc, erased, added_nodes, added_off);
      // base - erased + added must be cand as multisets, and the difference must be minimal
      std::map<Walk, long> have, want;
      std::vector<char> gone(base.size(), 0);
      int prev = -1;
      for (int32_t i : erased) {
        if (i <= prev || i < 0 || (size_t)i >= base.size()) { printf("FAIL: erased indices not ascending / out of range\n"); return 1; }
        prev = i;
        gone[(size_t)i] = 1;
      }
      for (size_t i = 0; i < base.size(); i++)
        if (!gone[i]) have[base[i]]++;
      for (size_t a = 0; a + 1 < added_off.size(); a++)
        have[Walk(added_nodes.begin() + added_off[a], added_nodes.begin() + added_off[a + 1])]++;
      for (const Walk& w : cand) want[w]++;
      if (have != want) { printf("FAIL: base - erased + added != candidate (trial %d step %d)\n", trial, step); return 1; }
      std::map<Walk, long> cb, common;
      for (const Walk& w : base) cb[w]++;
      long inter = 0;
      for (auto& kv : want) {
        auto it = cb.find(kv.first);
        if (it != cb.end()) inter += std::min(it->second, kv.second);
      }
      if ((long)erased.size() != (long)base.size() - inter || (long)added_off.size() - 1 != (long)cand.size() - inter) {
        printf("FAIL: difference not minimal: erased %zu, added %zu, |base| %zu, |cand| %zu, common %ld\n", erased.size(), added_off.size() - 1,
               base.size(), cand.size(), inter);
        return 1;
      }
      checked++;
      if (rng() % 3 == 0) base = cand;   // the state moves along now and then
    }
  }
  printf("OK: %ld candidate lists, %.1f%% of their walks aligned with the current list\n", checked, 100.0 * (double)aligned_walks / (double)total_walks);
  return aligned_walks * 10 > total_walks * 9 ? 0 : 1;
}
