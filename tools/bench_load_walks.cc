// Host micro-benchmark of load_walks (walk_set.h) at BASELINE config 4's list shape: 10 000 walks, one move apart.
// g++ -O2 -std=c++17 -I gaml_b200/csrc tools/bench_load_walks.cc -o /tmp/blw && /tmp/blw
#include <chrono>
#include <cstdio>
#include <random>
#include "walk_set.h"
using namespace gaml;
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 10000;
  std::vector<std::vector<int>> base(n);
  for (int i = 0; i < n; i++) base[i] = {2 * i};
  std::mt19937 rng(5);
  auto flat = [](const std::vector<std::vector<int>>& w, std::vector<int32_t>& nodes, std::vector<int64_t>& offs) {
    nodes.clear(); offs.assign(1, 0);
    for (auto& x : w) { nodes.insert(nodes.end(), x.begin(), x.end()); offs.push_back((int64_t)nodes.size()); }
  };
  // a trajectory of joins and splits
  std::vector<std::vector<std::vector<int>>> lists(1, base);
  for (int k = 1; k < 64; k++) {
    auto cur = lists.back();
    if (rng() % 2 && cur.size() > 2) {
      int i = rng() % cur.size(), j = rng() % cur.size();
      if (i == j) j = (j + 1) % cur.size();
      cur[i].insert(cur[i].end(), cur[j].begin(), cur[j].end());
      cur.erase(cur.begin() + j);
    } else {
      int i = rng() % cur.size();
      if (cur[i].size() >= 2) { std::vector<int> tail(cur[i].begin() + 1, cur[i].end()); cur[i].resize(1); cur.push_back(tail); }
    }
    lists.push_back(cur);
  }
  std::vector<std::vector<int32_t>> fn(lists.size());
  std::vector<std::vector<int64_t>> fo(lists.size());
  for (size_t k = 0; k < lists.size(); k++) flat(lists[k], fn[k], fo[k]);
  WalkSet ws[2];
  WalkDiff d;
  int cur = 0;
  load_walks(ws[cur], nullptr, fn[0].data(), fo[0].data(), (int)lists[0].size(), d);
  cur ^= 1;
  std::vector<char> evict(8 << 20);
  double best = 1e9, sum = 0;
  int cnt = 0;
  for (int rep = 0; rep < 20; rep++)
    for (size_t k = 1; k < lists.size(); k++) {
      for (size_t i = 0; i < evict.size(); i += 64) evict[i]++;   // the host runs other code between evaluations
      auto t0 = std::chrono::steady_clock::now();
      load_walks(ws[cur], &ws[cur ^ 1], fn[k].data(), fo[k].data(), (int)lists[k].size(), d);
      double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      if (!d.valid) printf("invalid diff at %zu\n", k);
      cur ^= 1;
      best = std::min(best, us); sum += us; cnt++;
    }
  printf("n=%d load_walks: mean %.2f us, best %.2f us (changed %zu/%zu)\n", n, sum / cnt, best, d.old_changed.size(), d.new_changed.size());
}
