// TEST INFRASTRUCTURE ONLY — the oracle. Never imported, linked or executed by the product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs run it.
//
// A scalar CPU restatement of the GAML assembly-likelihood hot path, written from the behaviour of
// the reference (usamec/GAML) and pinned against the reference's own compiled code (oracle/_ref,
// see tests/test_oracle.py and tests/golden/). Each routine cites the reference lines it
// restates. Same CLI and file formats as oracle/ref_harness.cc:
//
//   gaml_oracle <workload GAMLWL1> <results GAMLRS1> [dump=0|1] [repeat=1]
//
// Arithmetic notes (they are what makes the paired state bit-reproducible):
//   * every per-alignment probability is mismatch^e * match^(len-e) from pow tables built exactly
//     like ReadSet::CalcMaxReadLen (graph.cc:1443-1454);
//   * a pair term is (p1*p2)*ins in that order (graph.cc:1889), ins from the table
//     [0, (int)(mean+5*std)) or the closed form (graph.cc:1593-1598, 1801-1804, 1877-1882);
//   * ScoringState::probs is updated by subtracting every erased-walk term and then adding every
//     added-walk term, one at a time, in enumeration order (graph.cc:1936-1950); erased walks come
//     out of an unordered_multiset keyed by the reference's vector hash (graph.h:21-45,
//     graph.cc:1745-1764), so the same container and hash are used here to get the same order.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

using Walk = std::vector<int>;
const int kWindowLen = 300;  // kMinSubpathLength, graph.cc:27

// graph.h:21-45 — the hash that decides the iteration order of GetChanges' multiset.
struct WalkHash {
  size_t operator()(const Walk& v) const {
    size_t seed = 0;
    for (size_t i = 0; i < v.size(); i++) seed ^= std::hash<int>()(v[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};

struct ShortRec { int pos, ed, read, orient; };                 // graph.h:211-215
struct LongRec { int pos, pos_end, read, pad; double logprob; };  // graph.h:516-520

template <class R>
using Cache = std::unordered_map<Walk, std::vector<R>, WalkHash>;

struct ReadSetData {
  int kind = 0, n_reads = 0, n_mates = 1;
  double mismatch = 0, match = 0, ins_mean = 0, ins_std = 1, mppb = 0, mps = 0, weight = 1, penalty = 0, step = 0;
  std::vector<int> len[2];
  std::vector<double> pow_match[2], pow_mismatch[2];
  Cache<ShortRec> cache[2];
  Cache<LongRec> lcache;
  int max_len[2] = {0, 0};
  // paired state (graph.h:612-619)
  std::vector<Walk> old_walks;
  std::vector<double> probs;
  int bad_bases = 0;
  // last per-read values (for dumps)
  std::vector<double> last;
};

struct Graph {
  std::vector<int> node_len, nmap;
};

// ---- helpers -------------------------------------------------------------------------------
void BuildPowTables(ReadSetData& rs, int m) {  // graph.cc:1443-1454
  int mx = 0;
  for (int v : rs.len[m]) mx = std::max(mx, v);
  rs.max_len[m] = mx;
  rs.pow_match[m].resize(mx + 7);
  rs.pow_mismatch[m].resize(mx + 7);
  for (size_t i = 0; i < rs.pow_match[m].size(); i++) {
    rs.pow_match[m][i] = pow(rs.match, (double)i);
    rs.pow_mismatch[m][i] = pow(rs.mismatch, (double)i);
  }
}

double InsertPdf(double d, double mean, double sd) {  // graph.cc:1593-1598
  double z = (d - mean) / sd;
  double e = exp(-z * z / 2.0);
  double c = sqrt(2 * M_PI) * sd;
  return e / c;
}

int WalkLength(const Graph& g, const Walk& w) {  // graph.cc:1766-1773
  int t = 0;
  for (int x : w) t += x < 0 ? -x : g.node_len[x];
  return t;
}

std::vector<Walk> SplitAtGaps(const Walk& w, std::vector<int>* gaps) {  // graph.cc:1666-1676, 1813-1824
  std::vector<Walk> ctgs(1);
  for (int x : w) {
    if (x < 0) {
      if (gaps) gaps->push_back(-x);
      ctgs.emplace_back();
    } else {
      ctgs.back().push_back(x);
    }
  }
  return ctgs;
}

Walk WindowKey(const Graph& g, const Walk& ctg, size_t i) {  // graph.cc:552-561, 618-627
  Walk key(1, ctg[i]);
  int beyond = 0;
  for (size_t j = i + 1; j < ctg.size(); j++) {
    beyond += g.node_len[ctg[j]];
    key.push_back(ctg[j]);
    if (beyond > kWindowLen) break;
  }
  return key;
}

// ---- final per-set reduction (graph.cc:1495-1537) ------------------------------------------
double MeanLogWithFloor(const std::vector<double>& p, int total_len, const ReadSetData& rs, bool paired,
                        int* floored) {
  if (total_len == 0) total_len = 1;
  *floored = 0;
  double acc = 0;
  int cnt = 0;
  for (size_t i = 0; i < p.size(); i++) {
    double v = p[i] / (2 * total_len);
    int l = rs.len[0][i] + (paired ? rs.len[1][i] : 0);
    double thr = exp(rs.mps + rs.mppb * l);
    if (v < thr) {
      (*floored)++;
      v = thr;
    }
    acc += log(v);
    cnt++;
  }
  return acc / cnt;
}

// ---- single reads (graph.cc:1650-1743 with AddPositions 600-649) ---------------------------
struct Placed { int pos, ed, orient; };

double ScoreSingle(const Graph& g, ReadSetData& rs, const std::vector<Walk>& walks, int* floored, int* total_len) {
  std::vector<std::vector<Placed>> at(rs.n_reads);
  int tl = 0;
  int stride = 0;  // +1000000 per walk, graph.cc:1685 (int arithmetic, wraps like the reference)
  for (const Walk& w : walks) {
    std::vector<int> gaps;
    std::vector<Walk> ctgs = SplitAtGaps(w, &gaps);
    for (size_t c = 0; c < ctgs.size(); c++) {
      if (c > 0) tl += gaps[c - 1];
      int cur = (int)((unsigned)stride + (unsigned)tl);
      for (size_t i = 0; i < ctgs[c].size(); i++) {
        tl += g.node_len[ctgs[c][i]];
        auto it = rs.cache[0].find(WindowKey(g, ctgs[c], i));
        if (it != rs.cache[0].end()) {
          for (const ShortRec& r : it->second) {
            int gp = (int)((unsigned)r.pos + (unsigned)cur);
            std::vector<Placed>& lst = at[r.read];
            bool dup = false;
            for (Placed& q : lst) {
              if (q.pos == gp) {  // same position: later record replaces the payload, graph.cc:635-641
                q.ed = r.ed;
                q.orient = r.orient;
                dup = true;
                break;
              }
            }
            if (!dup) lst.push_back(Placed{gp, r.ed, r.orient});
          }
        }
        cur = (int)((unsigned)cur + (unsigned)g.node_len[ctgs[c][i]]);
      }
    }
    stride = (int)((unsigned)stride + 1000000u);
  }
  std::vector<double>& p = rs.last;
  p.assign(rs.n_reads, 0.0);
  for (int r = 0; r < rs.n_reads; r++)
    for (const Placed& q : at[r]) p[r] += rs.pow_mismatch[0][q.ed] * rs.pow_match[0][rs.len[0][r] - q.ed];
  *total_len = tl;
  // Coverage penalty of single reads: the reference's sweep (graph.cc:1710-1733) only counts a gap when
  // last_event_type >= 3, and that variable is only ever assigned 1 or a negative type — bad_bases is identically 0.
  return MeanLogWithFloor(p, tl, rs, false, floored) - 0 * rs.penalty;
}

// ---- paired reads, incremental (graph.cc:1745-1989) ----------------------------------------
void GatherWalkMate(const Graph& g, const ReadSetData& rs, int m, const Walk& ctg, int start,
                    std::unordered_map<int, std::vector<Placed>>& at) {  // graph.cc:535-598
  int cur = start, seen_max = 0;
  for (size_t i = 0; i < ctg.size(); i++) {
    int node_max = 0;
    std::vector<Walk> keys(1, WindowKey(g, ctg, i));
    if (g.node_len[ctg[i]] > kWindowLen) keys.push_back(Walk(1, ctg[i]));
    for (const Walk& k : keys) {
      auto it = rs.cache[m].find(k);
      if (it == rs.cache[m].end()) continue;
      for (const ShortRec& r : it->second) {
        int gp = r.pos + cur;
        if (gp < seen_max - 5) continue;  // graph.cc:577
        node_max = std::max(node_max, gp);
        std::vector<Placed>& lst = at[r.read];
        bool dup = false;
        for (Placed& q : lst) {
          if (q.pos == gp) {
            q.ed = r.ed;
            q.orient = r.orient;
            dup = true;
            break;
          }
        }
        if (!dup) lst.push_back(Placed{gp, r.ed, r.orient});
      }
    }
    cur += g.node_len[ctg[i]];
    seen_max = std::max(seen_max, node_max);
  }
}

// Returns the walk's bad_bases (coverage-gap sweep, graph.cc:1893-1919); use_all_to_cov is always true on
// this path (prob_calculator.h:93).
int WalkPairTerms(const Graph& g, const ReadSetData& rs, const Walk& w, const std::vector<double>& ins_tab,
                  std::vector<std::pair<int, double>>& terms) {  // graph.cc:1794-1892
  std::vector<int> gaps;
  std::vector<Walk> ctgs = SplitAtGaps(w, &gaps);
  std::unordered_map<int, std::vector<Placed>> at1, at2;
  std::vector<std::pair<int, int>> events(1, std::make_pair(0, 1));   // (position, type): 1 = contig start, 3 = pair
  int cur = 0;
  for (size_t c = 0; c < ctgs.size(); c++) {
    if (c > 0) {
      cur += gaps[c - 1];
      events.push_back(std::make_pair(cur, 1));   // graph.cc:1835
    }
    GatherWalkMate(g, rs, 0, ctgs[c], cur, at1);
    GatherWalkMate(g, rs, 1, ctgs[c], cur, at2);
    cur += WalkLength(g, ctgs[c]);
  }
  for (auto& e : at1) {
    auto f = at2.find(e.first);
    if (f == at2.end()) continue;
    int r = e.first;
    int l1 = rs.len[0][r], l2 = rs.len[1][r];
    double cov_thr = exp(rs.mps + rs.mppb * (l2 + l2));   // mate-2 length twice, graph.cc:1855-1857
    for (const Placed& x : e.second) {
      double p1 = rs.pow_mismatch[0][x.ed] * rs.pow_match[0][l1 - x.ed];
      for (const Placed& y : f->second) {
        double p2 = rs.pow_mismatch[1][y.ed] * rs.pow_match[1][l2 - y.ed];
        if (x.orient == y.orient) continue;
        int d;
        if (x.pos < y.pos) {
          if (x.orient != 0 || y.orient != 1) continue;
          d = y.pos - x.pos + l2;
        } else {
          if (x.orient != 1 || y.orient != 0) continue;
          d = x.pos - y.pos + l1;
        }
        double ins = ((size_t)d < ins_tab.size()) ? ins_tab[d] : InsertPdf(d, rs.ins_mean, rs.ins_std);
        if (p1 * p2 * ins > cov_thr) {   // graph.cc:1883-1888
          events.push_back(std::make_pair(std::max(x.pos, y.pos), 3));
          events.push_back(std::make_pair(std::min(x.pos, y.pos), 3));
        }
        terms.push_back(std::make_pair(r, p1 * p2 * ins));
      }
    }
  }
  std::sort(events.begin(), events.end());
  int bad = 0, last_pos = 0, last_type = -1, last_begin = 0;
  for (auto& ev : events) {   // graph.cc:1901-1919; exp_cov_move is the set's `step`
    if (ev.second == 3 && ev.first - last_pos > rs.step && (last_type == 3 || last_type < 0) &&
        ev.first - last_begin > rs.ins_mean + 5 * rs.ins_std)
      bad += ev.first - last_pos;
    if (ev.second == 1) last_begin = ev.first;
    last_pos = ev.first;
    last_type = ev.second;
  }
  return bad;
}

double ScorePaired(const Graph& g, ReadSetData& rs, const std::vector<Walk>& walks, int* floored, int* total_len) {
  // GetChanges, graph.cc:1745-1764
  std::unordered_multiset<Walk, WalkHash> idx(rs.old_walks.begin(), rs.old_walks.end());
  std::vector<Walk> added, erased;
  for (const Walk& w : walks) {
    auto f = idx.find(w);
    if (f == idx.end()) added.push_back(w);
    else idx.erase(f);
  }
  erased.insert(erased.end(), idx.begin(), idx.end());
  if (rs.probs.empty()) rs.probs.resize(rs.n_reads);
  int tl = 0;
  for (const Walk& w : walks) tl += WalkLength(g, w);
  std::vector<double> ins_tab((int)(rs.ins_mean + 5 * rs.ins_std));  // graph.cc:1801-1804
  for (size_t i = 0; i < ins_tab.size(); i++) ins_tab[i] = InsertPdf((double)i, rs.ins_mean, rs.ins_std);
  std::vector<std::pair<int, double>> minus, plus;
  for (const Walk& w : erased) rs.bad_bases -= WalkPairTerms(g, rs, w, ins_tab, minus);   // graph.cc:1938
  for (const Walk& w : added) rs.bad_bases += WalkPairTerms(g, rs, w, ins_tab, plus);     // graph.cc:1946
  for (auto& t : minus) rs.probs[t.first] -= t.second;  // graph.cc:1936-1942
  for (auto& t : plus) rs.probs[t.first] += t.second;   // graph.cc:1944-1950
  rs.old_walks = walks;                                 // graph.cc:1986
  rs.last = rs.probs;
  *total_len = tl;
  return MeanLogWithFloor(rs.probs, tl, rs, true, floored) - rs.bad_bases * rs.penalty;   // graph.cc:1988
}

// ---- PacBio (graph.cc:3171-3261, 2410-2503, 3052-3088; logdouble.hpp:21-31) ----------------
inline void LseAdd(double& acc, double v) {
  if (std::isinf(acc) && acc < 0) { acc = v; return; }
  if (std::isinf(v) && v < 0) return;
  double hi = std::max(acc, v), lo = std::min(acc, v);
  acc = hi + log1p(exp(lo - hi));
}

double ScorePacbio(const Graph& g, ReadSetData& rs, const std::vector<Walk>& walks, int* floored, int* total_len) {
  std::vector<double>& lp = rs.last;
  lp.assign(rs.n_reads, -std::numeric_limits<double>::infinity());
  int tl = 0;
  int bad_bases = 0;
  // GetMinReadProb (graph.h:478-481): mismatch^(len/4) * match^(3 len/4) as a logdouble
  const double lmm = log(rs.mismatch), lm = log(rs.match);
  for (Walk w : walks) {
    for (int& x : w)
      if (x >= 0) x = g.nmap[x];  // graph.h:268-273
    size_t n = w.size();
    std::vector<int> begin(n), end(n);
    int off = 0;
    // coverage events of this walk (graph.cc:3197-3223): an artificial interval, one interval per node, ...
    std::vector<std::pair<int, int>> events;
    events.push_back({-1000, 1});
    events.push_back({2000, -3000});
    for (size_t i = 0; i < n; i++) {
      begin[i] = off;
      if (w[i] >= 0) {
        int cl = g.node_len[w[i]];
        events.push_back({off, 1});
        events.push_back({off + cl, -cl});
      }
      off += w[i] < 0 ? -w[i] : g.node_len[w[i]];
      end[i] = off;
    }
    tl += off;
    for (size_t i = 0; i < n; i++) {
      Walk key;
      for (size_t j = i; j < n; j++) {
        key.push_back(w[j]);
        auto it = rs.lcache.find(key);
        if (it != rs.lcache.end())
          for (const LongRec& r : it->second) {
            LseAdd(lp[r.read], r.logprob);  // no de-dup, graph.cc:2487-2500
            // ... and one per alignment at least as probable as the read's minimum (graph.cc:3213-3221)
            const double min_lp = lmm * (rs.len[0][r.read] * 0.25) + lm * (rs.len[0][r.read] * 0.75);
            if (r.logprob < min_lp) continue;
            events.push_back({begin[i] + r.pos, 1});
            events.push_back({begin[i] + r.pos_end, (begin[i] + r.pos) - (begin[i] + r.pos_end)});
          }
        if ((end[j] - begin[i]) - (end[i] - begin[i]) > rs.max_len[0]) break;  // graph.cc:2450
      }
    }
    // the sweep (graph.cc:3225-3250): stretches that begin further than exp_cov_move after the earliest open interval
    std::sort(events.begin(), events.end());
    std::multiset<int> inters;
    const int wl = off;
    for (size_t j = 0; j < events.size(); j++) {
      if (events[j].second == 1) inters.insert(events[j].first);
      else inters.erase(inters.find(events[j].first + events[j].second));
      int good_start = wl - 250;
      if (!inters.empty()) good_start = (int)(*inters.begin() + rs.step);
      if (j + 1 < events.size()) good_start = std::min(events[j + 1].first, good_start);
      good_start = std::min(good_start, wl - 250);
      if (good_start > std::max(2500, events[j].first)) bad_bases += good_start - std::max(2500, events[j].first);
    }
  }
  *total_len = tl;
  int den = tl == 0 ? 1 : tl;
  *floored = 0;
  double acc = log(1.0);  // logdouble total_prob = 1, graph.cc:3066
  int cnt = 0;
  double a = log(exp(rs.mps)), b = log(exp(rs.mppb));  // graph.cc:3075-3076
  for (int r = 0; r < rs.n_reads; r++) {
    double v = lp[r];
    double fl = a + b * (double)rs.len[0][r];
    if (v < fl) {
      (*floored)++;
      v = fl;
    }
    acc += v;
    cnt++;
  }
  return (acc / cnt - log((double)(2 * den))) - bad_bases * rs.penalty;   // graph.cc:3259-3260
}

// ---- ProbCalculator::CalcProb (prob_calculator.h:63-109) -----------------------------------
struct Calculator {
  Graph g;
  std::vector<ReadSetData> sets;
  void Reset() {
    for (auto& s : sets) {
      s.old_walks.clear();
      s.probs.clear();
      s.bad_bases = 0;
    }
  }
  double CalcProb(const std::vector<Walk>& walks, std::vector<std::pair<int, int>>& zeros, int& total_len) {
    zeros.assign(sets.size(), std::make_pair(0, 0));
    double prob = 0;
    for (int kind = 0; kind < 3; kind++) {  // single sets, then paired, then pacbio
      for (size_t s = 0; s < sets.size(); s++) {
        ReadSetData& rs = sets[s];
        if (rs.kind != kind) continue;
        int z = 0;
        double sc = kind == 0   ? ScoreSingle(g, rs, walks, &z, &total_len)
                    : kind == 1 ? ScorePaired(g, rs, walks, &z, &total_len)
                                : ScorePacbio(g, rs, walks, &z, &total_len);
        prob += sc * rs.weight;
        zeros[s] = std::make_pair(z, rs.n_reads);
      }
    }
    return prob;
  }
};

// ---- file IO -------------------------------------------------------------------------------
// ---- PacBio alignment probability (PacbioReadSet::AligmentProbability, graph.cc:2175-2297) -----------------------
// Forward DP in log space over the cells around an alignment's CIGAR path. The reference collects the cells as a list
// of (row, column) pairs and "uniquifies" it (graph.cc:2150-2173): per row, every column between the smallest and the
// largest one listed. Here a row's cells are kept as that [lo, hi] range from the start, which is the same set.
struct RowRanges {
  int first_row = 0;
  std::vector<int> lo, hi;   // per row from first_row; lo > hi = the row has no cell
  void Cover(int row, int c_lo, int c_hi) {
    if (lo.empty()) { first_row = row; lo.push_back(c_lo); hi.push_back(c_hi); return; }
    while (row < first_row) { lo.insert(lo.begin(), 1000000); hi.insert(hi.begin(), -1000000); first_row--; }
    while (row >= first_row + (int)lo.size()) { lo.push_back(1000000); hi.push_back(-1000000); }
    int i = row - first_row;
    lo[i] = std::min(lo[i], c_lo);
    hi[i] = std::max(hi[i], c_hi);
  }
};

double AlignmentLogProb(const std::string& s1, const std::string& s2, int posstart, const std::vector<std::pair<int, int>>& ops,
                        int band, double match, double mismatch) {
  std::string cigar;   // ExpandCigar, graph.cc:2127-2134
  for (auto& o : ops) cigar.append((size_t)o.first, (char)o.second);
  int bl = 0, el = 0;  // GetCigarEnds, graph.cc:2136-2149: leading / trailing insertion runs
  for (size_t i = 0; i < cigar.size(); i++)
    if (cigar[i] != 'I') { bl = (int)i; break; }
  for (int i = (int)cigar.size() - 1; i >= 0; i--)
    if (cigar[i] != 'I') { el = (int)cigar.size() - i; break; }
  bl = std::min(bl, 200);
  el = std::min(el, 200);
  RowRanges path;
  path.Cover(0, 0, 0);
  if (bl > 0)
    for (int i = -bl; i < 3; i++) path.Cover(i, 0, bl - 1);   // graph.cc:2188-2192
  int row = 0, col = 0;
  for (char c : cigar) {   // graph.cc:2193-2203
    if (c == 'M') { row++; col++; }
    else if (c == 'I') col++;
    else if (c == 'D') row++;
    path.Cover(row, col, col);
  }
  for (int i = row; i < row + el; i++) path.Cover(i, col - el, col);   // graph.cc:2204-2208
  RowRanges cells;   // every listed cell widened by the band in both directions, graph.cc:2210-2222
  for (size_t i = 0; i < path.lo.size(); i++) {
    if (path.lo[i] > path.hi[i]) continue;
    for (int d = -band; d <= band; d++) cells.Cover(path.first_row + (int)i + d, path.lo[i] - band, path.hi[i] + band);
  }
  const double lmatch = log(match), lmismatch = log(mismatch);
  auto match_lp = [&](char a, char b) {   // MatchProbability, graph.h:555-563 ('\n' separates contigs: probability 0)
    if (a == '\n' || b == '\n') return -std::numeric_limits<double>::infinity();
    return a != b ? lmismatch : lmatch;
  };
  const int n_rows = (int)cells.lo.size();
  std::vector<std::vector<double>> res(n_rows);
  for (int i = 0; i < n_rows; i++)
    if (cells.lo[i] <= cells.hi[i]) res[i].assign((size_t)(cells.hi[i] - cells.lo[i] + 1), -std::numeric_limits<double>::infinity());
  auto inside = [&](int r, int c) { return r >= 0 && c - cells.lo[r] >= 0 && c - cells.lo[r] < (int)res[r].size(); };
  double ret = log(0.0);   // logdouble ret = 0
  for (int i = 0; i < n_rows; i++)
    for (int c = cells.lo[i]; c <= cells.hi[i]; c++)
      if (c == 0) res[i][0 - cells.lo[i]] = 0.0;   // probability 1, graph.cc:2240-2244
  for (int i = 0; i < n_rows; i++) {
    const int r_abs = cells.first_row + i;
    for (int c = cells.lo[i]; c <= cells.hi[i]; c++) {
      if (c == 0) continue;
      if (c - 1 < 0 || c - 1 >= (int)s2.size()) continue;
      const int p1 = r_abs + posstart - 1;
      if (p1 < 0 || p1 >= (int)s1.size()) continue;
      double& cell = res[i][c - cells.lo[i]];
      if (inside(i - 1, c - 1)) LseAdd(cell, res[i - 1][c - 1 - cells.lo[i - 1]] + match_lp(s1[p1], s2[c - 1]));
      if (inside(i - 1, c)) LseAdd(cell, res[i - 1][c - cells.lo[i - 1]] + match_lp(s1[p1], '-'));
      if (inside(i, c - 1)) LseAdd(cell, res[i][c - 1 - cells.lo[i]] + match_lp('-', s2[c - 1]));
      if (c == (int)s2.size()) LseAdd(ret, cell);
    }
  }
  return ret;
}

struct Reader {
  std::vector<char> buf;
  size_t off = 8;
  explicit Reader(const char* path, const char* magic = "GAMLWL1\0") {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n);
    if (fread(buf.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
    if (n < 8 || memcmp(buf.data(), magic, 8) != 0) { fprintf(stderr, "bad magic\n"); exit(2); }
  }
  int i32() { int v; memcpy(&v, &buf[off], 4); off += 4; return v; }
  double f64() { double v; memcpy(&v, &buf[off], 8); off += 8; return v; }
  void ints(int* d, size_t n) { if (n) memcpy(d, &buf[off], 4 * n); off += 4 * n; }
};

}  // namespace

// --alnprob mode: same file formats as oracle/ref_harness.cc's
int AlignProbMode(const char* in_path, const char* out_path) {
  Reader rd(in_path, "GAMLAP1\0");
  const double match = rd.f64(), mismatch = rd.f64();
  const int band = rd.i32(), n = rd.i32();
  std::vector<double> out(n);
  double secs = 0;
  for (int a = 0; a < n; a++) {
    const int posstart = rd.i32();
    const int n1 = rd.i32();
    std::string s1(&rd.buf[rd.off], &rd.buf[rd.off] + n1);
    rd.off += n1;
    const int n2 = rd.i32();
    std::string s2(&rd.buf[rd.off], &rd.buf[rd.off] + n2);
    rd.off += n2;
    const int n_ops = rd.i32();
    std::vector<std::pair<int, int>> ops(n_ops);
    for (auto& o : ops) { o.first = rd.i32(); o.second = rd.i32(); }
    auto t0 = std::chrono::steady_clock::now();
    out[a] = AlignmentLogProb(s1, s2, posstart, ops, band, match, mismatch);
    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  FILE* f = fopen(out_path, "wb");
  if (!f) return 2;
  fwrite(out.data(), 8, out.size(), f);
  fwrite(&secs, 8, 1, f);
  fclose(f);
  fprintf(stderr, "gaml_oracle --alnprob: %d alignments, %.6f s\n", n, secs);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && strcmp(argv[1], "--alnprob") == 0) return AlignProbMode(argv[2], argv[3]);
  if (argc < 3) {
    fprintf(stderr, "usage: %s <workload> <results> [dump] [repeat]\n       %s --alnprob <alignments> <logvals>\n", argv[0], argv[0]);
    return 2;
  }
  bool dump = argc > 3 && atoi(argv[3]) != 0;
  int repeat = argc > 4 ? atoi(argv[4]) : 1;
  Reader rd(argv[1]);
  Calculator calc;
  int n_nodes = rd.i32();
  calc.g.node_len.resize(n_nodes);
  calc.g.nmap.resize(n_nodes);
  rd.ints(calc.g.node_len.data(), n_nodes);
  rd.ints(calc.g.nmap.data(), n_nodes);
  int n_sets = rd.i32();
  calc.sets.resize(n_sets);
  for (auto& rs : calc.sets) {
    rs.kind = rd.i32();
    rs.mismatch = rd.f64(); rs.match = rd.f64(); rs.ins_mean = rd.f64(); rs.ins_std = rd.f64();
    rs.mppb = rd.f64(); rs.mps = rd.f64(); rs.weight = rd.f64(); rs.penalty = rd.f64(); rs.step = rd.f64();
    rs.n_reads = rd.i32();
    rs.n_mates = rd.i32();
    for (int m = 0; m < rs.n_mates; m++) {
      rs.len[m].resize(rs.n_reads);
      rd.ints(rs.len[m].data(), rs.n_reads);
      BuildPowTables(rs, m);
    }
    for (int m = 0; m < rs.n_mates; m++) {
      int nk = rd.i32();
      for (int k = 0; k < nk; k++) {
        int kl = rd.i32();
        Walk key(kl);
        rd.ints(key.data(), kl);
        int nr = rd.i32();
        if (rs.kind == 2) {
          std::vector<LongRec>& v = rs.lcache[key];
          v.resize(nr);
          if (nr) memcpy(v.data(), &rd.buf[rd.off], sizeof(LongRec) * (size_t)nr);
          rd.off += sizeof(LongRec) * (size_t)nr;
        } else {
          std::vector<ShortRec>& v = rs.cache[m][key];
          v.resize(nr);
          rd.ints(reinterpret_cast<int*>(v.data()), 4 * (size_t)nr);
        }
      }
    }
  }
  int n_evals = rd.i32();
  std::vector<std::vector<Walk>> evals(n_evals);
  for (auto& ws : evals) {
    ws.resize(rd.i32());
    for (auto& w : ws) {
      w.resize(rd.i32());
      rd.ints(w.data(), w.size());
    }
  }
  FILE* out = fopen(argv[2], "wb");
  if (!out) return 2;
  fwrite("GAMLRS1\0", 1, 8, out);
  int hdr[3] = {n_evals * repeat, n_sets, dump ? 1 : 0};
  fwrite(hdr, 4, 3, out);
  double total = 0;
  for (int rep = 0; rep < repeat; rep++) {
    calc.Reset();
    for (auto& ws : evals) {
      std::vector<std::pair<int, int>> zeros;
      int tl = 0;
      auto t0 = std::chrono::steady_clock::now();
      double score = calc.CalcProb(ws, zeros, tl);
      double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      total += secs;
      fwrite(&score, 8, 1, out);
      int tli[2] = {tl, 0};
      fwrite(tli, 4, 2, out);
      fwrite(&secs, 8, 1, out);
      for (auto& z : zeros) {
        int zz[2] = {z.first, z.second};
        fwrite(zz, 4, 2, out);
      }
      if (dump) {
        for (auto& rs : calc.sets) {
          int n[2] = {(int)rs.last.size(), 0};
          fwrite(n, 4, 2, out);
          fwrite(rs.last.data(), 8, rs.last.size(), out);
        }
      }
    }
  }
  fclose(out);
  fprintf(stderr, "gaml_oracle: %d evals x %d, %.6f s scoring\n", n_evals, repeat, total);
  return 0;
}
