// Drop-in replacement for the reference's prob_calculator.h (usamec/GAML): same class, same constructor,
// same three CalcProb overloads, same public members — the bodies forward to the gaml_b200 CUDA library
// through its C ABI (include/gaml_b200.h). gaml.cc (Optimize, main) and moves.cc compile against this file
// UNCHANGED; graph.h / graph.cc stay the reference's (the ReadSet objects remain the host-side cache owners
// that the moves use, moves.cc:831-856, 948-964, and their aligners keep filling the caches).
//
// Build (see oracle/build_ref.sh, target gaml_gpu): put this directory BEFORE the reference on the include
// path and compile gaml.cc / moves.cc with -Dprivate=public so the adapter can read ReadSet::aligment_cache_
// (graph.h:427) — access specifiers do not change object layout, graph.o is the reference's own object file.
// A maintainer would instead add two accessors to ReadSet; INTEGRATION.md shows that patch.
//
// What stays on the host, exactly as the reference does it, is everything that FILLS the cache: the adapter
// calls the same precompute routines in the same places (graph.cc:1967-1968, 538-542, 605-609), so the same
// keys are aligned at the same time. It then mirrors keys it has not sent yet to the device and calls
// gaml_calc_prob. No score is ever computed on the CPU; if the CUDA context cannot be created the program
// aborts like the reference's asserts do.
#ifndef PROB_CALCULATOR_H__
#define PROB_CALCULATOR_H__

#include <cstdio>
#include <cstdlib>
#include <unordered_set>

#include "gaml_b200.h"
#include "graph.h"
#include "utility.h"

struct SingleReadConfig {   // reference prob_calculator.h:7-18
  SingleReadConfig() {}
  SingleReadConfig(double pc, double s, double mp, double mps, double w, bool a)
      : penalty_constant(pc), step(s), min_prob_per_base(mp), min_prob_start(mps), weight(w), advice(a) {}
  double penalty_constant;
  double step;
  double min_prob_per_base;
  double min_prob_start;
  double weight;
  bool advice;
};

struct PairedReadConfig {   // reference prob_calculator.h:20-35
  PairedReadConfig() {}
  PairedReadConfig(double pc, double s, double im, double is, double mp, double mps, double w, bool a)
      : penalty_constant(pc), step(s), insert_mean(im), insert_std(is), min_prob_per_base(mp), min_prob_start(mps),
        weight(w), advice(a) {}
  double penalty_constant;
  double step;
  double insert_mean;
  double insert_std;
  double min_prob_per_base;
  double min_prob_start;
  double weight;
  bool advice;
};

class ProbCalculator {
 public:
  ProbCalculator(const vector<pair<SingleReadConfig, ReadSet*>>& single_reads,
                 const vector<pair<PairedReadConfig, pair<ReadSet*, ReadSet*>>>& paired_reads,
                 const vector<pair<SingleReadConfig, PacbioReadSet*>>& pacbio_reads, Graph& gr)
      : single_reads(single_reads), paired_reads(paired_reads), pacbio_reads(pacbio_reads), gr(gr), ctx_(NULL) {
    paired_scoring_states.resize(paired_reads.size());   // kept for source compatibility; the state lives on the device
  }
  ~ProbCalculator() { gaml_ctx_destroy(ctx_); }

  double CalcProb(vector<vector<int>>& paths, vector<pair<int, int>>& zeros, int& total_len) {
    if (!ctx_) Init();   // lazy: gaml.cc constructs the calculator before PrepareReads fills the read sets
    FillAndMirrorCaches(paths);
    vector<int32_t> nodes;
    vector<int64_t> offs(1, 0);
    for (auto& p : paths) {
      nodes.insert(nodes.end(), p.begin(), p.end());
      offs.push_back((int64_t)nodes.size());
    }
    if (nodes.empty()) nodes.push_back(0);
    vector<int32_t> z(2 * (n_sets_ ? n_sets_ : 1));
    gaml_result res;
    Check(gaml_calc_prob(ctx_, nodes.data(), offs.data(), (int)paths.size(), &res, z.data()));
    zeros.clear();
    for (size_t s = 0; s < n_sets_; s++) zeros.push_back(make_pair(z[2 * s], z[2 * s + 1]));
    total_len = res.total_len;
    return res.prob;
  }
  double CalcProb(vector<vector<int>>& paths, int& total_len) {
    vector<pair<int, int>> zeros;
    return CalcProb(paths, zeros, total_len);
  }
  double CalcProb(vector<vector<int>>& paths) {
    int tl;
    return CalcProb(paths, tl);
  }

  vector<pair<SingleReadConfig, ReadSet*>> single_reads;
  vector<pair<PairedReadConfig, pair<ReadSet*, ReadSet*>>> paired_reads;
  vector<pair<SingleReadConfig, PacbioReadSet*>> pacbio_reads;
  vector<ScoringState> paired_scoring_states;
  Graph& gr;

 private:
  struct Mirror {   // one ReadSet <-> one (set, mate) store of the library
    ReadSet* rs;
    int set, mate;
    bool paired;
    unordered_set<vector<int>> sent;
  };

  void Check(int rc) {
    if (rc < 0) {
      fprintf(stderr, "gaml_b200: %s\n", gaml_last_error(ctx_));
      abort();
    }
  }

  static gaml_readset_config Cfg(int kind, double pc, double step, double mppb, double mps, double w, double match,
                                 double mismatch) {
    gaml_readset_config c;
    c.kind = kind;
    c.reserved = 0;
    c.mismatch_prob = mismatch;
    c.match_prob = match;
    c.insert_mean = 0;
    c.insert_std = 1;
    c.min_prob_per_base = mppb;
    c.min_prob_start = mps;
    c.weight = w;
    c.penalty_constant = pc;
    c.step = step;
    return c;
  }

  static vector<int32_t> Lens(const ReadSet& rs) {
    vector<int32_t> l(rs.GetNumberOfReads());
    for (size_t i = 0; i < l.size(); i++) l[i] = rs.GetReadLen((int)i);
    return l;
  }

  void Init() {
    Check(gaml_ctx_create(getenv("GAML_GPU_DEVICE") ? atoi(getenv("GAML_GPU_DEVICE")) : 0, &ctx_));
    vector<int32_t> node_len(gr.nodes.size());
    for (size_t i = 0; i < gr.nodes.size(); i++) node_len[i] = (int32_t)gr.nodes[i]->s.length();
    vector<int32_t> nmap(gr.normalize_map.begin(), gr.normalize_map.end());
    Check(gaml_set_graph(ctx_, (int)node_len.size(), node_len.data(), nmap.size() == node_len.size() ? nmap.data() : NULL));
    // CalcProb reports single sets, then paired, then pacbio (reference prob_calculator.h:70-107)
    for (auto& e : single_reads) {
      gaml_readset_config c = Cfg(GAML_KIND_SINGLE, e.first.penalty_constant, e.first.step, e.first.min_prob_per_base,
                                  e.first.min_prob_start, e.first.weight, e.second->match_prob_, e.second->mismatch_prob_);
      vector<int32_t> l = Lens(*e.second);
      int id = gaml_add_readset(ctx_, &c, (int64_t)l.size(), 0, (int64_t)l.size(), l.data(), NULL, -1, -1);
      Check(id);
      mirrors_.push_back(Mirror{e.second, id, 0, false, {}});
    }
    for (auto& e : paired_reads) {
      gaml_readset_config c = Cfg(GAML_KIND_PAIRED, e.first.penalty_constant, e.first.step, e.first.min_prob_per_base,
                                  e.first.min_prob_start, e.first.weight, e.second.first->match_prob_,
                                  e.second.first->mismatch_prob_);
      c.insert_mean = e.first.insert_mean;
      c.insert_std = e.first.insert_std;
      vector<int32_t> l1 = Lens(*e.second.first), l2 = Lens(*e.second.second);
      assert(l1.size() == l2.size());   // reference graph.cc:1962
      int id = gaml_add_readset(ctx_, &c, (int64_t)l1.size(), 0, (int64_t)l1.size(), l1.data(), l2.data(), -1, -1);
      Check(id);
      mirrors_.push_back(Mirror{e.second.first, id, 0, true, {}});
      mirrors_.push_back(Mirror{e.second.second, id, 1, true, {}});
    }
    if (!pacbio_reads.empty()) {
      fprintf(stderr, "gaml_b200 adapter: PacBio sets need a pre-filled cache (blasr drivers are not mirrored yet)\n");
      abort();
    }
    n_sets_ = single_reads.size() + paired_reads.size();
  }

  // Window key of node i inside a contig (reference graph.cc:552-561 / 618-627).
  vector<int> WindowKey(const vector<int>& ctg, size_t i) {
    vector<int> key(1, ctg[i]);
    int beyond = 0;
    for (size_t j = i + 1; j < ctg.size(); j++) {
      beyond += gr.nodes[ctg[j]]->s.length();
      key.push_back(ctg[j]);
      if (beyond > 300) break;   // kMinSubpathLength, graph.cc:27
    }
    return key;
  }

  void MirrorKey(Mirror& m, const vector<int>& key) {
    if (m.sent.count(key)) return;
    auto it = m.rs->aligment_cache_.find(key);
    if (it == m.rs->aligment_cache_.end()) return;   // never aligned: the reference treats it as an empty list
    m.sent.insert(key);
    static_assert(sizeof(Aligment) == sizeof(gaml_alignment), "Aligment is the ABI record");
    Check(gaml_cache_insert(ctx_, m.set, m.mate, key.data(), (int)key.size(),
                            reinterpret_cast<const gaml_alignment*>(it->second.data()), (int64_t)it->second.size(),
                            INT32_MIN));
  }

  void FillAndMirrorCaches(const vector<vector<int>>& paths) {
    // (1) run the reference's own cache-filling calls where CalcScoreForPathsNew / GetPositionsOnlyPath /
    //     AddPositions run them, so the same keys get aligned (internal min-hash aligner or bowtie2).
    for (auto& e : paired_reads) {
      e.second.first->PrecomputeAlignmentForPaths(paths, gr);    // graph.cc:1967
      e.second.second->PrecomputeAlignmentForPaths(paths, gr);   // graph.cc:1968
    }
    for (auto& m : mirrors_) {
      for (auto& path : paths) {
        vector<int> ctg;
        for (size_t i = 0; i <= path.size(); i++) {
          if (i == path.size() || path[i] < 0) {
            unordered_set<vector<int>> missing;
            m.rs->GetSubpathsFromPath(ctg, gr, missing);          // graph.cc:538-542, 605-609
            if (!missing.empty()) m.rs->PrecomputeAligmentForSubpaths(gr, USetToVector(missing));
            // (2) mirror every key this contig looks up and the device does not hold yet
            for (size_t k = 0; k < ctg.size(); k++) {
              MirrorKey(m, WindowKey(ctg, k));
              if (m.paired && gr.nodes[ctg[k]]->s.length() > 300) MirrorKey(m, vector<int>(1, ctg[k]));   // graph.cc:563-566
            }
            ctg.clear();
          } else {
            ctg.push_back(path[i]);
          }
        }
      }
    }
  }

  gaml_ctx* ctx_;
  size_t n_sets_ = 0;
  vector<Mirror> mirrors_;
};

#endif
