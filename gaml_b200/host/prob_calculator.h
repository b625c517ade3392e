// C++ host mirror of the reference's likelihood interface over the gaml_b200 C ABI.
//
// Same names, constructor shapes, argument meaning and statefulness as the reference's
// prob_calculator.h:7-124 and the parts of graph.h the likelihood path touches (Aligment graph.h:211-231,
// ReadSet graph.h:344-442, PacbioReadSet graph.h:444-593, Graph graph.h:233-273), so a caller written
// against the reference (gaml.cc:105,284; the 17 CalcProb call sites in moves.cc) reads the same here.
// Bodies are NOT the reference's: every score is computed by the CUDA library; this header only mirrors
// cache inserts to the device and forwards CalcProb. There is no CPU scoring path: construction throws
// std::runtime_error when the CUDA context cannot be created.
//
// In a build of the real GAML the reference's own graph.h types stay and INTEGRATION.md's adapter is used
// instead; this header is the self-contained form (the GPU box has no reference sources).
#pragma once
#include <climits>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "gaml_b200.h"

namespace gaml_b200 {

struct VecHash {   // graph.h:21-45
  size_t operator()(const std::vector<int>& v) const {
    size_t seed = 0;
    for (size_t i = 0; i < v.size(); i++) seed ^= std::hash<int>()(v[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};

struct Aligment {   // graph.h:211-231 (the reference's spelling)
  int position, edit_dist, read_id, orientation;
  Aligment() {}
  Aligment(int pos, int ed, int rid, int ori) : position(pos), edit_dist(ed), read_id(rid), orientation(ori) {}
  bool operator<(const Aligment& b) const { return position == b.position ? read_id < b.read_id : position < b.position; }
};
static_assert(sizeof(Aligment) == sizeof(gaml_alignment), "Aligment is the ABI record");

struct PacbioAligment {   // graph.h:516-535
  int position, position_end, read_id, pad;
  double logprob;
};
static_assert(sizeof(PacbioAligment) == sizeof(gaml_pacbio_alignment), "PacbioAligment is the ABI record");

// Only what the likelihood needs from Graph/Node: sequence length per node id and normalize_map.
struct Graph {
  std::vector<int> node_len;        // gr.nodes[i]->s.length()
  std::vector<int> normalize_map;   // graph.h:247-266 (empty = identity)
};

template <class Rec>
class CachedReadSet {
 public:
  CachedReadSet(const std::string& name, const std::string& filename, double match_prob, double mismatch_prob)
      : match_prob_(match_prob), mismatch_prob_(mismatch_prob), name_(name), filename_(filename) {}
  int GetNumberOfReads() const { return (int)read_lens_.size(); }
  int GetReadLen(int read_id) const { return read_lens_[read_id]; }
  const std::string& GetName() const { return name_; }
  void SetReadLens(const std::vector<int>& lens) { read_lens_ = lens; }
  // aligment_cache_[key] = records (graph.h:427 / 587); mirrored to the device at the next CalcProb.
  void InsertAligments(const std::vector<int>& key, const std::vector<Rec>& records) {
    aligment_cache_[key] = records;
    pending_.push_back(key);
  }
  bool HasAligments(const std::vector<int>& key) const { return aligment_cache_.count(key) != 0; }
  double match_prob_, mismatch_prob_;

 private:
  friend class ProbCalculator;
  std::string name_, filename_;
  std::vector<int> read_lens_;
  std::unordered_map<std::vector<int>, std::vector<Rec>, VecHash> aligment_cache_;
  std::vector<std::vector<int>> pending_;
};
using ReadSet = CachedReadSet<Aligment>;
using PacbioReadSet = CachedReadSet<PacbioAligment>;

struct SingleReadConfig {   // prob_calculator.h:7-18
  SingleReadConfig() {}
  SingleReadConfig(double pc, double s, double mp, double mps, double w, bool a)
      : penalty_constant(pc), step(s), min_prob_per_base(mp), min_prob_start(mps), weight(w), advice(a) {}
  double penalty_constant = 0, step = 50, min_prob_per_base = -0.7, min_prob_start = -10, weight = 1;
  bool advice = false;
};

struct PairedReadConfig {   // prob_calculator.h:20-35
  PairedReadConfig() {}
  PairedReadConfig(double pc, double s, double im, double is, double mp, double mps, double w, bool a)
      : penalty_constant(pc), step(s), insert_mean(im), insert_std(is), min_prob_per_base(mp), min_prob_start(mps),
        weight(w), advice(a) {}
  double penalty_constant = 0, step = 0, insert_mean = 0, insert_std = 1, min_prob_per_base = -0.7, min_prob_start = -10,
         weight = 1;
  bool advice = false;
};

class ProbCalculator {   // prob_calculator.h:37-124
 public:
  ProbCalculator(const std::vector<std::pair<SingleReadConfig, ReadSet*>>& single_reads,
                 const std::vector<std::pair<PairedReadConfig, std::pair<ReadSet*, ReadSet*>>>& paired_reads,
                 const std::vector<std::pair<SingleReadConfig, PacbioReadSet*>>& pacbio_reads, Graph& gr, int device = 0)
      : single_reads(single_reads), paired_reads(paired_reads), pacbio_reads(pacbio_reads), gr(gr) {
    Check(gaml_ctx_create(device, &ctx_), nullptr);
    Check(gaml_set_graph(ctx_, (int)gr.node_len.size(), gr.node_len.data(),
                         gr.normalize_map.empty() ? nullptr : gr.normalize_map.data()),
          ctx_);
    // sets are added in the order CalcProb reports them: single, paired, pacbio (prob_calculator.h:70-107)
    for (auto& e : this->single_reads) {
      gaml_readset_config c = Config(GAML_KIND_SINGLE, e.first.penalty_constant, e.first.step, e.first.min_prob_per_base,
                                     e.first.min_prob_start, e.first.weight, *e.second);
      AddSet(c, e.second, nullptr);
    }
    for (auto& e : this->paired_reads) {
      gaml_readset_config c = Config(GAML_KIND_PAIRED, e.first.penalty_constant, e.first.step, e.first.min_prob_per_base,
                                     e.first.min_prob_start, e.first.weight, *e.second.first);
      c.insert_mean = e.first.insert_mean;
      c.insert_std = e.first.insert_std;
      AddSet(c, e.second.first, e.second.second);
    }
    for (auto& e : this->pacbio_reads) {
      gaml_readset_config c = Config(GAML_KIND_PACBIO, e.first.penalty_constant, e.first.step, e.first.min_prob_per_base,
                                     e.first.min_prob_start, e.first.weight, *e.second);
      const int64_t n = e.second->GetNumberOfReads();
      int id = gaml_add_readset(ctx_, &c, n, 0, n, e.second->read_lens_.data(), nullptr, -1, -1);
      Check(id, ctx_);
      pacbio_ids_.push_back(id);
    }
  }
  ~ProbCalculator() { gaml_ctx_destroy(ctx_); }
  ProbCalculator(const ProbCalculator&) = delete;
  ProbCalculator& operator=(const ProbCalculator&) = delete;

  double CalcProb(std::vector<std::vector<int>>& paths, std::vector<std::pair<int, int>>& zeros, int& total_len) {
    SyncCaches();
    std::vector<int32_t> nodes;
    std::vector<int64_t> offs(1, 0);
    for (auto& p : paths) {
      nodes.insert(nodes.end(), p.begin(), p.end());
      offs.push_back((int64_t)nodes.size());
    }
    if (nodes.empty()) nodes.push_back(0);
    const size_t n_sets = single_ids_.size() + paired_ids_.size() + pacbio_ids_.size();
    std::vector<int32_t> z(2 * (n_sets ? n_sets : 1));
    gaml_result res;
    Check(gaml_calc_prob(ctx_, nodes.data(), offs.data(), (int)paths.size(), &res, z.data()), ctx_);
    zeros.clear();
    for (size_t s = 0; s < n_sets; s++) zeros.push_back(std::make_pair(z[2 * s], z[2 * s + 1]));
    total_len = res.total_len;
    return res.prob;
  }
  double CalcProb(std::vector<std::vector<int>>& paths, int& total_len) {
    std::vector<std::pair<int, int>> zeros;
    return CalcProb(paths, zeros, total_len);
  }
  double CalcProb(std::vector<std::vector<int>>& paths) {
    int tl;
    return CalcProb(paths, tl);
  }

  std::vector<std::pair<SingleReadConfig, ReadSet*>> single_reads;
  std::vector<std::pair<PairedReadConfig, std::pair<ReadSet*, ReadSet*>>> paired_reads;
  std::vector<std::pair<SingleReadConfig, PacbioReadSet*>> pacbio_reads;
  Graph& gr;
  gaml_ctx* context() { return ctx_; }

 private:
  static gaml_readset_config Config(int kind, double pc, double step, double mppb, double mps, double w, const ReadSet& rs) {
    gaml_readset_config c{};
    c.kind = kind;
    c.mismatch_prob = rs.mismatch_prob_;
    c.match_prob = rs.match_prob_;
    c.insert_std = 1;
    c.min_prob_per_base = mppb;
    c.min_prob_start = mps;
    c.weight = w;
    c.penalty_constant = pc;
    c.step = step;
    return c;
  }
  static gaml_readset_config Config(int kind, double pc, double step, double mppb, double mps, double w,
                                    const PacbioReadSet& rs) {
    gaml_readset_config c{};
    c.kind = kind;
    c.mismatch_prob = rs.mismatch_prob_;
    c.match_prob = rs.match_prob_;
    c.insert_std = 1;
    c.min_prob_per_base = mppb;
    c.min_prob_start = mps;
    c.weight = w;
    c.penalty_constant = pc;
    c.step = step;
    return c;
  }
  void AddSet(const gaml_readset_config& c, ReadSet* a, ReadSet* b) {
    const int64_t n = a->GetNumberOfReads();
    if (b && b->GetNumberOfReads() != n) throw std::runtime_error("paired read sets differ in size (graph.cc:1962)");
    int id = gaml_add_readset(ctx_, &c, n, 0, n, a->read_lens_.data(), b ? b->read_lens_.data() : nullptr, -1, -1);
    Check(id, ctx_);
    (b ? paired_ids_ : single_ids_).push_back(id);
  }
  template <class RS>
  void SyncOne(int set, int mate, RS* rs) {
    for (auto& key : rs->pending_) {
      auto& recs = rs->aligment_cache_[key];
      Insert(set, mate, key, recs);
    }
    rs->pending_.clear();
  }
  void Insert(int set, int mate, const std::vector<int>& key, const std::vector<Aligment>& recs) {
    Check(gaml_cache_insert(ctx_, set, mate, key.data(), (int)key.size(), reinterpret_cast<const gaml_alignment*>(recs.data()),
                            (int64_t)recs.size(), INT32_MIN),
          ctx_);
  }
  void Insert(int set, int, const std::vector<int>& key, const std::vector<PacbioAligment>& recs) {
    Check(gaml_cache_insert_pacbio(ctx_, set, key.data(), (int)key.size(),
                                   reinterpret_cast<const gaml_pacbio_alignment*>(recs.data()), (int64_t)recs.size()),
          ctx_);
  }
  void SyncCaches() {
    for (size_t i = 0; i < single_reads.size(); i++) SyncOne(single_ids_[i], 0, single_reads[i].second);
    for (size_t i = 0; i < paired_reads.size(); i++) {
      SyncOne(paired_ids_[i], 0, paired_reads[i].second.first);
      SyncOne(paired_ids_[i], 1, paired_reads[i].second.second);
    }
    for (size_t i = 0; i < pacbio_reads.size(); i++) SyncOne(pacbio_ids_[i], 0, pacbio_reads[i].second);
  }
  static void Check(int rc, gaml_ctx* ctx) {
    if (rc < 0) throw std::runtime_error(std::string("gaml_b200: ") + gaml_last_error(ctx));
  }
  gaml_ctx* ctx_ = nullptr;
  std::vector<int> single_ids_, paired_ids_, pacbio_ids_;
};

}  // namespace gaml_b200
