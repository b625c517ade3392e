// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// Cache-injection harness around the UNMODIFIED reference scoring code (usamec/GAML).
// It is compiled by oracle/build_ref.sh against /root/reference/{graph.cc,prob_calculator.h,...}
// where they lie; the binary lands in oracle/_ref/ (git-ignored). It
//   1. reads a GAMLWL1 workload (gaml_b200/workload.py),
//   2. fills the reference's own ReadSet / PacbioReadSet objects (graph.h:344-593) with the
//      workload's alignment caches (the private members listed in SURVEY.md §8c),
//   3. drives ONE reference ProbCalculator (prob_calculator.h:37-124) through every walk set in
//      order, exactly as gaml.cc:105,284 does, timing each CalcProb call,
//   4. writes a GAMLRS1 result file (score, total_len, floored counts, optional per-read values).
//
// usage: ref_harness <workload> <results> [dump=0|1] [repeat=1]
//        ref_harness --alnprob <alignments> <logvals>     (PacbioReadSet::AligmentProbability on injected alignments)
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <random>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define private public
#include "graph.h"
#include "prob_calculator.h"
#undef private

// Free functions with external linkage in the reference's graph.cc (not declared in graph.h).
void PositionsToReadProbs(int num_reads, const vector<vector<pair<int, pair<int, int> > > >& positions,
                          const ReadSet& read_set, vector<double>& read_probs);        // graph.cc:1482
void AddPositionsToReadProbsPacbio(const vector<vector<pair<pair<int, int>, logdouble> > >& positions,
                                   vector<logdouble>& read_probs);                      // graph.cc:3052

string gBowtiePath, gBlasrPath;  // defined in gaml.cc in the reference; graph.cc only needs the symbols

namespace {

struct Reader {
  std::vector<char> buf;
  size_t off = 0;
  explicit Reader(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n);
    if (fread(buf.data(), 1, n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
    fclose(f);
  }
  int i32() { int v; memcpy(&v, &buf[off], 4); off += 4; return v; }
  double f64() { double v; memcpy(&v, &buf[off], 8); off += 8; return v; }
  void ints(int* dst, size_t n) { memcpy(dst, &buf[off], 4 * n); off += 4 * n; }
};

struct SetInfo {
  int kind;
  ReadSet* rs1 = nullptr;
  ReadSet* rs2 = nullptr;
  PacbioReadSet* pb = nullptr;
  int n_reads = 0;
};

void FillReadSet(ReadSet* rs, int n_reads, const std::vector<int>& lens) {
  rs->reads_num_ = n_reads;
  rs->read_lens_ = lens;
  rs->CalcMaxReadLen();  // builds match_probs_/mismatch_probs_ pow tables, graph.cc:1443-1454
  rs->load_success_ = true;
}

}  // namespace

// Second mode: the reference's PacBio alignment probability (PacbioReadSet::AligmentProbability, graph.cc:2175-2297)
// on injected alignments. Input (GAMLAP1): f64 match, f64 mismatch, i32 band, i32 n; per alignment i32 posstart,
// i32 |s1|, s1, i32 |s2|, s2, i32 n_ops, n_ops x (i32 length, i32 op character). Output: n x f64 logval, then seconds.
int AlignProbMode(const char* in_path, const char* out_path) {
  Reader rd(in_path);
  if (memcmp(rd.buf.data(), "GAMLAP1\0", 8) != 0) { fprintf(stderr, "bad magic\n"); return 2; }
  rd.off = 8;
  const double match = rd.f64(), mismatch = rd.f64();
  const int band = rd.i32(), n = rd.i32();
  PacbioReadSet pb("/nonexistent/pbap", "", match, mismatch);
  std::vector<double> out(n);
  double secs = 0;
  for (int a = 0; a < n; a++) {
    PacbioReadSet::PacbioAligmentData ad;
    ad.posstart = rd.i32();
    const int n1 = rd.i32();
    std::string s1(&rd.buf[rd.off], &rd.buf[rd.off] + n1);
    rd.off += n1;
    const int n2 = rd.i32();
    std::string s2(&rd.buf[rd.off], &rd.buf[rd.off] + n2);
    rd.off += n2;
    const int n_ops = rd.i32();
    for (int k = 0; k < n_ops; k++) {
      const int len = rd.i32(), op = rd.i32();
      ad.cigar.push_back(make_pair(len, (char)op));
    }
    auto t0 = std::chrono::steady_clock::now();
    out[a] = pb.AligmentProbability(s1, s2, ad, band).logval;
    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  FILE* f = fopen(out_path, "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", out_path); return 2; }
  fwrite(out.data(), 8, out.size(), f);
  fwrite(&secs, 8, 1, f);
  fclose(f);
  fprintf(stderr, "ref_harness --alnprob: %d alignments, %.6f s in AligmentProbability\n", n, secs);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && strcmp(argv[1], "--alnprob") == 0) return AlignProbMode(argv[2], argv[3]);
  if (argc < 3) {
    fprintf(stderr, "usage: %s <workload> <results> [dump] [repeat]\n       %s --alnprob <alignments> <logvals>\n", argv[0], argv[0]);
    return 2;
  }
  const bool dump = argc > 3 && atoi(argv[3]) != 0;
  const int repeat = argc > 4 ? atoi(argv[4]) : 1;
  Reader rd(argv[1]);
  if (memcmp(rd.buf.data(), "GAMLWL1\0", 8) != 0) { fprintf(stderr, "bad magic\n"); return 2; }
  rd.off = 8;

  Graph gr;
  int n_nodes = rd.i32();
  std::vector<int> node_len(n_nodes), nmap(n_nodes);
  rd.ints(node_len.data(), n_nodes);
  rd.ints(nmap.data(), n_nodes);
  gr.nodes.resize(n_nodes);
  for (int i = 0; i < n_nodes; i++) {
    Node* n = new Node;
    n->id = i;
    n->s = string(node_len[i], 'A');  // scoring only ever uses s.length()
    gr.nodes[i] = n;
  }
  gr.normalize_map = nmap;

  vector<pair<SingleReadConfig, ReadSet*>> single_reads;
  vector<pair<PairedReadConfig, pair<ReadSet*, ReadSet*>>> paired_reads;
  vector<pair<SingleReadConfig, PacbioReadSet*>> pacbio_reads;
  std::vector<SetInfo> sets;
  long long total_records = 0;

  int n_sets = rd.i32();
  for (int s = 0; s < n_sets; s++) {
    SetInfo si;
    si.kind = rd.i32();
    double mismatch = rd.f64(), match = rd.f64(), ins_mean = rd.f64(), ins_std = rd.f64();
    double mppb = rd.f64(), mps = rd.f64(), weight = rd.f64(), pen = rd.f64(), step = rd.f64();
    si.n_reads = rd.i32();
    int n_mates = rd.i32();
    std::vector<std::vector<int>> lens(n_mates, std::vector<int>(si.n_reads));
    for (int m = 0; m < n_mates; m++) rd.ints(lens[m].data(), si.n_reads);
    char nm[64];
    if (si.kind == 2) {
      snprintf(nm, sizeof nm, "/nonexistent/pb%d", s);
      si.pb = new PacbioReadSet(nm, "", match, mismatch);
      si.pb->reads_num_ = si.n_reads;
      si.pb->read_lens_ = lens[0];
      si.pb->CalcMaxReadLen();
      si.pb->read_seq_.resize(si.n_reads);
      for (int i = 0; i < si.n_reads; i++) {
        char rn[32];
        snprintf(rn, sizeof rn, "r%d", i);
        si.pb->read_map_inv_[i] = rn;  // GetReadName is called per read per evaluation, graph.cc:3074
        si.pb->read_map_[rn] = i;
      }
      si.pb->load_success_ = true;
      int n_keys = rd.i32();
      for (int k = 0; k < n_keys; k++) {
        int kl = rd.i32();
        vector<int> key(kl);
        rd.ints(key.data(), kl);
        int nr = rd.i32();
        vector<PacbioReadSet::PacbioAligment>& dst = si.pb->aligment_cache_[key];
        dst.reserve(nr);
        for (int r = 0; r < nr; r++) {
          int pos = rd.i32(), pos_end = rd.i32(), rid = rd.i32();
          rd.i32();
          double lp = rd.f64();
          logdouble ld;
          ld.logval = lp;
          dst.push_back(PacbioReadSet::PacbioAligment(pos, pos_end, rid, ld));
        }
        total_records += nr;
      }
      pacbio_reads.push_back(make_pair(SingleReadConfig(pen, step, mppb, mps, weight, false), si.pb));
    } else {
      ReadSet* rss[2] = {nullptr, nullptr};
      for (int m = 0; m < n_mates; m++) {
        snprintf(nm, sizeof nm, "/nonexistent/rs%d_%d", s, m);
        rss[m] = new ReadSet(nm, "", match, mismatch);
        FillReadSet(rss[m], si.n_reads, lens[m]);
      }
      for (int m = 0; m < n_mates; m++) {
        int n_keys = rd.i32();
        for (int k = 0; k < n_keys; k++) {
          int kl = rd.i32();
          vector<int> key(kl);
          rd.ints(key.data(), kl);
          int nr = rd.i32();
          vector<Aligment>& dst = rss[m]->aligment_cache_[key];
          dst.resize(nr);
          static_assert(sizeof(Aligment) == 16, "Aligment is 4 x int32");
          rd.ints(reinterpret_cast<int*>(dst.data()), 4 * (size_t)nr);
          total_records += nr;
        }
      }
      si.rs1 = rss[0];
      si.rs2 = rss[1];
      if (si.kind == 0) {
        single_reads.push_back(make_pair(SingleReadConfig(pen, step, mppb, mps, weight, false), si.rs1));
      } else {
        paired_reads.push_back(make_pair(
            PairedReadConfig(pen, step, ins_mean, ins_std, mppb, mps, weight, false), make_pair(si.rs1, si.rs2)));
      }
    }
    sets.push_back(si);
  }

  int n_evals = rd.i32();
  std::vector<vector<vector<int>>> evals(n_evals);
  for (int e = 0; e < n_evals; e++) {
    int nw = rd.i32();
    evals[e].resize(nw);
    for (int w = 0; w < nw; w++) {
      int ln = rd.i32();
      evals[e][w].resize(ln);
      rd.ints(evals[e][w].data(), ln);
    }
  }

  // CalcProb reports sets in the order single, paired, pacbio (prob_calculator.h:70-107); map the
  // workload's set order onto that.
  std::vector<int> zero_slot(n_sets);
  {
    int ns = 0, np = 0, nb = 0;
    for (auto& si : sets) { if (si.kind == 0) ns++; else if (si.kind == 1) np++; else nb++; }
    int is = 0, ip = 0, ib = 0;
    for (int s = 0; s < n_sets; s++) {
      if (sets[s].kind == 0) zero_slot[s] = is++;
      else if (sets[s].kind == 1) zero_slot[s] = ns + ip++;
      else zero_slot[s] = ns + np + ib++;
    }
  }

  // The reference chats on stdout (and PacBio rewrites ./rp.dat every call, graph.cc:3071).
  fflush(stdout);
  FILE* devnull = freopen("/dev/null", "w", stdout);
  (void)devnull;

  FILE* out = fopen(argv[2], "wb");
  if (!out) { fprintf(stderr, "cannot write %s\n", argv[2]); return 2; }
  fwrite("GAMLRS1\0", 1, 8, out);
  int hdr[3] = {n_evals * repeat, n_sets, dump ? 1 : 0};
  fwrite(hdr, 4, 3, out);

  double total_secs = 0;
  for (int rep = 0; rep < repeat; rep++) {
    ProbCalculator pc(single_reads, paired_reads, pacbio_reads, gr);  // fresh ScoringState per repeat
    for (int e = 0; e < n_evals; e++) {
      vector<pair<int, int>> zeros;
      int total_len = 0;
      auto t0 = std::chrono::steady_clock::now();
      double score = pc.CalcProb(evals[e], zeros, total_len);
      auto t1 = std::chrono::steady_clock::now();
      double secs = std::chrono::duration<double>(t1 - t0).count();
      total_secs += secs;
      fwrite(&score, 8, 1, out);
      int tl[2] = {total_len, 0};
      fwrite(tl, 4, 2, out);
      fwrite(&secs, 8, 1, out);
      for (int s = 0; s < n_sets; s++) {
        int z[2] = {zeros[zero_slot[s]].first, zeros[zero_slot[s]].second};
        fwrite(z, 4, 2, out);
      }
      if (dump) {
        int ip = 0;
        for (int s = 0; s < n_sets; s++) {
          std::vector<double> vals;
          if (sets[s].kind == 0) {
            vector<double> rp;
            PositionsToReadProbs(sets[s].n_reads, sets[s].rs1->GetPositions(), *sets[s].rs1, rp);
            vals = rp;
          } else if (sets[s].kind == 1) {
            vals = pc.paired_scoring_states[ip++].probs;
          } else {
            // Rebuild read_probs the way CalcScoreForPacbio does (graph.cc:3176-3223).
            vector<logdouble> rp(sets[s].n_reads);
            for (auto path : evals[e]) {
              gr.NormalizePath(path);
              int tl2;
              AddPositionsToReadProbsPacbio(sets[s].pb->GetReadProbabilities(gr, path, tl2), rp);
            }
            vals.resize(rp.size());
            for (size_t i = 0; i < rp.size(); i++) vals[i] = rp[i].logval;
          }
          int n[2] = {(int)vals.size(), 0};
          fwrite(n, 4, 2, out);
          fwrite(vals.data(), 8, vals.size(), out);
        }
      }
    }
  }
  fclose(out);
  fprintf(stderr, "ref_harness: %d evals x %d, %lld cached records, %.6f s in CalcProb\n", n_evals, repeat,
          total_records, total_secs);
  return 0;
}
