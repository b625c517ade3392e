// GPU test of the C++ host mirror (gaml_b200/host/prob_calculator.h): drives ProbCalculator::CalcProb over
// a GAMLWL1 workload exactly as gaml.cc does with the reference's class, and compares with a GAMLRS1 result
// file written by the real reference (tests/golden/*.ref.res). Prints PASS / FAIL.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "prob_calculator.h"

using namespace gaml_b200;

struct Reader {
  std::vector<char> buf;
  size_t off = 8;
  explicit Reader(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n);
    if (fread(buf.data(), 1, n, f) != (size_t)n) exit(2);
    fclose(f);
  }
  int i32() { int v; memcpy(&v, &buf[off], 4); off += 4; return v; }
  double f64() { double v; memcpy(&v, &buf[off], 8); off += 8; return v; }
  void raw(void* d, size_t n) { if (n) memcpy(d, &buf[off], n); off += n; }
};

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s workload.wl reference.res\n", argv[0]); return 2; }
  Reader rd(argv[1]);
  Graph gr;
  int n_nodes = rd.i32();
  gr.node_len.resize(n_nodes);
  gr.normalize_map.resize(n_nodes);
  rd.raw(gr.node_len.data(), 4 * (size_t)n_nodes);
  rd.raw(gr.normalize_map.data(), 4 * (size_t)n_nodes);
  std::vector<std::pair<SingleReadConfig, ReadSet*>> single_reads;
  std::vector<std::pair<PairedReadConfig, std::pair<ReadSet*, ReadSet*>>> paired_reads;
  std::vector<std::pair<SingleReadConfig, PacbioReadSet*>> pacbio_reads;
  int n_sets = rd.i32();
  std::vector<int> kinds;
  for (int s = 0; s < n_sets; s++) {
    int kind = rd.i32();
    kinds.push_back(kind);
    double mismatch = rd.f64(), match = rd.f64(), im = rd.f64(), is = rd.f64(), mppb = rd.f64(), mps = rd.f64(),
           weight = rd.f64(), pen = rd.f64(), step = rd.f64();
    int n_reads = rd.i32(), n_mates = rd.i32();
    std::vector<std::vector<int>> lens(n_mates, std::vector<int>(n_reads));
    for (auto& l : lens) rd.raw(l.data(), 4 * (size_t)n_reads);
    if (kind == 2) {
      PacbioReadSet* pb = new PacbioReadSet("pb", "", match, mismatch);
      pb->SetReadLens(lens[0]);
      int nk = rd.i32();
      for (int k = 0; k < nk; k++) {
        std::vector<int> key(rd.i32());
        rd.raw(key.data(), 4 * key.size());
        std::vector<PacbioAligment> recs(rd.i32());
        rd.raw(recs.data(), sizeof(PacbioAligment) * recs.size());
        pb->InsertAligments(key, recs);
      }
      pacbio_reads.push_back({SingleReadConfig(pen, step, mppb, mps, weight, false), pb});
    } else {
      ReadSet* rs[2] = {nullptr, nullptr};
      for (int m = 0; m < n_mates; m++) {
        rs[m] = new ReadSet("rs", "", match, mismatch);
        rs[m]->SetReadLens(lens[m]);
      }
      for (int m = 0; m < n_mates; m++) {
        int nk = rd.i32();
        for (int k = 0; k < nk; k++) {
          std::vector<int> key(rd.i32());
          rd.raw(key.data(), 4 * key.size());
          std::vector<Aligment> recs(rd.i32());
          rd.raw(recs.data(), sizeof(Aligment) * recs.size());
          rs[m]->InsertAligments(key, recs);
        }
      }
      if (kind == 0) single_reads.push_back({SingleReadConfig(pen, step, mppb, mps, weight, false), rs[0]});
      else paired_reads.push_back({PairedReadConfig(pen, step, im, is, mppb, mps, weight, false), {rs[0], rs[1]}});
    }
  }
  int n_evals = rd.i32();
  std::vector<std::vector<std::vector<int>>> evals(n_evals);
  for (auto& ws : evals) {
    ws.resize(rd.i32());
    for (auto& w : ws) {
      w.resize(rd.i32());
      rd.raw(w.data(), 4 * w.size());
    }
  }
  // slot of workload set s in CalcProb's zeros order (single, paired, pacbio)
  std::vector<int> slot(n_sets);
  {
    int ns = 0, np = 0;
    for (int k : kinds) { ns += k == 0; np += k == 1; }
    int is = 0, ip = 0, ib = 0;
    for (int s = 0; s < n_sets; s++) slot[s] = kinds[s] == 0 ? is++ : kinds[s] == 1 ? ns + ip++ : ns + np + ib++;
  }
  Reader rr(argv[2]);
  int r_evals = rr.i32(), r_sets = rr.i32(), dump = rr.i32();
  if (r_evals != n_evals || r_sets != n_sets) { printf("FAIL: result file shape\n"); return 1; }

  ProbCalculator pc(single_reads, paired_reads, pacbio_reads, gr);
  bool ok = true;
  double worst = 0;
  for (int e = 0; e < n_evals; e++) {
    std::vector<std::pair<int, int>> zeros;
    int total_len = -1;
    double prob = pc.CalcProb(evals[e], zeros, total_len);
    double ref = rr.f64();
    int ref_tl = rr.i32();
    rr.i32();
    rr.f64();
    for (int s = 0; s < n_sets; s++) {
      int z = rr.i32(), n = rr.i32();
      if (zeros[slot[s]].first != z || zeros[slot[s]].second != n) {
        printf("eval %d set %d: zeros %d/%d vs %d/%d\n", e, s, zeros[slot[s]].first, zeros[slot[s]].second, z, n);
        ok = false;
      }
    }
    if (dump)
      for (int s = 0; s < n_sets; s++) {
        int n = rr.i32();
        rr.i32();
        rr.off += 8 * (size_t)n;
      }
    double rel = std::fabs(prob - ref) / std::fabs(ref);
    worst = std::max(worst, rel);
    if (!(rel <= 1e-9) || total_len != ref_tl) {
      printf("eval %d: prob %.17g vs %.17g, total_len %d vs %d\n", e, prob, ref, total_len, ref_tl);
      ok = false;
    }
  }
  printf("%s: %d evaluations, worst relative difference %.3g\n", ok ? "PASS" : "FAIL", n_evals, worst);
  return ok ? 0 : 1;
}
