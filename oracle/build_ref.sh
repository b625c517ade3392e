#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the real reference (usamec/GAML) scoring code from the sources
# WHERE THEY LIE under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box):
#   oracle/_ref/gaml_ref      the reference's own `gaml` binary (gaml.cc main + Optimize + moves)
#   oracle/_ref/ref_harness   cache-injection harness (oracle/ref_harness.cc) around ProbCalculator
# No reference source is copied. Two deviations, both recorded in DESIGN.md:
#   * Boost is not installed: -I oracle/boost_shim supplies split/is_any_of and no-op archives, and the
#     four vendored Boost-serialization helpers (unordered_*.hpp) are switched off through their own
#     include guards (their only job is on-disk archives, which the harness never touches).
#   * graph.cc:1478 returns a reference to a temporary (segfaults under g++ 13 -O2); the one line is
#     patched in the compiler's input stream (sed | g++ -x c++ -), meaning "empty alignment list".
set -euo pipefail
REF=${GAML_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/graph.cc" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
FLAGS=(-std=c++11 -O2 -w
       -DBOOST_SERIALIZATION_UNORDERED_MAP_HPP -DBOOST_SERIALIZATION_UNORDERED_SET_HPP
       -include unordered_map -include unordered_set -I"$HERE/boost_shim" -I"$REF")
cd "$OUT"
sed '1478s/return vector<Aligment>();/{ static const vector<Aligment> kEmptyAl; return kEmptyAl; }/' "$REF/graph.cc" \
  | g++ "${FLAGS[@]}" -x c++ - -c -o graph.o &
for f in moves input_output graph_from_assembly gaml; do
  g++ "${FLAGS[@]}" -c "$REF/$f.cc" -o "$f.o" &
done
g++ "${FLAGS[@]}" -c "$HERE/ref_harness.cc" -o ref_harness.o &
wait
g++ -o gaml_ref gaml.o graph.o moves.o input_output.o graph_from_assembly.o 2>/dev/null
g++ -o ref_harness ref_harness.o graph.o
# gaml_gpu: the reference's OWN gaml.cc / moves.cc (annealing loop + moves, unchanged) compiled against the
# drop-in integration/prob_calculator.h and linked with the CUDA library. -Dprivate=public lets the adapter
# read ReadSet::aligment_cache_; graph.o etc. are the reference objects built above.
LIBDIR="$HERE/../gaml_b200"
if [ -f "$LIBDIR/libgaml_b200.so" ]; then
  GFLAGS=(-std=c++11 -O2 -w -Dprivate=public
          -DBOOST_SERIALIZATION_UNORDERED_MAP_HPP -DBOOST_SERIALIZATION_UNORDERED_SET_HPP
          -include unordered_map -include unordered_set -include "$HERE/../integration/prob_calculator.h"
          -I"$HERE/../include" -I"$HERE/boost_shim" -I"$REF")
  # (quote-includes search the including file's own directory first, so the reference's prob_calculator.h would
  #  win over any -I; force-including the adapter first defines the shared include guard PROB_CALCULATOR_H__
  #  and turns the reference header into a no-op.)
  g++ "${GFLAGS[@]}" -c "$REF/gaml.cc" -o gaml_gpu.o &
  g++ "${GFLAGS[@]}" -c "$REF/moves.cc" -o moves_gpu.o &
  wait
  # gpu_harness: the SAME cache-injection harness as ref_harness, compiled against the drop-in adapter instead of the
  # reference's prob_calculator.h — every read-set kind (PacBio included) through ProbCalculator::CalcProb over the CUDA
  # library, and the per-call cost of the drop-in boundary (bench.py's dropin leg).
  g++ "${GFLAGS[@]}" -c "$HERE/ref_harness.cc" -o gpu_harness.o &
  wait
  g++ -o gaml_gpu gaml_gpu.o moves_gpu.o graph.o input_output.o graph_from_assembly.o \
      -L"$LIBDIR" -lgaml_b200 '-Wl,-rpath,$ORIGIN/../../gaml_b200' 2>/dev/null
  g++ -o gpu_harness gpu_harness.o graph.o -L"$LIBDIR" -lgaml_b200 '-Wl,-rpath,$ORIGIN/../../gaml_b200'
  # gaml_gpu_batched: the same program with the three places where moves.cc scores a LIST of alternatives
  # one CalcProb at a time rewritten — in the compiler's input stream, like the graph.cc:1478 line — to hand the whole
  # list to ProbCalculator::CalcProbBatch (one device batch) through GamlBatchReplay (integration/prob_calculator.h):
  #   LocalChange2     moves.cc:108-113   the two sampled extensions of each step
  #   FixGapLength     moves.cc:717-720   the two interior points of the ternary search
  #   FixRepForNode2   moves.cc:1158-1305 the three candidate loops (pairs of positions, doubles, palindromes)
  sed -e '108i\    GamlBatchReplay gb(prob_calc); for (int gb_ph = 0; gb_ph < 2; gb_ph++) { gb.Begin(gb_ph); scores.clear();' \
      -e '110s/prob_calc.CalcProb(new_paths)/gb.Score(new_paths)/' \
      -e '113a\    gb.End(); }' \
      -e '717i\  double mid1_p = 0, mid2_p = 0; { GamlBatchReplay gb(prob_calc); for (int gb_ph = 0; gb_ph < 2; gb_ph++) { gb.Begin(gb_ph);' \
      -e '718s/double mid1_p = prob_calc.CalcProb(paths)/mid1_p = gb.Score(paths)/' \
      -e '720s/double mid2_p = prob_calc.CalcProb(paths)/mid2_p = gb.Score(paths)/' \
      -e '720a\  gb.End(); } }' \
      -e '1157a\  GamlBatchReplay gb(prob_calc);' \
      -e '1158i\  for (int gb_ph = 0; gb_ph < 2; gb_ph++) { gb.Begin(gb_ph);' \
      -e '1189s/prob_calc.CalcProb(paths2)/gb.Score(paths2)/' \
      -e '1204a\  gb.End(); }' \
      -e '1205i\  for (int gb_ph = 0; gb_ph < 2; gb_ph++) { gb.Begin(gb_ph);' \
      -e '1264s/prob_calc.CalcProb(paths2)/gb.Score(paths2)/' \
      -e '1281a\  gb.End(); }' \
      -e '1282i\  for (int gb_ph = 0; gb_ph < 2; gb_ph++) { gb.Begin(gb_ph);' \
      -e '1290s/prob_calc.CalcProb(paths2)/gb.Score(paths2)/' \
      -e '1305a\  gb.End(); }' \
      "$REF/moves.cc" > moves_batched.ii.cc
  if [ "$(grep -c 'gb.Score' moves_batched.ii.cc)" != "6" ]; then echo "build_ref: moves.cc patch did not apply (reference changed?)" >&2; exit 1; fi
  g++ "${GFLAGS[@]}" -I"$REF" -c moves_batched.ii.cc -o moves_gpu_batched.o
  rm -f moves_batched.ii.cc
  g++ -o gaml_gpu_batched gaml_gpu.o moves_gpu_batched.o graph.o input_output.o graph_from_assembly.o \
      -L"$LIBDIR" -lgaml_b200 '-Wl,-rpath,$ORIGIN/../../gaml_b200' 2>/dev/null
  echo "build_ref: built $OUT/gaml_gpu (reference annealing loop + moves over the CUDA ProbCalculator) and $OUT/gpu_harness"
fi
echo "build_ref: built $OUT/gaml_ref and $OUT/ref_harness"
