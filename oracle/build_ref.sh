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
  echo "build_ref: built $OUT/gaml_gpu (reference annealing loop + moves over the CUDA ProbCalculator) and $OUT/gpu_harness"
fi
echo "build_ref: built $OUT/gaml_ref and $OUT/ref_harness"
