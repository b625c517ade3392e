#!/usr/bin/env bash
# SA iterations/s of the reference's UNCHANGED annealing loop (gaml.cc Optimize + moves.cc) on a warm cache, for the three
# builds of oracle/build_ref.sh: gaml_ref (reference ProbCalculator, CPU), gaml_gpu (drop-in CUDA ProbCalculator),
# gaml_gpu_batched (the same + the moves' candidate lists scored in device batches). One synthetic paired-end data set
# (tools/make_e2e_dataset.py); every binary runs it twice — 1 iteration and N iterations — so that read loading, index
# building and the first alignments drop out: rate = (N - 1) / (T_N - T_1). Traces must be identical.
#   tools/sa_rate.sh <outdir> <N iterations> <unique 10-kbp nodes> <read pairs>
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=${1:-$ROOT/gpurun_out/sa_rate}; N=${2:-1000}; UNIQUE=${3:-100}; READS=${4:-400000}
mkdir -p "$OUT"
D=/tmp/sa_rate_data
rm -rf "$D"
python "$ROOT/tools/make_e2e_dataset.py" "$D" --kind paired --n-unique "$UNIQUE" --n-reads "$READS" --iterations "$N" > "$OUT/gen.log"
sed "s/^max_iterations=.*/max_iterations=1/" "$D/gaml.cfg" > "$D/gaml1.cfg"
{
echo "data set: $(cat $OUT/gen.log)"
for bin in gaml_ref gaml_gpu gaml_gpu_batched; do
  [ -x "$ROOT/oracle/_ref/$bin" ] || continue
  t0=$(date +%s.%N); ( cd "$D" && stdbuf -oL "$ROOT/oracle/_ref/$bin" gaml1.cfg > $bin.1.log 2>&1 ); t1=$(date +%s.%N)
  ( cd "$D" && stdbuf -oL "$ROOT/oracle/_ref/$bin" gaml.cfg > $bin.log 2>&1 ); t2=$(date +%s.%N)
  python -c "
t1, tn, n = $t1 - $t0, $t2 - $t1, $N
print('$bin: %d iterations in %.2f s, 1 iteration in %.2f s -> %.1f SA iterations/s (loop alone), %.2f ms per iteration' % (n, tn, t1, (n - 1) / max(tn - t1, 1e-9), 1e3 * (tn - t1) / (n - 1)))"
done
for bin in gaml_gpu gaml_gpu_batched; do
  [ -f "$D/$bin.log" ] && python "$ROOT/tools/compare_traces.py" "$D/gaml_ref.log" "$D/$bin.log" | tail -2
done
} | tee "$OUT/sa_rate.txt"
