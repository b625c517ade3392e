#!/usr/bin/env bash
# SA iterations/s of the reference's UNCHANGED annealing loop (gaml.cc Optimize + moves.cc) on a warm cache, for the three
# builds of oracle/build_ref.sh: gaml_ref (reference ProbCalculator, CPU), gaml_gpu (drop-in CUDA ProbCalculator),
# gaml_gpu_batched (the same + the moves' candidate lists scored in device batches). One synthetic paired-end data set
# (tools/make_e2e_dataset.py). The driver prints one "itnum" line per iteration (gaml.cc:330); each binary's stdout is
# timestamped line by line, and the rate is taken between iteration WARM and the last one, so read loading, index building,
# context creation and the first (cache-filling) iterations drop out. Traces must be identical.
#   tools/sa_rate.sh <outdir> <N iterations> <unique 10-kbp nodes> <read pairs> [warm-up iterations]
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=${1:-$ROOT/gpurun_out/sa_rate}; N=${2:-1000}; UNIQUE=${3:-100}; READS=${4:-400000}; WARM=${5:-100}
mkdir -p "$OUT"
D=/tmp/sa_rate_data
rm -rf "$D"
python "$ROOT/tools/make_e2e_dataset.py" "$D" --kind paired --n-unique "$UNIQUE" --n-reads "$READS" --iterations "$N" > "$OUT/gen.log"
{
echo "data set: $(cat $OUT/gen.log)"
for bin in gaml_ref gaml_gpu gaml_gpu_batched; do
  [ -x "$ROOT/oracle/_ref/$bin" ] || continue
  ( cd "$D" && stdbuf -oL "$ROOT/oracle/_ref/$bin" gaml.cfg 2>/dev/null | python -u -c "
import sys, time
t = []
with open('$bin.log', 'w') as f:
    for line in sys.stdin:
        f.write(line)
        if line.startswith('itnum '):
            t.append(time.perf_counter())
w = min($WARM, max(len(t) - 2, 0))
if len(t) > w + 1:
    dt = t[-1] - t[w]
    print('$bin: %d iterations, iterations %d..%d in %.3f s -> %.1f SA iterations/s, %.3f ms per iteration' % (len(t), w, len(t) - 1, dt, (len(t) - 1 - w) / dt, 1e3 * dt / (len(t) - 1 - w)))
else:
    print('$bin: too few iterations (%d)' % len(t))
" )
done
for bin in gaml_gpu gaml_gpu_batched; do
  [ -f "$D/$bin.log" ] && python "$ROOT/tools/compare_traces.py" "$D/gaml_ref.log" "$D/$bin.log" | tail -2
done
} | tee "$OUT/sa_rate.txt"
