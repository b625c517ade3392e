#!/usr/bin/env python
"""Developer check (GPU box): CUDA path vs oracle on a battery of small seeded workloads.

The pytest suite (tests/ -m gpu) is the judged version of this; the script prints more detail.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaml_b200 import api, synth, workload  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle", "gaml_oracle")


def run_oracle(wl, tmp, name):
    p = os.path.join(tmp, name + ".wl")
    r = os.path.join(tmp, name + ".res")
    workload.write_workload(p, wl)
    subprocess.run([ORACLE, p, r, "1"], check=True, stderr=subprocess.DEVNULL, cwd=tmp)
    return workload.read_results(r)


def compare(name, wl, tmp, verbose=True):
    ref = run_oracle(wl, tmp, name)
    pc = api.ProbCalculator.from_workload(wl)
    worst_total = 0.0
    worst_read = 0.0
    bit_exact_state = True
    ok = True
    for e, walks in enumerate(wl.evals):
        prob, zeros, tl = pc.calc_prob(walks)
        r = ref[e]
        rel = abs(prob - r.score) / max(abs(r.score), 1e-300)
        worst_total = max(worst_total, rel)
        if tl != r.total_len or zeros != r.zeros or not (rel <= 1e-9):
            ok = False
            print(f"  [{name}] eval {e}: MISMATCH prob {prob!r} vs {r.score!r} rel {rel:.3e} tl {tl}/{r.total_len} zeros {zeros}/{r.zeros}")
        for s, spec in enumerate(wl.sets):
            v = pc.read_values(s)
            rv = r.per_read[s]
            if spec.kind == workload.KIND_PACBIO:
                fin = np.isfinite(rv)
                if not np.array_equal(fin, np.isfinite(v)):
                    ok = False
                    print(f"  [{name}] eval {e} set {s}: -inf pattern differs")
                d = np.abs(v[fin] - rv[fin]) / np.maximum(np.abs(rv[fin]), 1e-300)
                worst_read = max(worst_read, float(d.max()) if d.size else 0.0)
            else:
                if not np.array_equal(v, rv):
                    bit_exact_state = False
                    nz = rv != 0
                    d = np.abs(v[nz] - rv[nz]) / np.abs(rv[nz])
                    worst_read = max(worst_read, float(d.max()) if d.size else 0.0)
                    bad = np.nonzero(v != rv)[0]
                    print(f"  [{name}] eval {e} set {s}: {len(bad)} per-read values differ, first {bad[:4]} {v[bad[:4]]} {rv[bad[:4]]}")
    st = pc.stats()
    print(f"{name}: {'OK ' if ok else 'FAIL'} evals={len(wl.evals)} worst total rel={worst_total:.2e} "
          f"per-read bit-exact={bit_exact_state} worst per-read rel={worst_read:.2e} launches={st.kernel_launches} "
          f"overflow(last)={st.last_overflow_reads}")
    pc.close()
    return ok and worst_read <= 1e-12


def main():
    ok = True
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(3):
            ok &= compare(f"single{seed}", synth.single_workload(10, 2500, 3000, n_evals=25, seed=seed), tmp)
            ok &= compare(f"paired{seed}", synth.paired_workload(14, 2500, 5000, n_evals=40, seed=seed), tmp)
            ok &= compare(f"mixed{seed}", synth.mixed_workload(10, 6000, 3000, 300, n_single=1000, n_evals=20,
                                                               seed=seed, pacbio_len=5000), tmp)
        ok &= compare("paired_big", synth.paired_workload(60, 8000, 200000, n_evals=30, seed=9), tmp)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
