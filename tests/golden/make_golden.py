#!/usr/bin/env python
"""Generates tests/golden/*.wl (inputs) and *.ref.res (outputs of the REAL reference).

Run in the build container, where /root/reference exists: oracle/build_ref.sh compiles the reference's
own graph.cc / prob_calculator.h into oracle/_ref/ref_harness, and every expected value below comes out of
that binary — nothing is computed by this repo's own code. The fixtures are committed so the GPU box
(which has no /root/reference) can still pin the oracle and the CUDA path against the reference.
"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cases import golden_cases  # noqa: E402
from gaml_b200 import workload  # noqa: E402


def main():
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        subprocess.run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")], check=True)
    for name, wl in golden_cases().items():
        wp = os.path.join(HERE, name + ".wl")
        rp = os.path.join(HERE, name + ".ref.res")
        workload.write_workload(wp, wl)
        with tempfile.TemporaryDirectory() as tmp:   # the reference drops rp.dat into its cwd
            subprocess.run([harness, wp, rp, "1"], check=True, cwd=tmp)
        res = workload.read_results(rp)
        print(f"{name}: {len(res)} evals, {os.path.getsize(wp)} + {os.path.getsize(rp)} bytes, "
              f"scores {res[0].score!r} .. {res[-1].score!r}")
    alnprob_golden(harness)


def alnprob_golden(harness):
    """PacBio alignment probabilities (PacbioReadSet::AligmentProbability) of synthetic alignments, from the reference."""
    from cases import alnprob_cases
    from gaml_b200 import alnprob
    for name, (alns, match, mismatch, band) in alnprob_cases().items():
        ip = os.path.join(HERE, name + ".ap")
        rp = os.path.join(HERE, name + ".ref.lp")
        alnprob.write_alignments(ip, alns, match, mismatch, band)
        subprocess.run([harness, "--alnprob", ip, rp], check=True)
        vals, _ = alnprob.read_logvals(rp, len(alns))
        with open(rp, "wb") as f:          # keep the values only (the harness appends its run time)
            f.write(vals.astype("<f8").tobytes())
        print(f"{name}: {len(alns)} alignments, {os.path.getsize(ip)} + {os.path.getsize(rp)} bytes, logvals {vals[:3]}")


if __name__ == "__main__":
    main()
