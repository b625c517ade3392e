#!/usr/bin/env python
"""Developer tool (GPU box): throughput of gaml_pacbio_alignment_logprob against the CPU restatement (oracle) on the
same alignments. Prints alignments/s and DP cells/s for both."""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaml_b200 import api, alnprob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
read_len = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 100
base = alnprob.make_alignments(200, read_len, seed=7)
alns = base * copies
rows = sum(sum(l for l, op in a.cigar if op in "MD") for a in base)
cells = rows * 7 * copies   # ~ (2 band + 1) + 2 columns per row
pc = api.ProbCalculator([100], None)
pc.pacbio_alignment_logprob(base, 0.85, 0.05, 2)
flat = alnprob.flatten(alns)   # the C ABI's layout (what a C++ caller holds)
for rep in range(4):
    t0 = time.perf_counter()
    got = pc.pacbio_alignment_logprob_flat(flat, 0.85, 0.05, 2)
    dt = time.perf_counter() - t0
    st = pc.stats()
    print(f"cuda: {len(alns)} alignments of ~{read_len} bases in {dt * 1e3:.1f} ms through the C ABI (host arrays in, logvals out) -> {len(alns) / dt:.0f} alignments/s; "
          f"kernel {st.last_device_ms:.1f} ms -> {len(alns) / st.last_device_ms * 1e3:.0f} alignments/s, {cells / st.last_device_ms / 1e6:.2f} G cells/s; "
          f"host CIGAR summaries + upload enqueue {st.last_prepare_host_us / 1e3:.1f} ms", flush=True)
with tempfile.TemporaryDirectory() as tmp:
    ap, lp = os.path.join(tmp, "a.ap"), os.path.join(tmp, "a.lp")
    alnprob.write_alignments(ap, base, 0.85, 0.05, 2)
    for binary in ("oracle/gaml_oracle", "oracle/_ref/ref_harness"):
        b = os.path.join(ROOT, binary)
        if not os.path.exists(b):
            continue
        subprocess.run([b, "--alnprob", ap, lp], check=True, stderr=subprocess.DEVNULL)
        ref, secs = alnprob.read_logvals(lp, len(base))
        print(f"{binary}: {len(base)} alignments in {secs * 1e3:.1f} ms -> {len(base) / secs:.0f} alignments/s on one core", flush=True)
    fin = np.isfinite(ref)
    print("max relative difference cuda vs", binary, float(np.max(np.abs(got[:len(base)][fin] - ref[fin]) / np.abs(ref[fin]))))
