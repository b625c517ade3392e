#!/usr/bin/env python
"""Developer tool (GPU box): host cost of every kernel launch of a few evaluations (GAML_B200_TRACE_HOST=1)."""
import os, sys
os.environ["GAML_B200_TRACE_HOST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaml_b200 import api, synth
wl = synth.paired_workload(460, 10000, 2_000_000, n_evals=6, seed=42)
pc = api.ProbCalculator.from_workload(wl)
for rep in range(3):
    pc.reset_state()
    for w in wl.evals[:4]:
        print("---- evaluation", file=sys.stderr, flush=True)
        pc.calc_prob_partial(w)
        s = pc.stats()
        print(f"prepare {s.last_prepare_host_us:.1f} launch {s.last_launch_host_us:.1f} finish {s.last_finish_host_us:.1f}", file=sys.stderr, flush=True)
