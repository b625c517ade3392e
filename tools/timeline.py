#!/usr/bin/env python
"""Developer tool (GPU box): device timeline (globaltimer stamps) of full and incremental evaluations, L2 flushed
before each one (like bench.py's timed region)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaml_b200 import api, synth
wl = synth.paired_workload(460, 10000, 2_000_000, n_evals=12, seed=42)
pc = api.ProbCalculator.from_workload(wl)
flat = [api.FlatWalks(w) for w in wl.evals]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
sink = torch.zeros((), dtype=torch.int64, device="cuda")


def flush_l2():
    flush.fill_(1)
    sink.copy_(flush.view(torch.int32).sum())
    torch.cuda.synchronize()


pc.set_profiling(3)
for rep in range(4):
    pc.reset_state()
    flush_l2()
    pc.calc_prob_partial_flat(flat[0])
    print("full       ", {k: (round(a, 1), round(b, 1)) for k, (a, b) in pc.read_timeline().items()}, flush=True)
for it in flat[1:6]:
    flush_l2()
    pc.calc_prob_partial_flat(it)
    print("incremental", {k: (round(a, 1), round(b, 1)) for k, (a, b) in pc.read_timeline().items()}, flush=True)
