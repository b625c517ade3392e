#!/usr/bin/env python
"""Developer tool (GPU box): where does the wall time of a CalcProb through the C ABI go?
Python side: ctypes call; library side (gaml_stats): prepare (walk flattening + H2D enqueue), launch (kernel
enqueue), finish (D2H + stream synchronize, i.e. includes waiting for the device); device time from CUDA events."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaml_b200 import api, synth

n_unique = int(sys.argv[1]) if len(sys.argv) > 1 else 460        # walks (nodes) of the workload: 460 = config 2, 10000 = config 4
n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
wl = synth.paired_workload(n_unique, 10000, n_pairs, n_evals=202, seed=42)
pc = api.ProbCalculator.from_workload(wl)
pc.set_profiling(1)
flat0 = api.FlatWalks(wl.evals[0])
seq = [api.FlatWalks(w) for w in wl.evals[1:201]]


def run(kind, items, reset):
    acc = {"wall": 0.0, "prepare": 0.0, "launch": 0.0, "finish": 0.0, "device": 0.0}
    for it in items:
        if reset:
            pc.reset_state()
        t0 = time.perf_counter()
        pc.calc_prob_partial_flat(it)
        acc["wall"] += 1e6 * (time.perf_counter() - t0)
        s = pc.stats()
        acc["prepare"] += s.last_prepare_host_us
        acc["launch"] += s.last_launch_host_us
        acc["finish"] += s.last_finish_host_us
        acc["device"] += 1e3 * s.last_device_ms
    print(kind, {k: round(v / len(items), 1) for k, v in acc.items()}, "us per evaluation", flush=True)


for rep in range(3):
    run("full       ", [flat0] * 20, True)
    pc.reset_state()
    pc.calc_prob_partial_flat(flat0)
    run("incremental", seq, False)

# device timeline of one full and a few incremental evaluations (globaltimer stamps, chain intact)
pc.set_profiling(3)
pc.reset_state()
pc.calc_prob_partial_flat(flat0)
print("timeline full       ", {k: (round(a, 1), round(b, 1)) for k, (a, b) in pc.read_timeline().items()}, flush=True)
for it in seq[:6]:
    pc.calc_prob_partial_flat(it)
    print("timeline incremental", {k: (round(a, 1), round(b, 1)) for k, (a, b) in pc.read_timeline().items()}, flush=True)
pc.set_profiling(0)
