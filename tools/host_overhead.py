#!/usr/bin/env python
"""Developer tool (GPU box): where does the wall time of an incremental CalcProb go? Python flattening vs the C
ABI's prepare (host flatten + H2D enqueue) / launch (kernel enqueue) / finish (D2H + sync)."""
import os, sys, time
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaml_b200 import api, synth

wl = synth.paired_workload(460, 10000, 2_000_000, n_evals=202, seed=42)
pc = api.ProbCalculator.from_workload(wl)
pc.calc_prob(wl.evals[0])
seq = wl.evals[1:201]
flat = [api.flatten_walks(w) for w in seq]
t = {"py_flatten": 0.0, "prepare": 0.0, "launch": 0.0, "finish": 0.0}
lib, h = pc.lib, pc.h
part = np.zeros(api.PARTIAL_DOUBLES, dtype=np.float64); tl = C.c_int32()
i64p, i32p, dp = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
for rep in range(3):
    for k in t: t[k] = 0.0
    pc.reset_state(); pc.calc_prob(wl.evals[0])
    for w, (nodes, offs) in zip(seq, flat):
        t0 = time.perf_counter(); api.flatten_walks(w); t1 = time.perf_counter()
        lib.gaml_eval_prepare(h, nodes.ctypes.data_as(i32p), offs.ctypes.data_as(i64p), len(w)); t2 = time.perf_counter()
        lib.gaml_eval_launch(h); t3 = time.perf_counter()
        lib.gaml_eval_finish(h, part.ctypes.data_as(dp), C.byref(tl)); t4 = time.perf_counter()
        t["py_flatten"] += t1 - t0; t["prepare"] += t2 - t1; t["launch"] += t3 - t2; t["finish"] += t4 - t3
    print({k: round(1e6 * v / len(seq), 1) for k, v in t.items()}, "us per incremental eval; device ms", pc.stats().last_device_ms)
