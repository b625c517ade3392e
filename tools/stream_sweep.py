#!/usr/bin/env python
"""Measurement aid: the streaming kernel's (resident blocks, reads per lane) shapes on one workload, one process.
  python tools/stream_sweep.py [c2|c4shard] [shape ...]      e.g. 42 52 62 44 54 81 notab
Prints per shape: streaming-kernel ms (events around it), whole-evaluation device ms (graph), tier timeline."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def relabel_by_key(wl):
    """Experiment: renumber the reads so that ids ascend with (first key of mate 1, edit 1, edit 2, mate distance)."""
    spec = wl.sets[0]
    n = spec.n_reads
    big = np.iinfo(np.int64).max
    fk = [np.full(n, big, dtype=np.int64) for _ in range(2)]
    fe = [np.zeros(n, dtype=np.int64) for _ in range(2)]
    fp = [np.zeros(n, dtype=np.int64) for _ in range(2)]
    for m in range(2):
        for ki, (key, recs) in enumerate(spec.caches[m].items()):
            r = recs["read_id"]
            new = fk[m][r] == big
            rr = r[new]
            # first record of a read under this key only (records are position-sorted; keep the first occurrence)
            _, first_idx = np.unique(rr, return_index=True)
            sel = np.nonzero(new)[0][first_idx]
            fk[m][r[sel]] = ki
            fe[m][r[sel]] = recs["edit_dist"][sel]
            fp[m][r[sel]] = recs["position"][sel]
    order = np.lexsort((np.abs(fp[1] - fp[0]), fe[1], fe[0], fk[0]))
    newid = np.empty(n, dtype=np.int64)
    newid[order] = np.arange(n)
    for m in range(2):
        for key, recs in spec.caches[m].items():
            recs["read_id"] = newid[recs["read_id"]].astype(recs["read_id"].dtype)
            o = np.lexsort((recs["read_id"], recs["position"]))
            spec.caches[m][key] = recs[o]
    return wl


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "c2"
    sort_reads = kind.endswith("+sorted")
    kind = kind.replace("+sorted", "")
    shapes = sys.argv[2:] or ["42", "52", "62", "44", "54", "81"]
    steps = int(os.environ.get("SWEEP_STEPS", "15"))
    results = {}
    for shape in shapes:
        os.environ.pop("GAML_B200_NO_TERM_TABLE", None)
        os.environ.pop("GAML_B200_NO_PERMUTE", None)
        os.environ.pop("GAML_B200_TIER1_SHAPE", None)
        os.environ.pop("GAML_B200_STREAM_SHAPE", None)
        if shape.startswith("t"):   # two launches: tier 1 alone in this shape, then the rest
            os.environ["GAML_B200_TIER1_SHAPE"] = shape[1:]
            os.environ["GAML_B200_STREAM_SHAPE"] = "2"
        elif shape == "notab":
            os.environ["GAML_B200_NO_TERM_TABLE"] = "1"
            os.environ["GAML_B200_STREAM_SHAPE"] = "0"
        elif shape.endswith("np"):
            os.environ["GAML_B200_NO_PERMUTE"] = "1"
            os.environ["GAML_B200_STREAM_SHAPE"] = shape[:-2]
        else:
            os.environ["GAML_B200_STREAM_SHAPE"] = shape
        if "wl" not in results:
            t0 = time.time()
            results["wl"] = bench.make_workload(1, 0, int(os.environ.get("SWEEP_EVALS", "202")), kind=kind)
            if sort_reads:
                relabel_by_key(results["wl"][0])
            print(f"workload {kind} generated in {time.time() - t0:.1f}s", flush=True)
        wl, shard = results["wl"]
        pc, t_up, nbytes = bench.load_calculator(wl, shard, 0)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        sink = torch.zeros((), dtype=torch.int64, device="cuda")

        def step():
            pc.reset_state()
            pc.prepare(wl.evals[0])
            flush.fill_(1)
            sink.copy_(flush.view(torch.int32).sum())
            torch.cuda.synchronize()
            pc.launch()
            part, tl = pc.finish()
            st = pc.stats()
            return st.last_device_ms, st.last_score_kernel_ms, part

        pc.set_profiling(1)
        for _ in range(3):
            step()
        dev = sorted(step()[0] for _ in range(steps))
        pc.set_profiling(2)
        step()
        ker = sorted(step()[1] for _ in range(steps))
        st = pc.stats()
        pc.set_profiling(3)
        step(); step()
        tl = {k: [round(a, 1), round(b, 1)] for k, (a, b) in pc.read_timeline().items()}
        pc.set_profiling(0)
        part = step()[2]
        bytes_ = st.last_algorithmic_bytes
        print(f"shape {shape}: kernel med {ker[len(ker)//2]*1e3:.1f} us (min {ker[0]*1e3:.1f}) -> {bytes_/ker[len(ker)//2]/1e6:.0f} GB/s = "
              f"{bytes_/ker[len(ker)//2]/1e6/6547.2:.3f}; eval med {dev[len(dev)//2]*1e3:.1f} us; partial {part[:3]}; timeline {tl}", flush=True)
        pc.close()
        del flush


if __name__ == "__main__":
    main()
