"""Seeded and hand-made workloads used by the golden fixtures and the parity tests.

Hand-made cases hit what random data rarely does: empty walk sets, reads with no alignment at all,
duplicated walks (GetChanges is a multiset diff, graph.cc:1745-1764), reads with more placements than
the in-register path holds (scratch path), the same key occurring twice in one evaluation, out-of-table
insert distances, gaps, short nodes folded by normalize_map (graph.h:247-273), ragged read lengths.
"""
from __future__ import annotations

import numpy as np

from gaml_b200 import synth
from gaml_b200.workload import (ALN_DTYPE, KIND_PACBIO, KIND_PAIRED, KIND_SINGLE, PB_DTYPE, ReadSetSpec,
                                Workload)


def _aln(rows):
    return np.array(rows, dtype=ALN_DTYPE) if rows else np.zeros(0, dtype=ALN_DTYPE)


def _pb(rows):
    return np.array(rows, dtype=PB_DTYPE) if rows else np.zeros(0, dtype=PB_DTYPE)


def handmade_paired() -> Workload:
    """4 long nodes (ids 0,2,4,6 + twins) and a 120 bp node (8/9); 12 pairs with ragged lengths."""
    node_len = np.array([2000, 2000, 1500, 1500, 1800, 1800, 900, 900, 120, 120], dtype=np.int32)
    n = 12
    l1 = np.array([100, 100, 90, 100, 75, 100, 100, 100, 60, 100, 100, 100], dtype=np.int32)
    l2 = np.array([100, 80, 100, 100, 100, 100, 50, 100, 100, 100, 100, 100], dtype=np.int32)
    c1, c2 = {}, {}
    # node 0 alone: window key == single-node key
    c1[(0,)] = _aln([(101, 0, 0, 0), (301, 1, 1, 0), (700, 2, 2, 1), (1500, 0, 3, 0), (1500, 3, 3, 0),   # duplicate pos, read 3
                     (1850, 0, 4, 0), (40, 5, 11, 0)])
    c2[(0,)] = _aln([(320, 0, 0, 1), (560, 0, 1, 1), (480, 1, 2, 0), (1730, 2, 3, 1), (900, 0, 11, 1),
                     (5, 1, 5, 1)])                                                                 # read 5: mate 2 only
    # read 6: > 8 placements per mate inside node 2 (tandem repeat) -> scratch path
    c1[(2,)] = _aln([(50 + 7 * k, k % 3, 6, 0) for k in range(11)] + [(900, 0, 7, 0), (1300, 1, 9, 1)])
    c2[(2,)] = _aln([(330 + 5 * k, (k + 1) % 3, 6, 1) for k in range(10)] + [(1150, 0, 7, 1), (1010, 2, 9, 0),
                                                                            (1400, 0, 8, 1)])
    c1[(4,)] = _aln([(100, 0, 8, 0), (1700, 0, 10, 0)])
    c2[(4,)] = _aln([(20, 0, 8, 1)])
    c1[(6,)] = _aln([])
    c2[(6,)] = _aln([(150, 1, 10, 1)])
    # join 0 -> 2: window [0,2] holds the tail of 0 and head of 2 (positions relative to node 0's start)
    c1[(0, 2)] = _aln([(1850, 0, 4, 0), (1990, 1, 5, 0), (2050, 0, 6, 0)])
    c2[(0, 2)] = _aln([(2100, 0, 4, 1), (2230, 0, 5, 1)])
    # join 4 -> 8 -> 6 through the short node: windows [4,8,6], [8,6]
    c1[(4, 8, 6)] = _aln([(1700, 0, 10, 0)])
    c2[(4, 8, 6)] = _aln([(1925, 1, 10, 1)])
    c1[(8, 6)] = _aln([])
    c2[(8, 6)] = _aln([(125, 1, 10, 1), (271, 1, 10, 1)])
    # reverse-complement walk of node 0 (twin 1)
    c1[(1,)] = _aln([(1800, 0, 0, 1), (1600, 1, 1, 1)])
    c2[(1,)] = _aln([(1581, 0, 0, 0), (1341, 0, 1, 0)])
    # far-apart pair (insert distance beyond the 5-sigma table and beyond exp underflow)
    c1[(2, 4)] = _aln([(1400, 0, 2, 0)])
    c2[(2, 4)] = _aln([(1620, 0, 2, 1)])
    rs = ReadSetSpec(kind=KIND_PAIRED, n_reads=n, read_len=[l1, l2], caches=[c1, c2], insert_mean=300.0,
                     insert_std=30.0, step=250.0)
    evals = [
        [[0], [2], [4], [6]],
        [[0], [2], [4], [6]],                 # no change
        [[0, 2], [4], [6]],                   # extend
        [[0, 2], [4, 8, 6]],                  # extend through the short node
        [[0], [2], [4, 8, 6]],                # disconnect
        [[0], [0], [2], [4, 8, 6]],           # duplicated walk
        [[0], [2], [4, 8, 6]],                # one copy removed again
        [[1], [2, 4], [6]],                   # flip + new join
        [[1], [2, -250, 4], [6]],             # same with a gap instead
        [],                                   # empty assembly (total_len 0 -> 1)
        [[2], [2], [2]],
        [[0], [2], [4], [6]],
    ]
    return Workload(node_len=node_len, normalize_map=np.arange(len(node_len), dtype=np.int32), sets=[rs], evals=evals)


def handmade_single() -> Workload:
    node_len = np.array([1200, 1200, 800, 800, 90, 90, 2500, 2500], dtype=np.int32)
    n = 9
    ln = np.array([100, 100, 80, 100, 36, 100, 100, 100, 100], dtype=np.int32)
    c = {}
    c[(0,)] = _aln([(10, 0, 0, 0), (10, 2, 0, 1), (400, 1, 1, 1), (1100, 0, 2, 0)])
    c[(2,)] = _aln([(5, 0, 3, 0), (700, 3, 3, 0)] + [(20 + 3 * k, k % 4, 5, k % 2) for k in range(13)])   # read 5: scratch path
    c[(6,)] = _aln([(2000, 0, 6, 0)])
    c[(0, 2)] = _aln([(1100, 0, 2, 0), (1150, 1, 7, 0), (1205, 0, 3, 0)])       # 1205 == node2 pos 5 shifted: de-dup
    c[(2, 4, 6)] = _aln([(790, 0, 8, 0)])
    c[(4, 6)] = _aln([(60, 2, 8, 1)])
    rs = ReadSetSpec(kind=KIND_SINGLE, n_reads=n, read_len=[ln], caches=[c])
    evals = [
        [[0], [2], [6]],
        [[0, 2], [6]],
        [[0, 2, 4, 6]],
        [[6], [0, 2]],
        [[0], [0]],
        [],
        [[2, -40, 6], [0]],
    ]
    return Workload(node_len=node_len, normalize_map=np.arange(len(node_len), dtype=np.int32), sets=[rs], evals=evals)


def handmade_pacbio() -> Workload:
    """Two 2-bp nodes with equal sequence are folded onto one id by normalize_map (graph.h:256-264)."""
    node_len = np.array([3000, 3000, 2, 2, 2, 2, 2500, 2500, 4000, 4000], dtype=np.int32)
    nmap = np.arange(10, dtype=np.int32)
    nmap[4] = 2           # node 4 has the same <=3 bp sequence as node 2
    n = 7
    ln = np.array([2500, 1800, 3000, 900, 2000, 2600, 1000], dtype=np.int32)
    c = {}
    keys = [(0,), (0, 2), (0, 2, 6), (2,), (2, 6), (2, 6, 8), (6,), (6, 8), (8,), (0, 2, 6, 8), (2, 6, 8),
            (0, -100), (0, -100, 8), (-100,), (-100, 8)]
    for k in keys:
        c[k] = _pb([])
    c[(0,)] = _pb([(100, 2600, 0, 0, -1800.25), (400, 2200, 1, 0, -1500.5), (405, 2205, 1, 0, -1503.0)])
    c[(0, 2, 6)] = _pb([(1500, 4500, 2, 0, -2400.0), (100, 2600, 0, 0, -1800.25)])
    c[(6,)] = _pb([(10, 910, 3, 0, -700.125)] + [(20 + k, 2020 + k, 4, 0, -1600.0 - 0.5 * k) for k in range(12)])
    c[(6, 8)] = _pb([(2000, 4600, 5, 0, -2100.75)])
    c[(8,)] = _pb([(50, 1050, 6, 0, -3000.0)])            # below the floor: floored read
    c[(0, -100, 8)] = _pb([(2900, 3900, 6, 0, -650.0)])
    rs = ReadSetSpec(kind=KIND_PACBIO, n_reads=n, read_len=[ln], caches=[c], weight=0.5)
    evals = [
        [[0], [6], [8]],
        [[0, 4, 6], [8]],           # node 4 normalises to 2
        [[0, 2, 6, 8]],
        [[0, -100, 8], [6]],
        [],
        [[6], [6]],
    ]
    return Workload(node_len=node_len, normalize_map=nmap, sets=[rs], evals=evals)


def fill_missing_keys(wl: Workload) -> Workload:
    """The reference runs an aligner on any looked-up key that is absent from the cache
    (graph.cc:447-493, 2445-2478); give every enumerated key at least an empty list so that the
    injected cache is the whole input."""
    for rs in wl.sets:
        walks = synth.all_walks(wl.evals)
        if rs.kind == KIND_PACBIO:
            keys = synth.pacbio_keys_for_walks(walks, wl.node_len, wl.normalize_map, int(rs.read_len[0].max()))
            empty = np.zeros(0, dtype=PB_DTYPE)
        else:
            keys = synth.short_keys_for_walks(walks, wl.node_len, with_single_node=rs.kind == KIND_PAIRED)
            empty = np.zeros(0, dtype=ALN_DTYPE)
        for cache in rs.caches:
            for k in keys:
                cache.setdefault(k, empty)
    return wl


def golden_cases():
    """name -> Workload; small enough to commit (tests/golden/*.wl)."""
    return {
        "hand_paired": fill_missing_keys(handmade_paired()),
        "hand_single": fill_missing_keys(handmade_single()),
        "hand_pacbio": fill_missing_keys(handmade_pacbio()),
        "synth_single": synth.single_workload(8, 1500, 600, n_evals=14, seed=21),
        "synth_paired": synth.paired_workload(10, 1800, 900, n_evals=24, seed=22),
        "synth_mixed": synth.mixed_workload(8, 5000, 600, 80, n_single=300, n_evals=12, seed=23, pacbio_len=4000),
        "synth_paired_penalty": _with_penalty(synth.paired_workload(12, 2500, 350, n_evals=30, seed=24)),
        "synth_pacbio_penalty": _with_pacbio_penalty(synth.mixed_workload(8, 5000, 300, 60, n_evals=10, seed=25, pacbio_len=4000)),
        "synth_paired_bigedit": paired_big_edit(8, 2500, 700, n_evals=10, seed=26),
        "synth_paired_ragged": paired_ragged(8, 2000, 700, n_evals=10, seed=27),
    }


def paired_big_edit(n_unique: int, unique_len: int, n_pairs: int, n_evals: int, seed: int) -> Workload:
    """250 bp pairs with some edit distances of 128 and more: such records do not fit the 16-byte packed pairs / 8-byte
    packed tier-2 records of the streaming kernel (7 bits of edit distance), so the set streams the plain records."""
    wl = synth.paired_workload(n_unique, unique_len, n_pairs, n_evals=n_evals, seed=seed, read_len=250, insert_mean=600.0,
                               insert_std=40.0)
    bumped = 0
    for cache in wl.sets[0].caches:
        for key in sorted(cache.keys()):
            recs = cache[key]
            if len(recs) and bumped < 40:
                recs["edit_dist"][0] = 128 + bumped      # still <= read length + 6
                bumped += 1
    assert bumped >= 10
    return wl


def paired_ragged(n_unique: int, unique_len: int, n_pairs: int, n_evals: int, seed: int) -> Workload:
    """Per-pair lengths differ: the packed records are used, lengths / pow tables / thresholds come from their arrays."""
    wl = synth.paired_workload(n_unique, unique_len, n_pairs, n_evals=n_evals, seed=seed)
    spec = wl.sets[0]
    rng = np.random.default_rng(seed)
    for m, step in ((0, 3), (1, 5)):
        lens = np.asarray(spec.read_len[m]).copy()
        lens[::step] -= rng.integers(1, 20, size=len(lens[::step])).astype(lens.dtype)
        spec.read_len[m] = lens
    return wl


def _with_pacbio_penalty(wl: Workload, step: float = 700.0) -> Workload:
    """PacBio coverage penalty (graph.cc:3197-3250): sparse long reads, a penalty step below the typical gap, and
    match / mismatch probabilities that put GetMinReadProb (graph.h:478) inside the records' logprob range so that
    some alignments do not count as coverage."""
    pb = wl.sets[-1]
    pb.penalty_constant = 0.0004
    pb.step = step
    pb.match_prob = 0.9
    # min read prob at the mean read length = -2000, the middle of the synthetic logprob range [-2500, -1500]
    pb.mismatch_prob = float(np.exp((-2000.0 / float(np.mean(pb.read_len[0])) - 0.75 * np.log(0.9)) / 0.25))
    return wl


def _with_penalty(wl: Workload) -> Workload:
    """Sparse coverage + penalty_constant: exercises the coverage-gap sweep (example.cfg:13-14 style settings)."""
    wl.sets[0].penalty_constant = 0.00007
    wl.sets[0].step = wl.sets[0].insert_mean - 30.0
    return wl


def seeded_cases():
    """Larger randomized cases for oracle-vs-reference (here) and CUDA-vs-oracle (GPU box)."""
    out = {}
    for seed in range(3):
        out[f"single_s{seed}"] = synth.single_workload(10, 2500, 3000, n_evals=25, seed=seed)
        out[f"paired_s{seed}"] = synth.paired_workload(14, 2500, 5000, n_evals=40, seed=seed)
        out[f"mixed_s{seed}"] = synth.mixed_workload(10, 6000, 3000, 300, n_single=1000, n_evals=20, seed=seed,
                                                     pacbio_len=5000)
        out[f"pacbio_penalty_s{seed}"] = _with_pacbio_penalty(
            synth.mixed_workload(10, 6000, 1500, 120 + 60 * seed, n_evals=14, seed=30 + seed, pacbio_len=5000), step=500.0 + 400.0 * seed)
    return out


def alnprob_cases():
    """name -> (alignments, match, mismatch, band) for the PacBio alignment probability (graph.cc:2175-2297): golden
    fixtures small enough to commit."""
    from gaml_b200 import alnprob
    return {
        "alnprob_band2": (alnprob.make_alignments(24, 260, seed=41), 0.85, 0.05, 2),
        "alnprob_separators_band1": (alnprob.make_alignments(16, 180, seed=42, separators=True), 0.8, 0.066, 1),
        "alnprob_band3": (alnprob.make_alignments(10, 220, seed=43, clip=30), 0.9, 0.03, 3),
    }


def alnprob_seeded():
    """Larger cases, oracle vs reference here and CUDA vs oracle on the GPU box."""
    from gaml_b200 import alnprob
    return {
        "long_band2": (alnprob.make_alignments(40, 2500, seed=51), 0.85, 0.05, 2),
        "many_short": (alnprob.make_alignments(600, 150, seed=52, separators=True), 0.85, 0.05, 2),
        "clipped": (alnprob.make_alignments(30, 400, seed=53, clip=260), 0.85, 0.05, 2),
    }
