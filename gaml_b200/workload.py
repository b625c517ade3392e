"""Flat workload / result files shared by the reference harness, the oracle and the CUDA path.

A workload is everything `ProbCalculator::CalcProb` (reference prob_calculator.h:63-109) consumes:
node lengths (graph.h:74-77), read sets with their injected alignment caches
(`aligment_cache_`, graph.h:427 / 587), and an ordered list of walk sets ("evals") that are
scored one after another through ONE stateful calculator, as gaml.cc:105,284 does.

GAMLWL1 layout (little endian):
  char[8]  "GAMLWL1\\0"
  i32 n_nodes; i32 node_len[n_nodes]; i32 normalize_map[n_nodes]     (ids: 2k forward, 2k+1 twin)
  i32 n_sets; per set:
      i32 kind (0 single, 1 paired, 2 pacbio)
      f64 mismatch_prob, match_prob, insert_mean, insert_std, min_prob_per_base,
          min_prob_start, weight, penalty_constant, step
      i32 n_reads; i32 n_mates; per mate: i32 read_len[n_reads]
      per mate: i32 n_keys; per key: i32 key_len; i32 key[key_len]; i32 n_rec; records
          kind 0/1 record: i32 position, edit_dist, read_id, orientation         (graph.h:211-215)
          kind 2   record: i32 position, position_end, read_id, pad; f64 logprob  (graph.h:516-520)
  i32 n_evals; per eval: i32 n_walks; per walk: i32 len; i32 ids[len]

GAMLRS1 layout:
  char[8] "GAMLRS1\\0"; i32 n_evals; i32 n_sets; i32 dump
  per eval: f64 score; i32 total_len; i32 pad; f64 seconds; per set: i32 zero_reads, i32 n_reads
            if dump: per set: i32 n; i32 pad; f64 per_read[n]
               (single: sum of p1; paired: ScoringState::probs; pacbio: LSE logval)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

KIND_SINGLE, KIND_PAIRED, KIND_PACBIO = 0, 1, 2

ALN_DTYPE = np.dtype([("position", "<i4"), ("edit_dist", "<i4"), ("read_id", "<i4"), ("orientation", "<i4")])
PB_DTYPE = np.dtype([("position", "<i4"), ("position_end", "<i4"), ("read_id", "<i4"), ("pad", "<i4"),
                     ("logprob", "<f8")])
assert ALN_DTYPE.itemsize == 16 and PB_DTYPE.itemsize == 24

Key = Tuple[int, ...]


@dataclass
class ReadSetSpec:
    kind: int
    n_reads: int
    read_len: List[np.ndarray]                 # one int32 array per mate
    caches: List[Dict[Key, np.ndarray]]        # one dict per mate: key -> ALN_DTYPE / PB_DTYPE records
    mismatch_prob: float = 0.01
    match_prob: float = 0.96                   # 1 - 4*mismatch (gaml.cc:813,854)
    insert_mean: float = 0.0
    insert_std: float = 1.0
    min_prob_per_base: float = -0.7
    min_prob_start: float = -10.0
    weight: float = 1.0
    penalty_constant: float = 0.0
    step: float = 50.0
    name: str = ""

    @property
    def n_mates(self) -> int:
        return len(self.read_len)

    def n_records(self) -> int:
        return int(sum(len(v) for c in self.caches for v in c.values()))


@dataclass
class Workload:
    node_len: np.ndarray                        # int32 [n_nodes] (n_nodes even: node 2k and twin 2k+1)
    normalize_map: np.ndarray                   # int32 [n_nodes]  (graph.h:247-266)
    sets: List[ReadSetSpec] = field(default_factory=list)
    evals: List[List[List[int]]] = field(default_factory=list)   # eval -> walks -> node ids (neg = gap)
    meta: dict = field(default_factory=dict)


def _w_i32(f, *vals):
    f.write(struct.pack("<%di" % len(vals), *vals))


def write_workload(path: str, wl: Workload) -> None:
    with open(path, "wb") as f:
        f.write(b"GAMLWL1\0")
        _w_i32(f, len(wl.node_len))
        np.asarray(wl.node_len, dtype="<i4").tofile(f)
        np.asarray(wl.normalize_map, dtype="<i4").tofile(f)
        _w_i32(f, len(wl.sets))
        for s in wl.sets:
            _w_i32(f, s.kind)
            f.write(struct.pack("<9d", s.mismatch_prob, s.match_prob, s.insert_mean, s.insert_std,
                                s.min_prob_per_base, s.min_prob_start, s.weight, s.penalty_constant, s.step))
            _w_i32(f, s.n_reads, s.n_mates)
            for rl in s.read_len:
                assert len(rl) == s.n_reads
                np.asarray(rl, dtype="<i4").tofile(f)
            dt = PB_DTYPE if s.kind == KIND_PACBIO else ALN_DTYPE
            for cache in s.caches:
                _w_i32(f, len(cache))
                for key, recs in cache.items():
                    _w_i32(f, len(key), *key)
                    _w_i32(f, len(recs))
                    np.asarray(recs, dtype=dt).tofile(f)
        _w_i32(f, len(wl.evals))
        for walks in wl.evals:
            _w_i32(f, len(walks))
            for w in walks:
                _w_i32(f, len(w), *[int(x) for x in w])


def read_workload(path: str) -> Workload:
    buf = np.fromfile(path, dtype=np.uint8)
    assert bytes(buf[:8]) == b"GAMLWL1\0", "bad workload magic"
    off = 8

    def i32(n=1):
        nonlocal off
        v = np.frombuffer(buf, dtype="<i4", count=n, offset=off)
        off += 4 * n
        return v

    def f64(n=1):
        nonlocal off
        v = np.frombuffer(buf, dtype="<f8", count=n, offset=off)
        off += 8 * n
        return v

    n_nodes = int(i32()[0])
    node_len = i32(n_nodes).copy()
    nmap = i32(n_nodes).copy()
    wl = Workload(node_len=node_len, normalize_map=nmap)
    for _ in range(int(i32()[0])):
        kind = int(i32()[0])
        p = f64(9)
        n_reads, n_mates = (int(x) for x in i32(2))
        rls = [i32(n_reads).copy() for _ in range(n_mates)]
        dt = PB_DTYPE if kind == KIND_PACBIO else ALN_DTYPE
        caches = []
        for _m in range(n_mates):
            cache: Dict[Key, np.ndarray] = {}
            for _k in range(int(i32()[0])):
                kl = int(i32()[0])
                key = tuple(int(x) for x in i32(kl))
                nr = int(i32()[0])
                recs = np.frombuffer(buf, dtype=dt, count=nr, offset=off).copy()
                off += nr * dt.itemsize
                cache[key] = recs
            caches.append(cache)
        wl.sets.append(ReadSetSpec(kind=kind, n_reads=n_reads, read_len=rls, caches=caches,
                                   mismatch_prob=p[0], match_prob=p[1], insert_mean=p[2], insert_std=p[3],
                                   min_prob_per_base=p[4], min_prob_start=p[5], weight=p[6],
                                   penalty_constant=p[7], step=p[8]))
    for _ in range(int(i32()[0])):
        walks = []
        for _w in range(int(i32()[0])):
            ln = int(i32()[0])
            walks.append([int(x) for x in i32(ln)])
        wl.evals.append(walks)
    return wl


@dataclass
class EvalResult:
    score: float
    total_len: int
    seconds: float
    zeros: List[Tuple[int, int]]
    per_read: List[np.ndarray] = field(default_factory=list)


def read_results(path: str) -> List[EvalResult]:
    buf = np.fromfile(path, dtype=np.uint8)
    assert bytes(buf[:8]) == b"GAMLRS1\0", "bad result magic"
    off = 8
    n_evals, n_sets, dump = (int(x) for x in np.frombuffer(buf, "<i4", 3, off))
    off += 12
    out = []
    for _ in range(n_evals):
        score = float(np.frombuffer(buf, "<f8", 1, off)[0]); off += 8
        total_len = int(np.frombuffer(buf, "<i4", 1, off)[0]); off += 8
        secs = float(np.frombuffer(buf, "<f8", 1, off)[0]); off += 8
        z = np.frombuffer(buf, "<i4", 2 * n_sets, off); off += 8 * n_sets
        er = EvalResult(score, total_len, secs, [(int(z[2 * i]), int(z[2 * i + 1])) for i in range(n_sets)])
        if dump:
            for _s in range(n_sets):
                n = int(np.frombuffer(buf, "<i4", 1, off)[0]); off += 8
                er.per_read.append(np.frombuffer(buf, "<f8", n, off).copy()); off += 8 * n
        out.append(er)
    return out


def write_results(path: str, results: List[EvalResult], n_sets: int, dump: bool) -> None:
    with open(path, "wb") as f:
        f.write(b"GAMLRS1\0")
        _w_i32(f, len(results), n_sets, 1 if dump else 0)
        for r in results:
            f.write(struct.pack("<d", r.score))
            _w_i32(f, r.total_len, 0)
            f.write(struct.pack("<d", r.seconds))
            for z in r.zeros:
                _w_i32(f, z[0], z[1])
            if dump:
                for pr in r.per_read:
                    _w_i32(f, len(pr), 0)
                    np.asarray(pr, dtype="<f8").tofile(f)
