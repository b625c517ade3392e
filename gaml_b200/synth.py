"""Synthetic genomes, reads and injected alignment caches (BASELINE.json configs C1..C5).

Nothing here is on the scoring path: it manufactures the INPUT of the hot path — what the
reference's internal aligner (graph.cc:839-899) would have left in `aligment_cache_` — for a genome
made of unique nodes plus a few repeat nodes, and a scripted sequence of walk sets that mimics the
moves of the annealing loop (extend / disconnect / interchange, gaml.cc:173-211).

Cache keys follow the reference's lookup rules exactly:
  single / paired: window key per node (node i plus following nodes until their summed length
                   exceeds kMinSubpathLength=300, graph.cc:552-561, 618-627) and, for paired sets,
                   the single-node key when the node is longer than 300 (graph.cc:563-566);
  window content : what AlignSubpathInternal keeps — a long first (last) node of a multi-node window
                   is cropped to its last (first) 300 bp (graph.cc:849-853); positions are 1-based
                   and relative to the un-cropped start of the first node (graph.cc:890);
  pacbio         : every prefix window (i..j) until the part beyond node i exceeds the longest read
                   (graph.cc:2438-2454).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .workload import (ALN_DTYPE, KIND_PACBIO, KIND_PAIRED, KIND_SINGLE, PB_DTYPE, Key, ReadSetSpec,
                       Workload)

K_MIN_SUBPATH = 300  # kMinSubpathLength, graph.cc:27


# ----------------------------------------------------------------------------------------------
# genome
# ----------------------------------------------------------------------------------------------
@dataclass
class Genome:
    node_len: np.ndarray       # int32 [2*n]: node 2k and its twin 2k+1 share a length
    units: np.ndarray          # int32 [U]: forward node ids along the genome
    starts: np.ndarray         # int64 [U+1]: genome coordinate of each unit (+ total length)
    unique_units: np.ndarray   # indices into `units` that are unique (long) nodes

    @property
    def length(self) -> int:
        return int(self.starts[-1])

    def strand(self, s: int):
        """Unit ids and start coordinates of strand s (0 forward, 1 reverse complement)."""
        if s == 0:
            return self.units, self.starts
        u = (self.units[::-1] ^ 1).astype(np.int32)
        lens = self.node_len[u].astype(np.int64)
        st = np.concatenate([[0], np.cumsum(lens)])
        return u, st


def make_genome(n_unique: int, unique_len: int, jitter: float = 0.2, n_repeat_nodes: int = 3,
                repeat_len: int = 400, repeat_copies: int = 2, seed: int = 7) -> Genome:
    rng = np.random.default_rng(seed)
    ulen = (unique_len * (1.0 + jitter * (2 * rng.random(n_unique) - 1))).astype(np.int32)
    n_nodes = n_unique + n_repeat_nodes
    node_len = np.zeros(2 * n_nodes, dtype=np.int32)
    node_len[0:2 * n_unique:2] = ulen
    node_len[1:2 * n_unique:2] = ulen
    node_len[2 * n_unique::2] = repeat_len
    node_len[2 * n_unique + 1::2] = repeat_len
    units: List[int] = [2 * k for k in range(n_unique)]
    uniq_flag = [True] * n_unique
    # drop each repeat node `repeat_copies` times at distinct interior boundaries
    n_slots = n_unique - 1
    if n_repeat_nodes and n_slots >= n_repeat_nodes * repeat_copies:
        slots = rng.choice(n_slots, size=n_repeat_nodes * repeat_copies, replace=False)
        ins = sorted(((int(b) + 1, 2 * (n_unique + i // repeat_copies)) for i, b in enumerate(slots)),
                     reverse=True)
        for pos, nid in ins:
            units.insert(pos, nid)
            uniq_flag.insert(pos, False)
    units_a = np.asarray(units, dtype=np.int32)
    starts = np.concatenate([[0], np.cumsum(node_len[units_a].astype(np.int64))])
    return Genome(node_len, units_a, starts, np.nonzero(np.asarray(uniq_flag))[0])


# ----------------------------------------------------------------------------------------------
# alignments in genome coordinates
# ----------------------------------------------------------------------------------------------
@dataclass
class Alns:
    """Short-read alignments of ONE mate in forward-genome coordinates (0-based start)."""
    a: np.ndarray        # int64 start
    l: np.ndarray        # int32 aligned length (= read length here)
    read: np.ndarray     # int32 read id
    ed: np.ndarray       # int32 edit distance
    orient: np.ndarray   # int32 0 forward / 1 reverse

    def on_strand(self, s: int, glen: int) -> "Alns":
        if s == 0:
            out = self
        else:
            out = Alns(glen - self.a - self.l, self.l, self.read, self.ed, self.orient ^ 1)
        o = np.argsort(out.a, kind="stable")
        return Alns(out.a[o], out.l[o], out.read[o], out.ed[o], out.orient[o])


def _concat(xs: Sequence[Alns]) -> Alns:
    return Alns(*[np.concatenate([getattr(x, f) for x in xs]) for f in ("a", "l", "read", "ed", "orient")])


def make_single_reads(g: Genome, n_reads: int, read_len: int = 100, extra_frac: float = 0.1, seed: int = 11,
                      read_lo: int = 0, read_hi: Optional[int] = None, chunk: int = 1 << 20):
    """Uniform single-end reads; returns (read_len array for [read_lo,read_hi), Alns)."""
    read_hi = n_reads if read_hi is None else read_hi
    parts = []
    for c0 in range((read_lo // chunk) * chunk, read_hi, chunk):
        rng = np.random.default_rng([seed, c0 // chunk])
        n = min(chunk, n_reads - c0)
        a = rng.integers(0, g.length - read_len, size=n)
        orient = rng.integers(0, 2, size=n).astype(np.int32)
        ed = rng.integers(0, 3, size=n).astype(np.int32)
        has_x = rng.random(n) < extra_frac
        xa = rng.integers(0, g.length - read_len, size=n)
        xed = rng.integers(3, 6, size=n).astype(np.int32)
        xo = rng.integers(0, 2, size=n).astype(np.int32)
        ids = np.arange(c0, c0 + n, dtype=np.int32)
        keep = (ids >= read_lo) & (ids < read_hi)
        l = np.full(n, read_len, dtype=np.int32)
        parts.append(Alns(a[keep], l[keep], ids[keep], ed[keep], orient[keep]))
        kx = keep & has_x
        parts.append(Alns(xa[kx], l[kx], ids[kx], xed[kx], xo[kx]))
    return np.full(read_hi - read_lo, read_len, dtype=np.int32), _concat(parts)


def make_paired_reads(g: Genome, n_pairs: int, read_len: int = 100, insert_mean: float = 300.0,
                      insert_std: float = 30.0, min_insert: int = 200, extra_frac: float = 0.1, seed: int = 13,
                      read_lo: int = 0, read_hi: Optional[int] = None, chunk: int = 1 << 20):
    """Innie pairs: returns (len1, len2, Alns mate1, Alns mate2) for pairs [read_lo, read_hi).

    Pair from the forward strand: mate 1 forward at s, mate 2 reverse at s+f-len (so the reference's
    dist = y.pos - x.pos + len2 = f, graph.cc:1866-1870); from the reverse strand the roles swap
    (graph.cc:1871-1875). Generated per 1Mi-pair chunk from (seed, chunk) so any read-id shard can be
    produced independently on its own rank.
    """
    read_hi = n_pairs if read_hi is None else read_hi
    p1, p2 = [], []
    for c0 in range((read_lo // chunk) * chunk, read_hi, chunk):
        rng = np.random.default_rng([seed, c0 // chunk])
        n = min(chunk, n_pairs - c0)
        f = np.maximum(np.rint(rng.normal(insert_mean, insert_std, size=n)).astype(np.int64), min_insert)
        s = (rng.random(n) * (g.length - f)).astype(np.int64)
        strand = rng.integers(0, 2, size=n).astype(np.int32)
        ed1 = rng.integers(0, 3, size=n).astype(np.int32)
        ed2 = rng.integers(0, 3, size=n).astype(np.int32)
        a1 = np.where(strand == 0, s, s + f - read_len)
        a2 = np.where(strand == 0, s + f - read_len, s)
        x1 = rng.random(n) < extra_frac
        x2 = rng.random(n) < extra_frac
        xa1 = rng.integers(0, g.length - read_len, size=n)
        xa2 = rng.integers(0, g.length - read_len, size=n)
        xe1 = rng.integers(3, 6, size=n).astype(np.int32)
        xe2 = rng.integers(3, 6, size=n).astype(np.int32)
        xo1 = rng.integers(0, 2, size=n).astype(np.int32)
        xo2 = rng.integers(0, 2, size=n).astype(np.int32)
        ids = np.arange(c0, c0 + n, dtype=np.int32)
        keep = (ids >= read_lo) & (ids < read_hi)
        l = np.full(n, read_len, dtype=np.int32)
        p1.append(Alns(a1[keep], l[keep], ids[keep], ed1[keep], strand[keep]))
        p2.append(Alns(a2[keep], l[keep], ids[keep], ed2[keep], (strand ^ 1)[keep]))
        k1, k2 = keep & x1, keep & x2
        p1.append(Alns(xa1[k1], l[k1], ids[k1], xe1[k1], xo1[k1]))
        p2.append(Alns(xa2[k2], l[k2], ids[k2], xe2[k2], xo2[k2]))
    n_loc = read_hi - read_lo
    ln = np.full(n_loc, read_len, dtype=np.int32)
    return ln, ln.copy(), _concat(p1), _concat(p2)


# ----------------------------------------------------------------------------------------------
# key enumeration (the reference's lookup rules)
# ----------------------------------------------------------------------------------------------
def split_contigs(walk: Sequence[int]) -> List[List[int]]:
    ctgs: List[List[int]] = [[]]
    for x in walk:
        if x < 0:
            ctgs.append([])
        else:
            ctgs[-1].append(int(x))
    return ctgs


def window_key(ctg: Sequence[int], i: int, node_len: np.ndarray) -> Key:
    """graph.cc:552-561 / 618-627."""
    key = [ctg[i]]
    tot = 0
    for j in range(i + 1, len(ctg)):
        tot += int(node_len[ctg[j]])
        key.append(ctg[j])
        if tot > K_MIN_SUBPATH:
            break
    return tuple(key)


def short_keys_for_walks(walks: Iterable[Sequence[int]], node_len: np.ndarray, with_single_node: bool) -> List[Key]:
    keys: Dict[Key, None] = {}
    for w in walks:
        for ctg in split_contigs(w):
            for i in range(len(ctg)):
                keys[window_key(ctg, i, node_len)] = None
                if with_single_node and node_len[ctg[i]] > K_MIN_SUBPATH:
                    keys[(ctg[i],)] = None
    return list(keys)


def pacbio_keys_for_walks(walks: Iterable[Sequence[int]], node_len: np.ndarray, nmap: np.ndarray,
                          max_read_len: int) -> List[Key]:
    """graph.cc:2438-2454 (after NormalizePath, graph.h:268-273); gaps stay inside the walk."""
    keys: Dict[Key, None] = {}
    for w in walks:
        p = [int(nmap[x]) if x >= 0 else int(x) for x in w]
        el = [int(node_len[x]) if x >= 0 else -x for x in p]
        for i in range(len(p)):
            beyond = 0
            for j in range(i, len(p)):
                if j > i:
                    beyond += el[j]
                keys[tuple(p[i:j + 1])] = None
                if beyond > max_read_len:
                    break
    return list(keys)


# ----------------------------------------------------------------------------------------------
# cache construction
# ----------------------------------------------------------------------------------------------
def _find_subseq(units: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Start indices where `key` occurs contiguously in `units`."""
    k = len(key)
    if k > len(units):
        return np.zeros(0, dtype=np.int64)
    ok = units[:len(units) - k + 1] == key[0]
    for d in range(1, k):
        ok &= units[d:len(units) - k + 1 + d] == key[d]
    return np.nonzero(ok)[0]


class ShortCacheBuilder:
    """Fills `aligment_cache_`-shaped dicts for one mate from genome-coordinate alignments."""

    def __init__(self, g: Genome, alns: Alns):
        self.g = g
        self.strands = []
        for s in (0, 1):
            u, st = g.strand(s)
            self.strands.append((u, st, alns.on_strand(s, g.length)))

    def _take(self, s: int, lo: int, hi: int, shift: int, out: list) -> None:
        """Alignments of strand s fully inside genome interval [lo,hi) -> key coordinate a - shift."""
        _, _, al = self.strands[s]
        i0 = np.searchsorted(al.a, lo, side="left")
        i1 = np.searchsorted(al.a, hi, side="left")
        if i1 <= i0:
            return
        sl = slice(i0, i1)
        m = al.a[sl] + al.l[sl] <= hi
        if not m.any():
            return
        rec = np.empty(int(m.sum()), dtype=ALN_DTYPE)
        rec["position"] = (al.a[sl][m] - shift + 1).astype(np.int32)   # 1-based, graph.cc:890
        rec["edit_dist"] = al.ed[sl][m]
        rec["read_id"] = al.read[sl][m]
        rec["orientation"] = al.orient[sl][m]
        out.append(rec)

    def records(self, key: Key) -> np.ndarray:
        nl = self.g.node_len
        k = np.asarray(key, dtype=np.int32)
        lens = nl[k].astype(np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)])
        multi = len(key) > 1
        # cropped region of each node in key coordinates (graph.cc:848-856)
        reg = []
        for i in range(len(key)):
            lo, hi = int(offs[i]), int(offs[i + 1])
            if multi and i == 0 and lens[i] > K_MIN_SUBPATH:
                lo = hi - K_MIN_SUBPATH
            elif multi and i > 0 and i + 1 == len(key) and lens[i] > K_MIN_SUBPATH:
                hi = lo + K_MIN_SUBPATH
            reg.append((lo, hi))
        out: list = []
        for s in (0, 1):
            u, st, _ = self.strands[s]
            # whole-key occurrences: everything inside the cropped window, spanning reads included
            for p in _find_subseq(u, k):
                g0 = int(st[p])
                self._take(s, g0 + reg[0][0], g0 + reg[-1][1], g0, out)
            # every other copy of each node: reads fully inside that node's (cropped) part
            for i in range(len(key)):
                for p in np.nonzero(u == k[i])[0]:
                    g0 = int(st[p]) - int(offs[i])
                    self._take(s, g0 + reg[i][0], g0 + reg[i][1], g0, out)
        if not out:
            return np.zeros(0, dtype=ALN_DTYPE)
        rec = np.unique(np.concatenate(out))          # de-dup, as set<Aligment> does (graph.cc:841)
        o = np.lexsort((rec["read_id"], rec["position"]))   # Aligment::operator<, graph.h:227-230
        return rec[o]

    def build(self, keys: Iterable[Key]) -> Dict[Key, np.ndarray]:
        return {k: self.records(k) for k in keys}


# ----------------------------------------------------------------------------------------------
# scripted move sequences (walk sets)
# ----------------------------------------------------------------------------------------------
def invert_walk(w: Sequence[int]) -> List[int]:
    return [(x ^ 1) if x >= 0 else x for x in reversed(w)]


def start_walks(g: Genome, threshold: int = 500) -> List[List[int]]:
    """One walk per long node (gaml.cc:1002-1005)."""
    n_nodes = len(g.node_len) // 2
    return [[2 * k] for k in range(n_nodes) if g.node_len[2 * k] > threshold]


def scripted_evals(g: Genome, n_evals: int, seed: int = 5, p_true_join: float = 0.55, p_gap: float = 0.15,
                   p_reject: float = 0.35) -> List[List[List[int]]]:
    """A deterministic pseudo-annealing trajectory.

    Each step proposes one move on the current walk set and is "rejected" with probability p_reject
    (the next proposal then starts from the older set), which exercises the reference's rule that
    ScoringState follows the last EVALUATED set, not the last accepted one (graph.cc:1986).
    Moves: extend (join two walks: the true genome neighbour, through the repeat between them when there
    is one, or a random wrong partner; sometimes with a gap), disconnect (split a walk),
    interchange (swap the tails of two walks), flip (reverse-complement a walk).
    """
    rng = np.random.default_rng(seed)
    cur = start_walks(g)
    evals = [[list(w) for w in cur]]
    # genome successor of each unit occurrence, to propose true joins
    unit_pos = {}
    for idx, u in enumerate(g.units):
        unit_pos.setdefault(int(u), []).append(idx)

    def true_extension(last_node: int) -> Optional[List[int]]:
        """Nodes that follow `last_node` in the genome up to and including the next unique node."""
        fwd = last_node % 2 == 0
        base = last_node if fwd else last_node ^ 1
        if base not in unit_pos:
            return None
        idx = unit_pos[base][0]
        out = []
        step = 1 if fwd else -1
        j = idx + step
        uniq = set(int(x) for x in g.unique_units)
        while 0 <= j < len(g.units):
            n = int(g.units[j])
            out.append(n if fwd else n ^ 1)
            if j in uniq:
                return out
            j += step
        return None

    while len(evals) < n_evals:
        new = [list(w) for w in cur]
        r = rng.random()
        if r < 0.5 and len(new) >= 2:                      # extend
            i = int(rng.integers(len(new)))
            ext = true_extension(new[i][-1]) if rng.random() < p_true_join else None
            j = -1
            if ext is not None:
                tgt = ext[-1]
                for jj, w in enumerate(new):
                    if jj != i and w[0] == tgt:
                        j = jj
                        break
                    if jj != i and (w[-1] ^ 1) == tgt:
                        new[jj] = invert_walk(w)
                        j = jj
                        break
            if j >= 0:
                mid = ext[:-1]
            else:
                j = int(rng.integers(len(new) - 1))
                j += j >= i
                mid = []
            if rng.random() < p_gap:
                mid = [-int(rng.integers(1, 400))]
            joined = new[i] + mid + new[j]
            new = [w for t, w in enumerate(new) if t not in (i, j)] + [joined]
        elif r < 0.75:                                     # disconnect
            cand = [t for t, w in enumerate(new) if sum(1 for x in w if x >= 0) >= 2]
            if not cand:
                continue
            i = cand[int(rng.integers(len(cand)))]
            w = new[i]
            cut = int(rng.integers(1, len(w)))
            a, b = w[:cut], w[cut:]
            while a and a[-1] < 0:
                a.pop()
            while b and b[0] < 0:
                b.pop(0)
            if not a or not b:
                continue
            new = [x for t, x in enumerate(new) if t != i] + [a, b]
        elif r < 0.9 and len(new) >= 2:                    # interchange tails
            i, j = (int(x) for x in rng.choice(len(new), size=2, replace=False))
            wi, wj = new[i], new[j]
            ci = int(rng.integers(1, len(wi) + 1))
            cj = int(rng.integers(1, len(wj) + 1))
            a, b = wi[:ci] + wj[cj:], wj[:cj] + wi[ci:]
            if any(x and (x[0] < 0 or x[-1] < 0) for x in (a, b)) or not a or not b:
                continue
            new[i], new[j] = a, b
        else:                                              # flip
            i = int(rng.integers(len(new)))
            new[i] = invert_walk(new[i])
        evals.append(new)
        if rng.random() >= p_reject:
            cur = new
    return evals


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def all_walks(evals) -> List[List[int]]:
    seen = {}
    for ws in evals:
        for w in ws:
            seen[tuple(w)] = None
    return [list(w) for w in seen]


def paired_workload(n_unique: int, unique_len: int, n_pairs: int, n_evals: int = 8, seed: int = 1,
                    read_len: int = 100, insert_mean: float = 300.0, insert_std: float = 30.0,
                    n_repeat_nodes: int = 3, read_lo: int = 0, read_hi: Optional[int] = None,
                    evals=None, extra_frac: float = 0.1) -> Workload:
    g = make_genome(n_unique, unique_len, n_repeat_nodes=n_repeat_nodes, seed=seed)
    if evals is None:
        evals = scripted_evals(g, n_evals, seed=seed + 100)
    l1, l2, a1, a2 = make_paired_reads(g, n_pairs, read_len, insert_mean, insert_std, seed=seed + 200,
                                       read_lo=read_lo, read_hi=read_hi, extra_frac=extra_frac)
    keys = short_keys_for_walks(all_walks(evals), g.node_len, with_single_node=True)
    c1 = ShortCacheBuilder(g, a1).build(keys)
    c2 = ShortCacheBuilder(g, a2).build(keys)
    rs = ReadSetSpec(kind=KIND_PAIRED, n_reads=n_pairs, read_len=[l1, l2], caches=[c1, c2],
                     insert_mean=insert_mean, insert_std=insert_std, step=insert_mean - 50.0, name="paired")
    return Workload(node_len=g.node_len, normalize_map=np.arange(len(g.node_len), dtype=np.int32), sets=[rs],
                    evals=evals, meta={"genome_len": g.length, "read_lo": read_lo,
                                       "read_hi": n_pairs if read_hi is None else read_hi})


def single_workload(n_unique: int, unique_len: int, n_reads: int, n_evals: int = 8, seed: int = 2,
                    read_len: int = 100, n_repeat_nodes: int = 3, evals=None) -> Workload:
    g = make_genome(n_unique, unique_len, n_repeat_nodes=n_repeat_nodes, seed=seed)
    if evals is None:
        evals = scripted_evals(g, n_evals, seed=seed + 100)
    ln, al = make_single_reads(g, n_reads, read_len, seed=seed + 200)
    keys = short_keys_for_walks(all_walks(evals), g.node_len, with_single_node=False)
    cache = ShortCacheBuilder(g, al).build(keys)
    rs = ReadSetSpec(kind=KIND_SINGLE, n_reads=n_reads, read_len=[ln], caches=[cache], name="single")
    return Workload(node_len=g.node_len, normalize_map=np.arange(len(g.node_len), dtype=np.int32), sets=[rs],
                    evals=evals, meta={"genome_len": g.length})


def pacbio_readset(g: Genome, evals, n_reads: int, read_len: int = 10000, seed: int = 17,
                   weight: float = 0.5) -> ReadSetSpec:
    """PacBio-like reads: logprob ~ -1500 - U[0,1000), every third read gets a second alignment
    (logprob - 2.5) on the other copy of the window it sits in."""
    rng = np.random.default_rng(seed)
    nmap = np.arange(len(g.node_len), dtype=np.int32)
    lens = np.maximum((read_len * (1.0 + 0.3 * (2 * rng.random(n_reads) - 1))).astype(np.int32), 500)
    lens = np.minimum(lens, g.length // 2)
    max_len = int(lens.max())
    a = (rng.random(n_reads) * (g.length - lens)).astype(np.int64)
    strand = rng.integers(0, 2, size=n_reads)
    lp = -1500.0 - 1000.0 * rng.random(n_reads)
    keys = pacbio_keys_for_walks(all_walks(evals), g.node_len, nmap, max_len)
    cache: Dict[Key, list] = {k: [] for k in keys}
    keyset = set(keys)
    for s in (0, 1):
        u, st = g.strand(s)
        sel = np.nonzero(strand == s)[0]
        aa = a[sel] if s == 0 else g.length - a[sel] - lens[sel]
        bb = aa + lens[sel]
        ui = np.searchsorted(st, aa, side="right") - 1          # unit holding the first base
        uj = np.searchsorted(st, bb - 1, side="right") - 1      # unit holding the last base
        for t in range(len(sel)):
            key = tuple(int(x) for x in u[ui[t]:uj[t] + 1])
            if key not in keyset:
                continue
            rid = int(sel[t])
            base = int(st[ui[t]])
            cache[key].append((int(aa[t] - base), int(bb[t] - base), rid, 0, float(lp[rid])))
            if rid % 3 == 0:
                cache[key].append((int(aa[t] - base) + 3, int(bb[t] - base) + 3, rid, 0, float(lp[rid]) - 2.5))
    out: Dict[Key, np.ndarray] = {}
    for k, v in cache.items():
        rec = np.array(v, dtype=PB_DTYPE) if v else np.zeros(0, dtype=PB_DTYPE)
        out[k] = rec[np.argsort(rec["position"], kind="stable")]     # PacbioAligment::operator<, graph.h:532
    return ReadSetSpec(kind=KIND_PACBIO, n_reads=n_reads, read_len=[lens], caches=[out], weight=weight,
                       name="pacbio")


def mixed_workload(n_unique: int, unique_len: int, n_pairs: int, n_pacbio: int, n_single: int = 0,
                   n_evals: int = 8, seed: int = 3, pacbio_len: int = 10000) -> Workload:
    """C3-style: paired (+ optional single) + PacBio read sets over one genome, weights 1.0 / 0.5."""
    g = make_genome(n_unique, unique_len, seed=seed)
    evals = scripted_evals(g, n_evals, seed=seed + 100)
    wl = paired_workload(n_unique, unique_len, n_pairs, seed=seed, evals=evals)
    if n_single:
        ln, al = make_single_reads(g, n_single, 100, seed=seed + 300)
        keys = short_keys_for_walks(all_walks(evals), g.node_len, with_single_node=False)
        wl.sets.insert(0, ReadSetSpec(kind=KIND_SINGLE, n_reads=n_single, read_len=[ln],
                                      caches=[ShortCacheBuilder(g, al).build(keys)], name="single"))
    if n_pacbio:
        wl.sets.append(pacbio_readset(g, evals, n_pacbio, read_len=pacbio_len, seed=seed + 400))
    return wl
