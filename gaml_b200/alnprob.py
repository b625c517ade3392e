"""PacBio alignment probability (PacbioReadSet::AligmentProbability, graph.cc:2175-2297): synthetic alignments, the
GAMLAP1 file format shared with oracle/ref_harness.cc and oracle/gaml_oracle.cc (--alnprob), and the ctypes call of
gaml_pacbio_alignment_logprob (the device kernel).

GAMLAP1: "GAMLAP1\\0", f64 match, f64 mismatch, i32 band, i32 n; per alignment i32 posstart, i32 |s1|, s1, i32 |s2|, s2,
i32 n_ops, n_ops x (i32 length, i32 op character). Result file: n x f64 logval, f64 seconds.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


@dataclass
class Alignment:
    s1: bytes                      # the walk's sequence (or a window of it)
    s2: bytes                      # the read
    posstart: int                  # offset of the alignment in s1
    cigar: List[Tuple[int, str]]   # (length, 'M' | 'I' | 'D'); I = read base without a walk base


def make_alignments(n: int, read_len: int = 1500, seed: int = 1, sub: float = 0.02, ins: float = 0.10, dele: float = 0.04,
                    clip: int = 12, separators: bool = False) -> List[Alignment]:
    """PacBio-like alignments: a random walk window, a read derived from it by substitutions / insertions / deletions,
    the CIGAR of that derivation, soft-clip-like leading/trailing insertion runs, some alignments hanging over the ends
    of s1, and (optionally) a contig separator inside the window."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for a in range(n):
        ln = max(int(read_len * (0.6 + 0.8 * rng.random())), 20)
        margin = int(rng.integers(0, 40))
        s1 = alphabet[rng.integers(0, 4, size=ln + 2 * margin + 40)].copy()
        if separators and a % 5 == 0:
            s1[int(rng.integers(margin, margin + ln))] = ord("\n")
        ops: List[str] = []
        read: List[int] = []
        lead = int(rng.integers(0, clip + 1)) if a % 3 else 0
        for _ in range(lead):
            ops.append("I")
            read.append(int(alphabet[rng.integers(0, 4)]))
        i = margin
        end = margin + ln
        while i < end:
            u = rng.random()
            if u < ins:
                ops.append("I")
                read.append(int(alphabet[rng.integers(0, 4)]))
            elif u < ins + dele:
                ops.append("D")
                i += 1
            else:
                ops.append("M")
                b = int(s1[i])
                if rng.random() < sub:
                    b = int(alphabet[rng.integers(0, 4)])
                read.append(b)
                i += 1
        trail = int(rng.integers(0, clip + 1)) if a % 4 else 0
        for _ in range(trail):
            ops.append("I")
            read.append(int(alphabet[rng.integers(0, 4)]))
        cigar: List[Tuple[int, str]] = []
        for o in ops:
            if cigar and cigar[-1][1] == o:
                cigar[-1] = (cigar[-1][0] + 1, o)
            else:
                cigar.append((1, o))
        posstart = margin + 1 if a % 7 else max(margin - 3, 0)   # most rows index s1[row + posstart - 1] inside the window
        if a % 11 == 0:
            s1 = s1[: margin + ln - 5]                            # alignment hangs over the end of s1
        out.append(Alignment(bytes(s1.tobytes()), bytes(bytearray(read)), posstart, cigar))
    return out


def write_alignments(path: str, alns: List[Alignment], match: float, mismatch: float, band: int) -> None:
    with open(path, "wb") as f:
        f.write(b"GAMLAP1\0")
        f.write(struct.pack("<ddii", match, mismatch, band, len(alns)))
        for a in alns:
            f.write(struct.pack("<ii", a.posstart, len(a.s1)))
            f.write(a.s1)
            f.write(struct.pack("<i", len(a.s2)))
            f.write(a.s2)
            f.write(struct.pack("<i", len(a.cigar)))
            for ln, op in a.cigar:
                f.write(struct.pack("<ii", ln, ord(op)))


def read_logvals(path: str, n: int) -> Tuple[np.ndarray, float]:
    raw = open(path, "rb").read()
    vals = np.frombuffer(raw[: 8 * n], dtype="<f8").copy()
    secs = struct.unpack("<d", raw[8 * n: 8 * n + 8])[0]
    return vals, secs


def flatten(alns: List[Alignment]):
    """The C ABI's layout: concatenated sequences and CIGAR ops with offsets."""
    s1 = np.frombuffer(b"".join(a.s1 for a in alns), dtype=np.uint8)
    s2 = np.frombuffer(b"".join(a.s2 for a in alns), dtype=np.uint8)
    s1_off = np.zeros(len(alns) + 1, dtype=np.int64)
    s2_off = np.zeros(len(alns) + 1, dtype=np.int64)
    op_off = np.zeros(len(alns) + 1, dtype=np.int64)
    np.cumsum([len(a.s1) for a in alns], out=s1_off[1:])
    np.cumsum([len(a.s2) for a in alns], out=s2_off[1:])
    np.cumsum([len(a.cigar) for a in alns], out=op_off[1:])
    op_len = np.array([ln for a in alns for ln, _ in a.cigar], dtype=np.int32)
    op_chr = np.array([ord(op) for a in alns for _, op in a.cigar], dtype=np.uint8)
    posstart = np.array([a.posstart for a in alns], dtype=np.int32)
    return s1, s1_off, s2, s2_off, posstart, op_len, op_chr, op_off
