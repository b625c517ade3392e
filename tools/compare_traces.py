#!/usr/bin/env python
"""Diff two stdout traces of the reference's `gaml` driver (gaml.cc:330-338 prints one line per annealing
iteration). The wall-clock field is stripped; everything else — proposed moves, "accept" lines, new/cur/best
probabilities (6 decimals, the reference's own %lf), total length, walk count, floored-read counts — must match.

usage: compare_traces.py ref.log other.log  -> exit 0 if the annealing trajectories are identical
"""
import re
import sys


def parse(path):
    its, accepts, other = [], 0, []
    for line in open(path, errors="replace"):
        line = line.rstrip("\n")
        if line.startswith("itnum "):
            line = re.sub(r" time \d\d:\d\d:\d\d", "", line)
            its.append(line)
        elif line == "accept":
            accepts += 1
            its.append("accept@%d" % len(its))
        elif line.startswith(("start prob", "loc ", "clean ", "local save", "s t ")):
            other.append(line)
    return its, accepts, other


def main():
    a, acc_a, oa = parse(sys.argv[1])
    b, acc_b, ob = parse(sys.argv[2])
    n_it = sum(1 for x in a if x.startswith("itnum"))
    print(f"{sys.argv[1]}: {n_it} iterations, {acc_a} accepted; {sys.argv[2]}: "
          f"{sum(1 for x in b if x.startswith('itnum'))} iterations, {acc_b} accepted")
    ok = n_it > 0
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            print(f"first divergence at trace entry {i}:\n  ref: {x}\n  new: {y}")
            ok = False
            break
    if ok and len(a) != len(b):
        print(f"trace lengths differ: {len(a)} vs {len(b)}")
        ok = False
    if oa != ob:
        print(f"auxiliary lines differ ({len(oa)} vs {len(ob)})")
        ok = False
    print("IDENTICAL TRAJECTORY" if ok else "TRAJECTORIES DIFFER")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
