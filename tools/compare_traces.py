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


def margins(its):
    """(new_prob - cur_prob) of every iteration, as printed (gaml.cc:330 prints %lf: 6 decimals). The accept test is
    new_prob > cur_prob (gaml.cc:286), so an iteration whose printed margin is 0 while the two runs' arithmetic differs in
    the 1e-12 range is where a trajectory could fork: they are counted and listed (SURVEY §7.4.2 ii)."""
    out = []
    for line in its:
        m = re.match(r"itnum (\d+) .* new prob (\S+) (\S+) (\S+) ", line)
        if m:
            out.append((int(m.group(1)), float(m.group(2)) - float(m.group(3)), float(m.group(3))))
    return out


def main():
    a, acc_a, oa = parse(sys.argv[1])
    b, acc_b, ob = parse(sys.argv[2])
    ma, mb = margins(a), margins(b)
    if ma:
        nz = [abs(d) for _, d, _ in ma if d != 0.0]
        ties = [it for it, d, _ in ma if d == 0.0]
        rel = min((abs(d) / max(abs(c), 1e-300) for _, d, c in ma if d != 0.0), default=float("inf"))
        print(f"decision margins new-cur ({sys.argv[1]}): {len(ma)} iterations, {len(ties)} printed as exactly 0 (proposal reproduced the "
              f"current state or was not evaluated), smallest non-zero |margin| {min(nz) if nz else float('nan'):.6f} "
              f"(relative {rel:.3e}; the two arithmetics agree to ~1e-13 relative, a fork needs a margin below 1e-11 relative)")
        if rel < 1e-11:
            print("NOISE-LEVEL MARGIN: a non-zero decision margin below 1e-11 relative — the trajectories may legitimately fork here")
        if [x[:2] for x in ma] != [x[:2] for x in mb]:
            bad = next((i for i, (x, y) in enumerate(zip(ma, mb)) if x[:2] != y[:2]), min(len(ma), len(mb)))
            print(f"decision margins differ from iteration index {bad}")
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
