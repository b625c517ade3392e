#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS page) into the few numbers DESIGN.md / profiles/ quote."""
import collections, csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
            'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
            'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__cycles_elapsed.max',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
    kn = hdr.index('Kernel Name')
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:85s}", [f"{r[kn][:28]}={r[i]}" for r in rows[2:]] if False else [r[i] for r in rows[1:]])

def sass(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, isrc, ismp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    data = []
    for r in rows[2:]:
        if len(r) < 10: break
        data.append(r)
    tot = sum(int(r[ia]) for r in data)
    print("kernel:", rows[0][1][:80], "| sass lines", len(data), "| warp-instructions", tot, "| samples", sum(int(r[ismp]) for r in data))
    op = collections.Counter(); smp = collections.Counter()
    for r in data:
        t = r[isrc].split()
        o = t[1] if t[0].startswith('@') else t[0]
        op[o] += int(r[ia]); smp[o] += int(r[ismp])
    print("top opcodes (executed, samples):", [(o, n, smp[o]) for o, n in op.most_common(18)])
    for r in sorted(data, key=lambda r: -int(r[ismp]))[:top]:
        print(f"{r[ismp]:>6} {r[ia]:>9}  {r[isrc][:100]}")

if __name__ == "__main__":
    raw(sys.argv[1]); sass(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
