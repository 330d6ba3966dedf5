"""Per-iteration kernel-time table of the search loop's two captured training iterations from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv python tools/search_profile.py --graphs`).
An iteration = the launches between two consecutive fused optimiser steps (mt_step_kernel); iterations that contain the
encoder stem are task 1 (end to end, batch 32 @350x350), the others task 0 (decoder on cached features, batch 64).
Usage: python tools/summarize_search_launches.py X.csv OUT.txt"""
import csv
import gzip
import re
import sys


def main():
    path, out = sys.argv[1], sys.argv[2]
    op = gzip.open if path.endswith(".gz") else open
    rows = []
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    i_grid = hdr.index("Grid Size") if "Grid Size" in hdr else None
    for r in rd:
        if len(r) <= i_val:
            continue
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        name = re.sub(r"\(.*", "", r[i_name]).replace("nasb::", "").replace("void ", "")
        rows.append((name, us))
    segs, cur = [], []
    for name, us in rows:
        cur.append((name, us))
        if name.startswith("mt_step_kernel"):
            segs.append(cur)
            cur = []
    phases = {"task0": [], "task1": []}
    for s in segs:
        is1 = any("stem" in n for n, _ in s)
        phases["task1" if is1 else "task0"].append(s)
    with open(out, "w") as f:
        for ph, ss in phases.items():
            # the first iteration of a phase carries whatever preceded it (validation, cache population): drop it; drop eager
            # warm-up duplicates by taking the median-length segments
            ss = ss[1:] if len(ss) > 1 else ss
            if not ss:
                continue
            n = len(ss)
            tot = sum(us for s in ss for _, us in s)
            f.write("## %s: %d iterations, %.3f ms of kernel time and %.0f launches per iteration (ncu: serialised, per-kernel)\n"
                    % (ph, n, tot / n / 1e3, sum(len(s) for s in ss) / n))
            by = {}
            for s in ss:
                for name, us in s:
                    e = by.setdefault(name, [0, 0.0])
                    e[0] += 1
                    e[1] += us
            f.write("# ms_per_iteration  launches_per_iteration  avg_us  share  kernel\n")
            for name, (c, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
                f.write("%9.3f %7.1f %8.2f %5.1f%%  %s\n" % (us / n / 1e3, c / n, us / c, 100 * us / tot, name))
            f.write("\n")


if __name__ == "__main__":
    main()
