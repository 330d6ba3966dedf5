"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares, not absolutes)."""
import collections
import csv
import re
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.2f ms summed kernel time (cold-cache, serialised under ncu)" % (
        path, sum(v[0] for v in agg.values()), tot / 1e6))
    print("# %10s %7s %7s  kernel" % ("ms", "share", "calls"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%12.3f %6.1f%% %7d  %s" % (v[1] / 1e6, 100 * v[1] / tot, v[0], k[:120]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
