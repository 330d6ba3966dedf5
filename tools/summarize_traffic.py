"""Aggregate an ncu multi-metric CSV (tools/profile_r2.sh, step 2) by kernel: launches, time, DRAM bytes, achieved GB/s,
tensor-pipe instructions.  Usage: python tools/summarize_traffic.py gpurun_out/prof/traffic.csv.gz > profiles/r2_traffic_table.txt"""
import collections
import csv
import gzip
import re
import sys


def main(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(dict)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "") or 0)
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)  # -> us
        if row["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per[row["ID"]]["name"] = re.sub(r"\(.*", "", row["Kernel Name"])
        per[row["ID"]][row["Metric Name"]] = v
    agg = collections.defaultdict(lambda: collections.Counter())
    for d in per.values():
        a = agg[d["name"]]
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0)
        a["rd"] += d.get("dram__bytes_read.sum", 0)
        a["wr"] += d.get("dram__bytes_write.sum", 0)
        a["tensor"] += d.get("sm__inst_executed_pipe_tensor.sum", 0)
        a["regs"] = max(a["regs"], d.get("launch__registers_per_thread", 0))
    print("# %s -- per kernel over one eager training iteration (+ warm-up launches); time under ncu is cold-cache and serialised" % path)
    print("# %8s %6s %10s %10s %9s %12s %5s  kernel" % ("ms", "calls", "dram_rd_MB", "dram_wr_MB", "GB/s", "tensor_inst", "regs"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0
        print("%10.3f %6d %10.1f %10.1f %9.1f %12d %5d  %s" % (a["us"] / 1e3, a["n"], a["rd"] / 1e6, a["wr"] / 1e6, gbs, a["tensor"],
                                                             a["regs"], k[:100]))


if __name__ == "__main__":
    main(sys.argv[1])
