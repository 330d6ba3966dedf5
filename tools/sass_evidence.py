"""Per-kernel SASS opcode counts of libnasb200.so (the Blackwell-native claim made checkable): tcgen05 MMA (UTCHMMA), TMA
loads / stores (UTMALDG / UTMASTG), TMEM loads (LDTM), tcgen05 commit barriers (UTCBAR), packed fp32 FMA (FFMA2), and the
architectures of every embedded cubin.  Runs without a GPU.

    python tools/sass_evidence.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nas-segm-pytorch_b200", "libnasb200.so")
OPS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "FFMA2", "HFMA2", "FFMA", "LDG", "STG", "LDS", "ATOM", "RED")


def main():
    elf = subprocess.run(["cuobjdump", "--list-elf", LIB], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"sm_\d+a?", elf)))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + ".") or (o in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR") and op.startswith(o)):
                    counts[cur][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# %s : %d kernels, cubin architectures %s" % (os.path.relpath(LIB, ROOT), len(counts), archs))
    print("# columns: " + " ".join(OPS) + " | total instructions | kernel")
    tot = collections.Counter()
    rows = []
    for (k, c), name in zip(counts.items(), demangle):
        name = re.sub(r"\(.*", "", name)
        rows.append((c, name))
        tot.update(c)
    rows.sort(key=lambda r: (-r[0]["UTCHMMA"], -r[0]["UTMALDG"], -r[0]["FFMA2"], r[1]))
    for c, name in rows:
        print(" ".join("%6d" % c[o] for o in OPS) + " | %7d | %s" % (c["total"], name))
    print("# library totals: " + ", ".join("%s %d" % (o, tot[o]) for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
