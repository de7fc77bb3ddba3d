#!/usr/bin/env python
"""Per-kernel share of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        d[r[ki].split("(")[0]].append(v)
    tot = sum(sum(v) for v in d.values())
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / tot:.1f} % |")
    print(f"\ntotal {tot:.1f} us over {sum(len(v) for v in d.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
