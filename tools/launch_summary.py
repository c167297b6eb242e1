#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list:
per-kernel device time, launch count and share of the LAST training step in the log
(steps are delimited by the gather kernel, which runs exactly twice per step).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import collections
import csv
import io
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(io.StringIO("".join(lines))))


def short(name):
    n = name.split("(")[0]
    n = n.replace("void ", "").replace("mmi::", "").replace("tc::", "")
    return n[:78]


def main():
    rows = load(sys.argv[1])
    names = [r["Kernel Name"] for r in rows]
    gi = [i for i, n in enumerate(names) if "gather_l1norm" in n]
    start = gi[-2] if len(gi) >= 2 else 0
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[start:]:
        t = float(r["Metric Value"].replace(",", "")) / 1e6
        a = agg.setdefault(short(r["Kernel Name"]), [0.0, 0])
        a[0] += t
        a[1] += 1
        tot += t
    print(f"# {sys.argv[1]}: last step = launches {start}..{len(rows) - 1} ({len(rows) - start} launches), "
          f"sum of gpu__time_duration = {tot:.3f} ms (cold-cache, serialised: compare shares, not absolutes)")
    print(f"{'kernel':80s} {'ms':>9s} {'n':>5s} {'share':>7s}")
    for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{n:80s} {t:9.3f} {c:5d} {100 * t / tot:6.1f}%")


if __name__ == "__main__":
    main()
