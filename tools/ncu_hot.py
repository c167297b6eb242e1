#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel in an ncu report (source page), with context.

    python tools/ncu_hot.py rep.ncu-rep kernel_regex [launch_skip] [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[idx["# Samples"]].isdigit()]
    tot = sum(int(r[idx["# Samples"]]) for r in body)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print(f"# {rows[0][1][:80]}  total samples {tot}, {len(body)} instructions")
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]]))[:top]
    for i in sorted(ranked):
        r = body[i]
        n = int(r[idx["# Samples"]])
        st = sorted(((int(r[idx[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {100.0 * n / tot:5.1f}% ex={r[idx['Instructions Executed']]:>9s} {r[idx['Source']].strip()[:90]:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")


if __name__ == "__main__":
    main()
