#!/usr/bin/env python
"""Hottest CUDA source lines (stall samples) of the first launch of every distinct kernel in an ncu report.

    python tools/ncu_hot_lines.py rep.ncu-rep [top_n]      (needs -lineinfo + --import-source on)
"""
import csv
import io
import subprocess
import sys


def ncu(*a):
    return subprocess.run(["ncu", "-i", *a], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    ki = raw[0].index("Kernel Name")
    first = {}
    for n, r in enumerate(raw[2:]):
        first.setdefault(r[ki], n)
    for name, skip in first.items():
        out = ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip), "--launch-count", "1")
        rows = list(csv.reader(io.StringIO(out)))
        files, cur, hdr = [], None, None
        for r in rows:
            if r and r[0] == "File Path":
                cur = r[1]
            elif r and r[0] == "Line No":
                hdr = r
            elif hdr and r and r[0].isdigit() and len(r) > 6 and r[6].isdigit():
                files.append((cur, r))
        if hdr is None or not files:
            print(f"## {name[:100]}: no source page")
            continue
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = max(1, sum(int(r[i_s]) for _, r in files))
        print(f"## {name[:100]} (launch {skip}): {tot} stall samples, hottest source lines")
        for f, r in sorted(files, key=lambda fr: -int(fr[1][i_s]))[:top]:
            st = sorted(((int(r[i]), n) for i, n in stall), reverse=True)[:2]
            print(f"{100.0 * int(r[i_s]) / tot:5.1f}% {str(f).split('/')[-1]}:{r[0]:>4s} ex={r[i_e]:>10s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} | {r[1].strip()[:100]}")


if __name__ == "__main__":
    main()
