#!/usr/bin/env python
"""Per-CUDA-source-line stall samples of one kernel in an ncu report (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py rep.ncu-rep kernel_regex [launch_skip] [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    files, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit() and len(r) > 6 and r[6].isdigit():
            files.append((cur, r))
    i_s = hdr.index("# Samples")
    i_e = hdr.index("Instructions Executed")
    stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for _, r in files)
    print(f"# total samples {tot}")
    for f, r in sorted(files, key=lambda fr: -int(fr[1][i_s]))[:top]:
        st = sorted(((int(r[i]), n) for i, n in stall), reverse=True)[:2]
        print(f"{100.0 * int(r[i_s]) / tot:5.1f}% {f.split('/')[-1]}:{r[0]:>4s} ex={r[i_e]:>10s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} | {r[1].strip()[:110]}")


if __name__ == "__main__":
    main()
