#!/usr/bin/env python
"""Counts of the SASS mnemonics that prove what a kernel is built from (tcgen05 = UTCHMMA, TMA = UTMALDG / UTMASTG,
TMEM traffic = LDTM / STTM, mbarriers = SYNCS, 128-bit global access, MUFU, vector reductions) per kernel of the library.

    python tools/sass_mnemonics.py segmminterest_b200/libmmi_b200.so > profiles/rNN_sass_mnemonics.txt
"""
import collections
import re
import subprocess
import sys

PAT = [("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
       ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("LDG.128", r"\bLDG[.\w]*\.128"), ("STG.128", r"\bSTG[.\w]*\.128"),
       ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("REDG", r"\bREDG"), ("ATOMG", r"\bATOMG"), ("MUFU.EX2", r"MUFU\.EX2"),
       ("MUFU.TANH", r"MUFU\.TANH"), ("FFMA2", r"\bFFMA2"), ("SHFL", r"\bSHFL"), ("BAR.SYNC", r"\bBAR\.SYNC")]


def main():
    lib = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for l in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            cnt[cur] = collections.Counter()
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4,6}\*/", l):
            cnt[cur]["instr"] += 1
            for name, rx in PAT:
                if re.search(rx, l):
                    cnt[cur][name] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    print(f"# {lib}: cuobjdump -sass, static instruction counts per kernel (sm_100a)")
    for name, c in sorted(zip(names, cnt.values())):
        name = name.split("(")[0].replace("void ", "").replace("mmi::", "")
        print(f"{name[:84]:84s} instr={c['instr']:5d}  " + " ".join(f"{k}={c[k]}" for k, _ in PAT if c[k]))


if __name__ == "__main__":
    main()
