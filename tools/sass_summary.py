#!/usr/bin/env python
"""cuobjdump -sass of bbocr_b200/libbbocr.so -> counts of the tcgen05 / TMEM / TMA mnemonics per kernel (profiles/r2_sass_summary.txt).
usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA")


def main():
    so = os.path.join(ROOT, "bbocr_b200", "libbbocr.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            if op in MNEMONICS and not (op == "HMMA" and "UTC" in m.group(1)):
                per[cur][op] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    print("SASS evidence for the tcgen05 / TMEM / TMA kernels of bbocr_b200/libbbocr.so (cuobjdump -sass, sm_100a; round 2, fourth session; tools/sass_summary.py)")
    print("mnemonics: UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, "
          "SYNCS = mbarrier ops; HMMA (legacy mma.sync) must be absent")
    print()
    print("whole library: " + ", ".join(f"{k} {total[k]}" for k in sorted(total)) + ("" if total["HMMA"] == 0 else "   <-- legacy HMMA present"))
    print()
    print("per kernel:")
    rows = [(n, c) for n, c in zip(names, per.values()) if c["UTCHMMA"] or c["UTMALDG"] or c["UTMASTG"] or c["LDTM"]]
    for n, c in sorted(rows, key=lambda x: -x[1]["UTCHMMA"]):
        n = re.sub(r"\(CUtensorMap_st.*", "(...)", n)
        print("  " + n + ": " + ", ".join(f"{k} {c[k]}" for k in sorted(c)))


if __name__ == "__main__":
    main()
