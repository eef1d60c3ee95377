#!/usr/bin/env python
"""Per-kernel SASS evidence for the tcgen05 / TMEM / TMA claims: `cuobjdump -sass` of the in-tree library, with the
mnemonics of B200_PROFILING.md counted per kernel (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store,
LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UTMAPF = TMA L2 prefetch, SYNCS = mbarrier ops).

    python profiles/sass_summary.py [path/to/libcfm_b200.so] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "image-inpainting-and-super-resolution-using-diffusion-models-and-conditional-flow-matching_b200", "libcfm_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA", "MUFU.TANH",
             "MUFU.EX2", "LDG", "STG", "SHFL", "BAR.SYNC", "ACQBULK", "UBLKCP", "FFMA2", "FADD2", "FMUL2"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                counts[cur][mn] += 1
    names = list(counts)
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    except Exception:
        pass
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({os.path.getsize(LIB)} bytes)")
    print("# instruction counts per kernel (static SASS, sm_100a); columns with no hits are omitted per row")
    tot = collections.Counter()
    for fn, c in counts.items():
        name = demangle.get(fn, fn)
        name = re.sub(r"\((?:int|bool|unsigned)\)", "", name)      # template arguments print as casts
        name = re.sub(r"\(.*", "", name)
        hits = " ".join(f"{mn}={c[mn]}" for mn in MNEMONICS if c[mn])
        print(f"{name:60s} sass={c['_total']:6d}  {hits}")
        tot.update(c)
    print("TOTAL" + " " * 55 + f" sass={tot['_total']:6d}  " + " ".join(f"{mn}={tot[mn]}" for mn in MNEMONICS if tot[mn]))


if __name__ == "__main__":
    main()
