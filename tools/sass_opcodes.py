#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a recompiled one
(/opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG /
UTMASTG / UBLKCP, mma.sync -> HMMA, cp.async -> LDGSTS, packed fp32x2 FMA -> FFMA2).

    python tools/sass_opcodes.py [vcmeshconv_b200/libvcmesh.so] > profiles/sass_opcodes_r2.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "LDSM",
       "LDGSTS", "FFMA2", "FFMA", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "MULTIMEM", "BAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vcmeshconv_b200", "libvcmesh.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, total, cur = collections.OrderedDict(), collections.Counter(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for o in OPS:
                if op == o or (o in ("UTCHMMA", "UTCQMMA") and op.startswith(o)):
                    counts[cur][o] += 1
                    total[o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# %s: %d kernels; SASS mnemonic counts per kernel (static instruction counts)" % (os.path.basename(lib), len(counts)))
    print("# totals: " + ", ".join("%s %d" % (o, total[o]) for o in OPS if total[o]))
    cols = [o for o in OPS if total[o]]
    print("%-8s " % "instrs" + " ".join("%8s" % c[:8] for c in cols) + "  kernel")
    for (name, c), dm in zip(counts.items(), demangle):
        short = re.sub(r"\(anonymous namespace\)::|vcm::", "", dm).replace("void ", "")
        short = re.sub(r"\(.*$", "", short) if not short.startswith("(") else short
        short = short or name
        print("%-8d " % c["_all"] + " ".join("%8d" % c[o] for o in cols) + "  " + short[:110])


if __name__ == "__main__":
    main()
