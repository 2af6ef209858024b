#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library (no GPU needed): registers / spills / smem from ``cuobjdump -res-usage`` and
counts of the Blackwell-specific mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG,
tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, red.global -> REDG) from ``cuobjdump -sass``.

    python scripts/sass_summary.py [lib.so] > profiles/rNN_sass_summary.csv
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "thinkdiff_mlre_b200/libthinkdiff_b200.so"
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTMAREDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "REDG",
        "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::|td::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("(EpiKind)", "")


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+.*)", res):
        usage[m.group(1)] = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur, n_instr = {}, None, collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z_][A-Z0-9_.]*)", line)
        if cur and m:
            op = m.group(1)
            n_instr[cur] += 1
            head = op.split(".")[0]
            counts[cur][head] += 1
            if ".2CTA" in op:
                counts[cur][head + ".2CTA"] += 1
            if head.startswith("UCGABAR"):
                counts[cur]["UCGABAR"] += 1
    names = demangle(sorted(counts))
    print("kernel,regs,stack_bytes,static_smem,instructions," + ",".join(COLS))
    for k in sorted(counts, key=lambda k: short(names[k])):
        u = usage.get(k, {})
        row = [f'"{short(names[k])}"', u.get("REG", ""), u.get("STACK", ""), u.get("SHARED", ""), str(n_instr[k])]
        row += [str(counts[k].get(c, 0)) for c in COLS]
        print(",".join(row))


if __name__ == "__main__":
    main()
