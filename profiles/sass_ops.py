#!/usr/bin/env python
"""Per-kernel SASS evidence table for liblrp_b200.so (run in the build container, no GPU needed):
   counts of the Blackwell-native instructions (tcgen05 -> UTC*MMA, TMA -> UTMALDG/UTMASTG/UTMAREDG, TMEM ld/st -> LDTM/STTM),
   legacy tensor instructions (HMMA: none expected), MUFU.EX2, plus registers / spills / smem from the ptxas logs.
   usage: python profiles/sass_ops.py > profiles/r02_sass_ops.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lrp-explains-transformers_b200", "csrc")
SO = os.path.join(ROOT, "lrp-explains-transformers_b200", "lxt_b200", "lib", "liblrp_b200.so")
PAT = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU.EX2", "FMNMX3", "RED.E", "ATOM"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for p in PAT:
            if op.startswith(p):
                counts[cur][p] += 1
                break
    regs = {}
    for log in glob.glob(os.path.join(CSRC, "build", "*.ptxas.log")):
        fn = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                fn = m.group(1)
                regs[fn] = {}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and fn:
                regs[fn].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and fn:
                regs[fn]["regs"] = int(m.group(1))
    names = demangle(list(counts))
    print("# r02 — SASS evidence per kernel of liblrp_b200.so (`python profiles/sass_ops.py`, cuobjdump -sass + ptxas -v logs)\n")
    print("tcgen05.mma -> `UTCHMMA` (`.2CTA` = cta_group::2), TMA load/store/reduce -> `UTMALDG` / `UTMASTG` / `UTMAREDG`, "
          "tcgen05.ld/st -> `LDTM` / `STTM`, tcgen05.commit -> `UTCBAR`; `HMMA` (legacy mma.sync) must be 0.\n")
    cols = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU.EX2"]
    print("| kernel | " + " | ".join(cols) + " | regs | spill st/ld (B) |")
    print("|---|" + "---|" * (len(cols) + 2))
    tot = collections.Counter()
    for fn, c in counts.items():
        short = re.sub(r"\(.*", "", names.get(fn, fn).replace("(anonymous namespace)::", "")).replace("lrp::", "")
        short = re.sub(r"^void ", "", short)
        r = regs.get(fn, {})
        tot.update(c)
        if not any(c[k] for k in cols) and "kernel" not in short:
            continue
        print(f"| `{short}` | " + " | ".join(str(c[k]) for k in cols) + f" | {r.get('regs', '?')} | {r.get('spill_st', 0)}/{r.get('spill_ld', 0)} |")
    print("\nTotals: " + ", ".join(f"{k} {tot[k]}" for k in cols))


if __name__ == "__main__":
    main()
