#!/usr/bin/env python
"""tools/sass_summary.py — per-kernel SASS evidence of libvdl2gpu.so for profiles/: architecture, registers / shared memory / spills
(cuobjdump --dump-resource-usage) and counts of the instructions that matter on this path:
  UBLKCP (TMA bulk copy), SYNCS (mbarrier), LDGSTS (cp.async), FFMA2/FMUL2/FADD2 (packed f32x2), FFMA/FMUL/FADD, DFMA/DMUL/DADD,
  MUFU, LDS/STS, LDG/STG, ATOM/RED, SHFL/VOTE, and - expected absent, there is no dense contraction - UTC*MMA / HMMA / LDTM.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dumpvdl2_b200", "libvdl2gpu.so")
GROUPS = OrderedDict([
    ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("LDGSTS", r"^LDGSTS"), ("FFMA2/FMUL2/FADD2", r"^(FFMA2|FMUL2|FADD2)"),
    ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD)(\.|$)"), ("DFMA/DMUL/DADD", r"^(DFMA|DMUL|DADD)"), ("MUFU", r"^MUFU"),
    ("F2F/I2F/F2I", r"^(F2F|I2F|F2I|I2FP|F2FP)"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG(\.|$)"), ("STG", r"^STG"),
    ("LDL/STL", r"^(LDL|STL)"), ("ATOM/RED", r"^(ATOM|ATOMG|RED)"), ("SHFL/VOTE", r"^(SHFL|VOTE)"), ("BAR", r"^BAR"),
    ("tensor (UTC*MMA/HMMA/LDTM)", r"^(UTC.*MMA|HMMA|IMMA|LDTM|STTM|UTCBAR)")])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    print(f"# libvdl2gpu.so  arch = {', '.join(arch)}   (cuobjdump -sass / --dump-resource-usage)")
    fn, counts, total = None, None, 0
    out = []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn:
                out.append((fn, counts, total))
            fn, counts, total = m.group(1), Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            total += 1
            op = m.group(1)
            for g, pat in GROUPS.items():
                if re.match(pat, op):
                    counts[g] += 1
    if fn:
        out.append((fn, counts, total))
    demangle = subprocess.run(["c++filt"] + [f for f, _, _ in out], capture_output=True, text=True).stdout.splitlines()
    for (f, c, t), name in sorted(zip(out, demangle), key=lambda x: x[1]):
        print(f"\n{name}")
        print(f"  {usage.get(f, '')}")
        print(f"  instructions {t}: " + ", ".join(f"{g} {c[g]}" for g in GROUPS if c[g]))
    tens = sum(c["tensor (UTC*MMA/HMMA/LDTM)"] for _, c, _ in out)
    print(f"\n# tensor-core instructions in the library: {tens} (none expected: the path has no dense contraction)")


if __name__ == "__main__":
    main()
