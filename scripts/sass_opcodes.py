#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmscl_b200.so (cuobjdump -sass), the Blackwell evidence the profiling recipe asks
for: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UBLKCP = TMA, LDGSTS = cp.async,
HMMA would be the legacy mma.sync path (none expected).

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mscl_b200", "lib", "libmscl_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP",
       "UBLKRED", "LDGSTS", "SYNCS", "ELECT", "HMMA", "IMMA", "HGMMA", "MUFU", "ATOMG", "REDG", "RED", "ATOM", "ATOMS", "MEMBAR", "BAR",
       "SHFL", "LDG", "STG", "LDS", "STS", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instructions per kernel, selected opcodes (tensor / TMA / async first)")
    print(f"# {len(kernels)} kernels; arch sm_100a")
    for (mangled, cnt), name in zip(kernels.items(), names):
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        total = sum(cnt.values())
        sel = "  ".join(f"{k}={cnt[k]}" for k in KEY if cnt.get(k))
        print(f"{name:<64} total={total:<6} {sel}")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("# library totals: " + "  ".join(f"{k}={tot[k]}" for k in KEY if tot.get(k)))
    assert tot.get("HMMA", 0) == 0 and tot.get("HGMMA", 0) == 0, "legacy tensor-core path found"


if __name__ == "__main__":
    sys.exit(main())
