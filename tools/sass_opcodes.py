#!/usr/bin/env python
"""SASS opcode counts per kernel of the built library (tensor-core / TMA / TMEM / mbarrier families), from
`cuobjdump -sass`: the evidence that the contraction kernels are tcgen05 / TMA code.  No GPU needed.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "3m-asr-inference_b200/libb200moe.so"
FAMILIES = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMACCTL", "UBLKPF", "UBLKCP", "LDTM", "STTM", "UTCBAR",
            "UTCATOMSWS", "SYNCS", "ACQBULK", "HMMA", "UTMAPF")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1).rstrip(".")
        if op.startswith(FAMILIES):
            cur[op] += 1

print(f"SASS opcode counts per kernel (tensor-core / TMA / TMEM / mbarrier families) of {LIB},")
print("from `cuobjdump -sass` (tools/sass_opcodes.py); kernels without any of them (elementwise / SIMT kernels) are left out.")
print("ffn_kernel<OutT, mode, ctas, tf32>: mode 0 product / 1 per-event trace / 2 timeline marks; ctas 2 = cta_group::2 pairs.")
tot = collections.Counter()
for name, c in per.items():
    if not c or set(c) == {"ACQBULK"}:
        continue
    tot.update(c)
    short = re.sub(r"\(.*", "", demangle(name).replace("(anonymous namespace)::", "")).replace("void ", "")
    print(f"\n{short}\n    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
print("\nwhole library\n    " + ", ".join(f"{k} x{v}" for k, v in sorted(tot.items())))
