#!/usr/bin/env python
"""tools/sass_summary.py [LIB] > profiles/rN_sass_summary.txt
Per kernel of libgca.so: instruction count, the ten most frequent SASS mnemonics and the counts of the mnemonics that
prove how it moves data (UTMALDG = TMA tensor load, SYNCS = mbarrier, LDGSTS = cp.async, LDG.E.128 / STG.E.128 = 128-bit
global accesses, ATOMS / REDUX / SHFL / BAR / PRMT / POPC), from `cuobjdump -sass`; plus the lines around every TMA
instruction."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gym_cellular_automata_b200/libgca.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "ATOMS", "ATOMG",
        "REDG", "REDUX", "SHFL", "BAR", "PRMT", "POPC", "LOP3", "SHF", "IMAD", "HMMA", "UTC")
cur, kernels = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append(m.group(1).strip())
arch = re.search(r"arch = (sm_\w+)", txt)
print(f"{lib}: {len(kernels)} kernels, {arch.group(1) if arch else '?'}\n")
demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
for (name, ins), dn in zip(kernels.items(), demangle + [""] * len(kernels)):
    ops = collections.Counter()
    for i in ins:
        i = re.sub(r"^@!?U?P\d+\s+", "", i)
        ops[i.split()[0]] += 1
    short = dn.split("(")[0][-70:] if dn else name[-70:]
    print(f"== {short}: {len(ins)} instructions")
    print("   top: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(10)))
    hits = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in KEYS}
    print("   evidence: " + ", ".join(f"{k} {v}" for k, v in hits.items() if v))
    for idx, i in enumerate(ins):
        if "UTMALDG" in i or "UTMASTG" in i:
            print("   TMA: " + " | ".join(ins[max(0, idx - 2):idx + 3]))
