#!/usr/bin/env python
"""SASS evidence of the Blackwell features per kernel:  python tools/sass_counts.py > profiles/r2_sass_counts.md

Runs `cuobjdump -sass` on the shipped library and counts, per kernel, the mnemonics that prove tcgen05 / tensor memory /
TMA use (B200_PROFILING.md): UTCHMMA (tcgen05.mma), STTM / LDTM (tcgen05.st / ld), UTMALDG (cp.async.bulk.tensor),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus the FP64 and packed-FP32 pipes the other kernels rely on."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "admm-quantization_b200", "lib", "libadmmq.so")
MNEMONICS = ["UTCHMMA", "STTM", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "DFMA", "FFMA2", "FFMA", "ATOMS", "RED", "BAR"]

out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), stdout=subprocess.PIPE, text=True).stdout.splitlines()
names = dict(zip(re.findall(r"Function : (\S+)", out), demangle))
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
        cut = cur.rfind(">(")
        cur = (cur[:cut + 1] if cut >= 0 else cur[:cur.find("(")] if "(" in cur else cur).replace("admmq::", "").replace("(int)", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        if op in MNEMONICS:
            counts[cur][op] += 1
sha = subprocess.run(["sha256sum", LIB], stdout=subprocess.PIPE, text=True).stdout.split()[0][:16]
print(f"# r2: SASS mnemonic counts per kernel of libadmmq.so (sha256 {sha}..., `cuobjdump -sass`, sm_100a)\n")
print("UTCHMMA = tcgen05.mma, STTM / LDTM = tcgen05.st / tcgen05.ld (tensor memory), UTMALDG = TMA tile load, UTCBAR = tcgen05.commit, "
      "SYNCS = mbarrier operations, DFMA = FP64 pipe, FFMA2 = packed FP32.\n")
print("| kernel | SASS instr | " + " | ".join(MNEMONICS) + " |")
print("|---|---:|" + "---:|" * len(MNEMONICS))
tot = collections.Counter()
for k, c in counts.items():
    print(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
    tot.update(c)
print(f"| **all kernels** | {tot['_total']} | " + " | ".join(str(tot[m]) for m in MNEMONICS) + " |")
