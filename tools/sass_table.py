"""Per-kernel SASS opcode table of libb200fusion.so (evidence that the hot kernels use the Blackwell tensor / TMA paths):

    python tools/sass_table.py > profiles/r02_sass_opcodes.md

Counts of UTCHMMA[.2CTA] (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk
copy), HMMA (mma.sync), LDGSTS (cp.async), LDSM (ldmatrix), MUFU.EX2 per kernel, from `cuobjdump -sass`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "simple-multimodal_b200", "libb200fusion.so")
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "LDSM", "MOVM", "MUFU.EX2", "ATOMS", "RED"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts, order, cur, it = {}, [], None, iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", next(it))
        cur = re.sub(r"^void ", "", cur).replace("b200f::(anonymous namespace)::", "b200f::")
        if cur not in counts:
            counts[cur] = collections.Counter()
            order.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    if op.startswith("UTCHMMA"):
        counts[cur]["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
    elif op.startswith("MUFU.EX2"):
        counts[cur]["MUFU.EX2"] += 1
    else:
        for k in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "LDSM", "MOVM", "ATOMS", "RED"):
            if op.startswith(k):
                counts[cur][k] += 1
                break

print("# SASS opcode counts per kernel of libb200fusion.so (`cuobjdump -sass`, sm_100a)\n")
print("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,")
print("HMMA = mma.sync (the narrow-side attention kernels), LDGSTS = cp.async, LDSM = ldmatrix, MOVM = movmatrix.\n")
print("| kernel | SASS instr | " + " | ".join(OPS) + " |")
print("|---|---:|" + "---:|" * len(OPS))
tot = collections.Counter()
for k in sorted(order, key=lambda n: -sum(counts[n][o] for o in OPS[:7])):
    c = counts[k]
    if not any(c[o] for o in OPS[:11]):
        continue
    tot.update(c)
    print(f"| `{k[:110]}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
print(f"| **all kernels listed** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
