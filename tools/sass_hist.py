"""Opcode histogram per kernel of the shipped library (cuobjdump -sass): python tools/sass_hist.py > profiles/r2_sass_opcodes.txt
Evidence that the binary is sm_100a code using the units the design names (UBLKCP / SYNCS = bulk-TMA + mbarrier, UCGABAR =
cluster barrier, ATOMG / REDG = global atomics, STG.E.ENL2.256 = 256-bit stores, REDUX / MATCH / SHFL = warp collectives)."""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "muvo_b200", "libmuvo_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, hist, arch = None, collections.OrderedDict(), set()
for ln in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", ln)
    if m and kern:
        op = m.group(1)
        base = op.split(".")[0]
        hist[kern][base] += 1
        if base in ("STG", "LDG", "ATOMG", "REDG", "UBLKCP", "SYNCS", "ATOMS", "UCGABAR_ARV", "UCGABAR_WAIT", "REDUX", "MATCH"):
            hist[kern]["  " + op] += 1
print("library:", os.path.basename(lib), "architectures:", ", ".join(sorted(arch)), "kernels:", len(hist))
interesting = ("UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "ATOMG", "REDG", "ATOMS", "REDUX", "MATCH", "SHFL", "DADD", "DMUL", "DFMA",
               "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "BAR")
for k, h in hist.items():
    name = demangle(k)
    name = re.sub(r"\(anonymous namespace\)::|muvo::|\(.*$", "", name)[:90]
    tot = sum(v for o, v in h.items() if not o.startswith("  "))
    picks = " ".join(f"{o}={h[o]}" for o in interesting if h.get(o))
    wide = " ".join(f"{o.strip()}={v}" for o, v in h.items() if o.startswith("  ") and (".256" in o or ".128" in o or "MAX" in o or "CAS" in o))
    print(f"{name}: {tot} instr | {picks} | {wide}")
