"""Summarise an `ncu --page source --csv` dump: opcode histogram + hottest SASS lines.
usage: python tools/ncu_src.py file.csv [n_top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) > ix["Instructions Executed"] and r[0].startswith("0x")]
gi = lambda r, k: int(float(r[ix[k]] or 0))
tot_inst = sum(gi(r, "Instructions Executed") for r in body)
tot_samp = sum(gi(r, "# Samples") for r in body) or 1
print("total warp-inst", tot_inst, "samples", tot_samp, "sass lines", len(body))
c, cs = Counter(), Counter()
for r in body:
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    c[op] += gi(r, "Instructions Executed")
    cs[op] += gi(r, "# Samples")
for op, v in c.most_common(16):
    print(f"{op:10s} inst {100 * v / tot_inst:5.1f}%  samples {100 * cs[op] / tot_samp:5.1f}%")
print("--- top sampled instructions")
for r in sorted(body, key=lambda r: -gi(r, "# Samples"))[:ntop]:
    print(str(gi(r, "# Samples")).rjust(6), str(gi(r, "Instructions Executed")).rjust(9), r[ix["Source"]][:100])
