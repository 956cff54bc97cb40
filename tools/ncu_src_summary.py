"""Summarise an `ncu --page source --csv` dump: instruction mix by opcode and the hottest
SASS lines by stall samples.  Usage: python tools/ncu_src_summary.py file.csv [topN]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr):
        data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
ie, ss, sn = ci["Instructions Executed"], ci["Warp Stall Sampling (All Samples)"], ci["Warp Stall Sampling (Not-issued Samples)"]
tot = sum(int(r[ie] or 0) for r in data)
tots = sum(int(r[ss] or 0) for r in data)
mix = collections.Counter()
smix = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
    op = m.group(2).split(".")[0] if m else "?"
    mix[op] += int(r[ie] or 0)
    smix[op] += int(r[ss] or 0)
print(f"total warp-instructions {tot}, stall samples {tots}, SASS lines {len(data)}")
print("opcode       inst%   stall%")
for op, n in mix.most_common(28):
    print(f"{op:12s} {100*n/tot:6.2f}  {100*smix[op]/max(tots,1):6.2f}")
print("\nhottest lines by stall samples:")
for r in sorted(data, key=lambda r: -int(r[ss] or 0))[:top]:
    print(f"{int(r[ss] or 0):7d} {int(r[sn] or 0):7d} {int(r[ie] or 0):9d}  {r[ci['Source']][:110]}")
