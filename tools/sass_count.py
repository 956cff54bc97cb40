#!/usr/bin/env python
"""Static SASS statistics per kernel of an object file / shared library (host only, no GPU):
instruction count, opcode histogram, local-memory (spill) accesses.  For the fully unrolled round loops of the
fused SS2D kernels the static count of the loop body is what every thread executes per round, so
(instructions / 18 tokens) can be compared between variants before spending GPU time.

    python tools/sass_count.py tramba_b200/_obj/ss2d_fused_bwd.o [name-filter]
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, cur = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, cur


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    path = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, ins in kernels(path):
        dn = demangle(name)
        if flt and flt not in dn:
            continue
        ops = collections.Counter()
        for _, text in ins:
            t = re.sub(r"^@!?U?P\w+\s+", "", text)
            ops[t.split()[0].split(".")[0]] += 1
        # backward branches = loops; the largest one is the round loop
        loops = []
        for addr, text in ins:
            m = re.search(r"\bBRA\S*\s+(?:!?U?P\w+,\s*)?(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", text)
            if m and m.group(1) and int(m.group(1), 16) < addr:
                tgt = int(m.group(1), 16)
                loops.append((sum(1 for a, _ in ins if tgt <= a <= addr), tgt, addr))
        loops.sort(reverse=True)
        short = re.sub(r"\(anonymous namespace\)::|tramba::", "", dn)[:100]
        print(f"{short}\n  {len(ins)} instructions; spills LDL {ops['LDL']} STL {ops['STL']}; "
              f"MUFU {ops['MUFU']} LDS {ops['LDS']} STS {ops['STS']} LDGSTS {ops['LDGSTS']} "
              f"IMAD {ops['IMAD']} IADD3 {ops['IADD3']} LEA {ops['LEA']} BRA {ops['BRA']}")
        print("  largest loops (instructions in body): " + ", ".join(str(l[0]) for l in loops[:5]))


if __name__ == "__main__":
    main()
