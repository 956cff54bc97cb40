#!/usr/bin/env python
"""Per-instruction warp-stall REASONS of one kernel from an .ncu-rep (ncu --set full --import-source on), read on
the CPU box:  ncu -i <rep> --page source --csv | python tools/ncu_stall_reasons.py [top]
Prints the share of every stall reason and, for the biggest reasons, the SASS instructions that collect it
(the instruction a stall is attributed to is the one WAITING, i.e. the consumer)."""
import csv
import sys


def main():
    top = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rows = list(csv.reader(sys.stdin))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    tot = {n: sum(int(r[col[n]] or 0) for r in body) for n in reasons}
    allsum = sum(tot.values())
    inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
    print(f"{len(body)} SASS lines, {inst} warp-instructions, {allsum} stall samples")
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"  {100 * v / allsum:6.2f}%  {n}")
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:5]:
        if n in ("stall_selected", "stall_not_selected"):
            continue
        print(f"\n{n}: instructions that wait (samples, executed, SASS)")
        for r in sorted(body, key=lambda r: -int(r[col[n]] or 0))[:top]:
            print(f"  {int(r[col[n]]):6d} {int(r[col['Instructions Executed']] or 0):9d}  {r[col['Source']].strip()[:90]}")
    # shared-memory conflicts
    if "L1 Wavefronts Shared" in col:
        w = sum(int(r[col["L1 Wavefronts Shared"]] or 0) for r in body)
        wi = sum(int(r[col["L1 Wavefronts Shared Ideal"]] or 0) for r in body)
        print(f"\nshared-memory wavefronts {w}, ideal {wi} ({w / max(wi, 1):.2f}x)")
        print("instructions with the most excessive wavefronts (excess, wavefronts, ideal, SASS)")
        ex = col["L1 Wavefronts Shared Excessive"]
        for r in sorted(body, key=lambda r: -int(r[ex] or 0))[:top]:
            print(f"  {int(r[ex] or 0):8d} {int(r[col['L1 Wavefronts Shared']] or 0):8d} "
                  f"{int(r[col['L1 Wavefronts Shared Ideal']] or 0):8d}  {r[col['Source']].strip()[:80]}")


if __name__ == "__main__":
    main()
