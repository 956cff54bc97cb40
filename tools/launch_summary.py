"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/launch_summary.py launches.csv [top_n]  -> prints a table (share of device time, ms, launches)."""
import collections
import csv
import re
import sys


def main(path, top=60):
    lines = [ln for ln in open(path, errors="replace") if not ln.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in r:
        if len(row) <= vi:
            continue
        try:
            v = float(row[vi].replace(",", ""))
        except ValueError:
            continue
        u = row[ui]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v   # -> us
        name = row[ki]
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
        if name.startswith("at::native::"):
            m = re.search(r"at::native::(\w+)<.*?at::native::(\w+)", name)
            name = "at::" + (m.group(1) + "<" + m.group(2) + ">" if m else name[12:72])
        name = name[:90]
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"{n} launches, {tot / 1e3:.2f} ms device time (serialised, cold-cache ncu replay: shares, not absolutes)")
    ours = sum(v[1] for k, v in agg.items() if "tramba" in k)
    print(f"tramba:: kernels: {100 * ours / tot:.2f}% of device time")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * v[1] / tot:6.2f}% {v[1] / 1e3:9.3f} ms {v[0]:6d} x {v[1] / v[0]:9.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
