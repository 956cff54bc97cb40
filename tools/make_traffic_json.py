"""profiles/roofline_traffic.json from an ncu per-launch DRAM-byte list of one bench step:
    ncu --profile-from-start off -k regex:'ss2d_|ln2d_|dwconv_|dct2d_|batch_sum|proj_' \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv python bench.py ...
Usage: python tools/make_traffic_json.py X.csv profiles/roofline_traffic.json
Writes {bench kind: mean (dram read + write) bytes per launch} -- bench.py's roofline.traffic."""
import collections
import csv
import json
import sys

KIND = {"ss2d_bwd_kernel": "ss2d_fused_bwd", "ss2d_fwd_kernel": "ss2d_fused_fwd", "batch_sum": "batch_sum",
        "ln2d_fwd": "layernorm2d_fwd", "ln2d_bwd": "layernorm2d_bwd", "dct2d_fwd": "dct2d_fwd", "dct2d_bwd": "dct2d_bwd",
        "proj_gather": "proj_gather", "proj_scatter": "proj_scatter", "dwconv_fwd": "dwconv_fwd", "dwconv_bwd": "dwconv_bwd"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(src, dst):
    lines = [ln for ln in open(src, errors="replace") if not ln.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ii, ki, mi, vi, ui = (hdr.index(h) for h in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    per = collections.defaultdict(float)      # (launch id, kind) -> bytes
    for row in r:
        if len(row) <= vi or not row[mi].startswith("dram__bytes"):
            continue
        kind = next((v for k, v in KIND.items() if k in row[ki]), None)
        if kind is None:
            continue
        per[(row[ii], kind)] += float(row[vi].replace(",", "")) * UNIT.get(row[ui], 1.0)
    agg = collections.defaultdict(list)
    for (_, kind), b in per.items():
        agg[kind].append(b)
    out = {k: sum(v) / len(v) for k, v in agg.items()}
    out["_launches"] = {k: len(v) for k, v in agg.items()}
    out["_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch over one eager bench step"
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
