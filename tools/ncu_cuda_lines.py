"""Per-CUDA-source-line summary of an `ncu --set full --import-source on` report (kernels built with -lineinfo):
    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > dump.csv
    python tools/ncu_cuda_lines.py dump.csv [topN]
Prints, per source line, the warp-instructions executed (share of the kernel) and the stall samples."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, errors="replace")))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            ss = hdr.index("Warp Stall Sampling (All Samples)")
            continue
        if hdr and len(r) >= len(hdr) and r[0].strip().isdigit():      # a CUDA-C line (SASS rows have an empty Line No)
            try:
                lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ie] or 0), int(r[ss] or 0)))
            except ValueError:
                pass
    tot_i = sum(x[3] for x in lines) or 1
    tot_s = sum(x[4] for x in lines) or 1
    print(f"{len(lines)} source lines, {tot_i} warp-instructions, {tot_s} stall samples")
    print("inst%  stall%  file:line  source")
    for f, ln, src, ni, ns in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"{100 * ni / tot_i:5.2f}  {100 * ns / tot_s:5.2f}  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
