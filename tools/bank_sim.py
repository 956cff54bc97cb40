#!/usr/bin/env python
"""Shared-memory bank-conflict model of the fused SS2D kernels' table-driven gathers (host only, no GPU).

The fused kernels (tramba_b200/csrc/ss2d_fused_*.cu) read x / dy / the merged y through the scan-order table:
at unrolled step m, thread j of a plane touches element skew(idx[k][j*S + m]) of its plane (S = 18 tokens per
thread, planes laid out back to back with stride LSK).  A warp-level LDS costs max-over-banks(distinct 32-bit
words) wavefronts.  This script counts wavefronts per warp access for every (order, H, element size) of the
model under a candidate skew function, so layouts can be compared before spending GPU time.

    python tools/bank_sim.py                  # the current skew  p + 2*(p >> 6)
    python tools/bank_sim.py --pad 4 --seg 128   # p + 4*(p >> 7)  (16-byte aligned segments: bulk-copy friendly)
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
S, NT = 18, 512
WS = {12: 4, 24: 8, 48: 12, 96: 16}


def tables():
    from tramba_b200 import _lib
    from tramba_b200.ops import order_table_host
    out = {}
    for H in (12, 24, 48, 96):
        out[("std", H)] = order_table_host(_lib.ORDER_STD, H)
        out[("dilation", H)] = order_table_host(_lib.ORDER_DILATION, H, 4)
        out[("window", H)] = order_table_host(_lib.ORDER_WINDOW, H, WS[H])
        out[("line", H)] = order_table_host(_lib.ORDER_LINE, H)
    return out


def wavefronts(word, nlanes=32):
    """word: (..., 32) int array of 32-bit word addresses of one warp access -> wavefronts (..,)."""
    bank = word % 32
    w = np.zeros(word.shape[:-1], dtype=np.int64)
    for b in range(32):
        sel = np.where(bank == b, word, -1)
        sel = np.sort(sel, axis=-1)
        distinct = (np.diff(sel, axis=-1) != 0).sum(-1) + 1 - (sel[..., 0] == -1)   # drop the -1 group
        w = np.maximum(w, distinct)
    return w


def simulate(tab, H, es, pad, seg_shift, plane_extra=2):
    """Mean wavefronts per warp-level gather for one order table (K, L), element size es (2 or 4 bytes)."""
    K, L = tab.shape
    TP = L // S
    PP = NT // TP
    lsk = L + pad * ((L + (1 << seg_shift) - 1) >> seg_shift) + plane_extra
    tid = np.arange(NT)
    q, j = tid // TP, tid % TP
    res = []
    for k in range(K):
        tot = 0.0
        for m in range(S):
            p = tab[k][j * S + m].astype(np.int64)
            e = q * lsk + p + pad * (p >> seg_shift)           # element position in the staging area
            word = (e * es) // 4
            tot += wavefronts(word.reshape(NT // 32, 32)).mean()
        res.append(tot / S)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pad", type=int, default=2, help="pad elements inserted per segment")
    ap.add_argument("--seg", type=int, default=64, help="segment length in elements (power of two)")
    args = ap.parse_args()
    sh = args.seg.bit_length() - 1
    assert 1 << sh == args.seg
    tabs = tables()
    print(f"skew(p) = p + {args.pad}*(p >> {sh}); wavefronts per warp gather (1.0 = conflict-free)")
    print(f"{'order':9s} {'H':>3s}  " + "  ".join(f"{'es=%d k=%d' % (es, k):>9s}" for es in (2, 4) for k in range(4)) + "   mean2  mean4")
    grand = {2: [], 4: []}
    for (fam, H), tab in sorted(tabs.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        row = {es: simulate(tab, H, es, args.pad, sh) for es in (2, 4)}
        for es in (2, 4):
            grand[es] += row[es]
        kk = tab.shape[0]
        cells = "  ".join(f"{row[es][k]:9.2f}" for es in (2, 4) for k in (range(4) if kk == 4 else range(4, 8)))
        print(f"{fam:9s} {H:3d}  {cells}   {np.mean(row[2]):5.2f}  {np.mean(row[4]):5.2f}")
    print(f"all: es=2 {np.mean(grand[2]):.3f}   es=4 {np.mean(grand[4]):.3f}")


if __name__ == "__main__":
    main()
