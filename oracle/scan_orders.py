"""CPU ORACLE (test infrastructure, NOT product code) -- scan-order tables.

numpy restatement of the reference's scan-order generators for the 2D selective-scan
hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.

Every function returns the *flat* pixel index `idx[k, t]` such that
    xs[b, k, c, t] = x[b, c].reshape(-1)[idx[k, t]]
for a square H x W map (the reference only supports H == W).

Pinned: bit-for-bit against the reference generators themselves, run in the build
container by tests/golden/make_golden.py (tables committed in tests/golden/orders.npz).

Reference lines followed:
  std       Models/SS2D/csms6s.py:13-31        (CrossScan.forward)
  dilation  Models/SS2D/Dilation.py:3-45,66    (generate_dilation_indices, flat = c0*H+c1)
  window    Models/SS2D/Window.py:3-35,56      (generate_window_indices,   flat = c0*H+c1)
  line      Models/SS2D/SpiralLine.py:3-82,103 (bresenham, generate_indices, flat = x+y*H)
Ablation families (pinned by tests/golden/ablation_orders.npz = the reference
classes' own outputs, made by tests/golden/make_golden_ablation.py):
  line4     Models/SS2D/csms6s.py:219-265      (the four line lists alone)
  spiral    Models/SS2D/csms6s.py:264-339 + Spiral.py:3-82
  hilbert   Models/SS2D/csms6s.py:374-405 + Hilbert.py:9-331 (sides 2^m and 3*2^m only)
  diagonal  Models/SS2D/csms6s.py:478-549
  ds        Models/SS2D/csms6s.py:619-631
  ab2 / ab1 Models/SS2D/csms6s.py:678-685, 713-718
"""
import numpy as np

# reference constants: Models/SS2D/csms6s.py:58-61 (dilation), :107-110 (window)
REF_SIZES = (12, 24, 48, 96)
REF_DILATION_RATE = {12: 4, 24: 4, 48: 4, 96: 4}
REF_WINDOW_SIZE = {12: 4, 24: 8, 48: 12, 96: 16}
# build-side extension for the 768^2 stress config: dilation 4, window 32
EXT_DILATION_RATE = {192: 4}
EXT_WINDOW_SIZE = {192: 32}
ABLATION = ("line4", "spiral", "hilbert", "diagonal", "ds", "ab2", "ab1")


def std_indices(H, W):
    """csms6s.py:19-21: k0 row-major, k1 col-major (x.transpose(2,3).flatten), k2/k3 = flips."""
    L = H * W
    t = np.arange(L, dtype=np.int64)
    k0 = t
    # xs[:,1][t] = x^T.flatten()[t] = x[t % H, t // H]  (x^T has shape (W, H))
    k1 = (t % H) * W + (t // H)
    return np.stack([k0, k1, k0[::-1], k1[::-1]])


def dilation_indices(H, W, rate=4):
    """Dilation.py:16-43.  coords[0]=(i,j) raster, coords[1]=(j,i), coords[2/3]=reversed
    (lines 17-24); each sequence is regrouped by phase p % rate (26-36); the returned list
    is [from c0, from c1, from c2, from c3] (43).  flat = c[0]*H + c[1] (66)."""
    L = H * W
    ii, jj = np.divmod(np.arange(L, dtype=np.int64), W)
    c0 = np.stack([ii, jj], 1)
    c1 = np.stack([jj, ii], 1)
    seqs = [c0, c1, c0[::-1], c1[::-1]]
    pos = np.arange(L)
    order = np.concatenate([pos[pos % rate == r] for r in range(rate)])
    out = []
    for c in seqs:
        cc = c[order]
        out.append(cc[:, 0] * H + cc[:, 1])
    return np.stack(out)


def window_indices(H, W, ws=4):
    """Window.py:15-34.  horizontal: (i+x, j+y) x outer / y inner; vertical: (j+x, i+y)
    with y outer / x inner (26); list = [h, h reversed, v, v reversed].  flat = c0*H+c1."""
    assert 0 < ws < H
    hor, ver = [], []
    for i in range(0, H, ws):
        for j in range(0, W, ws):
            for x in range(ws):
                for y in range(ws):
                    hor.append((i + x, j + y))
            for y in range(ws):
                for x in range(ws):
                    ver.append((j + x, i + y))
    hor = np.asarray(hor, dtype=np.int64)
    ver = np.asarray(ver, dtype=np.int64)
    seqs = [hor, hor[::-1], ver, ver[::-1]]
    return np.stack([c[:, 0] * H + c[:, 1] for c in seqs])


def _bresenham(x0, y0, x1, y1):
    """SpiralLine.py:3-24 verbatim semantics (both ifs evaluated each step)."""
    pts = []
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx = 1 if x0 < x1 else -1
    sy = 1 if y0 < y1 else -1
    err = dx - dy
    while True:
        pts.append((x0, y0))
        if x0 == x1 and y0 == y1:
            break
        e2 = 2 * err
        if e2 > -dy:
            err -= dy
            x0 += sx
        if e2 < dx:
            err += dx
            y0 += sy
    return pts


def line_indices(H, W):
    """SpiralLine.py:27-82: families E / E-flip / O / O-flip of Bresenham lines through the
    centre; per-line reversal for the flip lists.  flat = x + y*H (103)."""
    E, Ef, O, Of = [], [], [], []
    for r in range(0, H, 2):
        p = _bresenham(0, r, H - 1, W - 1 - r)
        E += p
        Ef += p[::-1]
    start = 0 if H % 2 == 0 else 2
    for c in range(start, W, 2):
        p = _bresenham(c, W - 1, H - 1 - c, 0)
        E += p
        Ef += p[::-1]
    for r in range(1, H, 2):
        p = _bresenham(0, r, H - 1, W - 1 - r)
        O += p
        Of += p[::-1]
    if H % 2 != 0:
        p = _bresenham(0, W - 1, H - 1, 0)
        O += p
        Of += p[::-1]
    for c in range(1, W, 2):
        p = _bresenham(c, W - 1, H - 1 - c, 0)
        O += p
        Of += p[::-1]
    out = []
    for pts in (E, Ef, O, Of):
        a = np.asarray(pts, dtype=np.int64)
        out.append(a[:, 0] + a[:, 1] * H)
    return out  # list: lengths equal H*W only for even H


def spiral_indices(H, W):
    """Spiral.py:3-60 + 62-71: clockwise from the top-left, counter-clockwise from the bottom-right; the
    walks emit (row, col) and the gather index is `y*H + x` with (x, y) = (row, col): the TRANSPOSED pixel."""
    def walk(ccw):
        top, bottom, left, right = 0, H - 1, 0, W - 1
        pts = []
        while top <= bottom and left <= right:
            if not ccw:
                pts += [(top, i) for i in range(left, right + 1)]
                top += 1
                pts += [(i, right) for i in range(top, bottom + 1)]
                right -= 1
                if top <= bottom:
                    pts += [(bottom, i) for i in range(right, left - 1, -1)]
                    bottom -= 1
                if left <= right:
                    pts += [(i, left) for i in range(bottom, top - 1, -1)]
                    left += 1
            else:
                pts += [(bottom, i) for i in range(right, left - 1, -1)]
                bottom -= 1
                pts += [(i, left) for i in range(bottom, top - 1, -1)]
                left += 1
                if top <= bottom:
                    pts += [(top, i) for i in range(left, right + 1)]
                    top += 1
                if left <= right:
                    pts += [(i, right) for i in range(top, bottom + 1)]
                    right -= 1
        a = np.asarray(pts, dtype=np.int64)
        return a[:, 1] * H + a[:, 0]
    return np.stack([walk(False), walk(True)])


def diagonal_indices(H, W):
    """csms6s.py:478-497: gather(3, (h + w) % W) / gather(3, (w - h) % W), transpose(-1, -2), flatten:
    token t = w*H + h reads x[h, (h + w) % W] (diagonal) and x[h, (w - h) % W] (anti-diagonal)."""
    t = np.arange(H * W, dtype=np.int64)
    w, h = t // H, t % H
    return np.stack([h * W + (h + w) % W, h * W + (w - h) % W])


def hilbert_indices(H):
    """Hilbert.py:9-331 restricted to what a square of side 2^m or 3*2^m reaches: even-even blocks are
    quartered (:187-200), odd-odd 3x3 blocks split 2+1 (:242-257), blocks of <= 6 cells are rendered as
    strips (:33-64).  Returns the (L,) walk `int(y-1)*H + int(x-1)` (:8); raises ValueError elsewhere."""
    out = []

    def strip(x0, y0, ax, ay, ox, oy, back):
        n = abs(ax + ay)
        ddx, ddy = ax / n, ay / n
        for ii in (range(n - 1, -1, -1) if back else range(n)):
            out.append(int(y0 + ii * ddy + (oy + ddy - 1) / 2 - 1) * H + int(x0 + ii * ddx + (ox + ddx - 1) / 2 - 1))

    def go(x0, y0, xl, yl, xr, yr, d):
        nl, nr = abs(xl + yl), abs(xr + yr)
        if nl * nr <= 6:
            if nl == 1:
                return strip(x0, y0, xr, yr, xl, yl, False)
            if nr == 1:
                return strip(x0, y0, xl, yl, xr, yr, False)
            if d == "l":
                strip(x0, y0, xr, yr, xl / 2, yl / 2, False)
                return strip(x0, y0, xr, yr, xl + xl / 2, yl + yl / 2, True)
            if d == "r":
                strip(x0, y0, xl, yl, xr / 2, yr / 2, False)
                return strip(x0, y0, xl, yl, xr + xr / 2, yr + yr / 2, True)
            raise ValueError("hilbert: leaf outside the restricted set")
        if 2 * nl > 3 * nr or 2 * nr > 3 * nl or d == "m":
            raise ValueError("hilbert: side not 2^m or 3*2^m")
        xl2, yl2, xr2, yr2 = round(xl / 2), round(yl / 2), round(xr / 2), round(yr / 2)
        if nl % 2 == 0 and nr % 2 == 0:
            if abs(xl2 + yl2 + xr2 + yr2) % 2:
                raise ValueError("hilbert: side not 2^m or 3*2^m")
            mid = None
        elif nl % 2 and nr % 2:
            xl2 = xl - xl2 if xl2 % 2 else xl2
            yl2 = yl - yl2 if yl2 % 2 else yl2
            xr2 = xr - xr2 if xr2 % 2 else xr2
            yr2 = yr - yr2 if yr2 % 2 else yr2
            mid = "m"
        else:
            raise ValueError("hilbert: side not 2^m or 3*2^m")
        if d == "l":
            go(x0, y0, xl2, yl2, xr2, yr2, "r")
            if mid:
                go(x0 + xr2, y0 + yr2, xl2, yl2, xr - xr2, yr - yr2, "m")
                go(x0 + xr + xl2, y0 + yr + yl2, xr2 - xr, yr2 - yr, xl - xl2, yl - yl2, "m")
            else:
                go(x0 + xr2, y0 + yr2, xl2, yl2, xr - xr2, yr - yr2, "l")
                go(x0 + xr2 + xl2, y0 + yr2 + yl2, xl - xl2, yl - yl2, xr - xr2, yr - yr2, "l")
            go(x0 + xr2 + xl, y0 + yr2 + yl, xl2 - xl, yl2 - yl, -xr2, -yr2, "r")
        else:
            go(x0, y0, xl2, yl2, xr2, yr2, "l")
            if mid:
                go(x0 + xl2, y0 + yl2, xl - xl2, yl - yl2, xr2, yr2, "m")
                go(x0 + xl + xr2, y0 + yl + yr2, xr - xr2, yr - yr2, xl2 - xl, yl2 - yl, "m")
            else:
                go(x0 + xl2, y0 + yl2, xl - xl2, yl - yl2, xr2, yr2, "r")
                go(x0 + xr2 + xl2, y0 + yr2 + yl2, xl - xl2, yl - yl2, xr - xr2, yr - yr2, "r")
            go(x0 + xl2 + xr, y0 + yl2 + yr, -xl2, -yl2, xr2 - xr, yr2 - yr, "l")

    go(1, 1, H, 0, 0, H, "l")
    a = np.asarray(out, dtype=np.int64)
    if len(a) != H * H or len(np.unique(a)) != H * H:
        raise ValueError("hilbert: walk is not a permutation")
    return a


def order_table(family, H, param=0):
    """(K, L) int64 table for a named family as the reference's CrossScan* classes use it.
    family: 'std' | 'dilation' | 'window' | 'line' (line = std k0-3 + 4 line lists,
    csms6s.py:167-171) | ablation: 'line4' | 'spiral' | 'hilbert' | 'diagonal' | 'ds' | 'ab2' | 'ab1'."""
    if family == "std":
        return std_indices(H, H)
    if family == "dilation":
        return dilation_indices(H, H, param or 4)
    if family == "window":
        return window_indices(H, H, param)
    if family == "line":
        ln = line_indices(H, H)
        assert all(len(a) == H * H for a in ln), "line family needs even H"
        return np.concatenate([std_indices(H, H), np.stack(ln)])
    if family in ("line4", "ds"):
        ln = line_indices(H, H)
        assert all(len(a) == H * H for a in ln), "line family needs even H"
        if family == "line4":
            return np.stack(ln)
        dg = diagonal_indices(H, H)
        return np.concatenate([np.stack(ln), dg, dg[:, ::-1]])          # csms6s.py:624-630
    if family == "spiral":
        sp = spiral_indices(H, H)
        return np.concatenate([std_indices(H, H), sp, sp[:, ::-1]])      # csms6s.py:321-326
    if family == "diagonal":
        dg = diagonal_indices(H, H)
        return np.concatenate([std_indices(H, H), dg, dg[:, ::-1]])      # csms6s.py:539-547
    if family == "hilbert":
        h0 = hilbert_indices(H)
        h1 = (H - 1 - h0 // H) * H + h0 % H                               # flip(input, [2]) (csms6s.py:401)
        return np.stack([h0, h1, h0[::-1], h1[::-1]])
    if family in ("ab2", "ab1"):
        t = np.arange(H * H, dtype=np.int64)
        return np.stack([t, t, t[::-1], t[::-1]] if family == "ab2" else [t] * 4)
    raise ValueError(family)


def cross_scan(x, table):
    """x: (B, C, H, W) -> (B, K, C, L); csms6s.py:18-22 / Dilation.py:60-69."""
    B, C, H, W = x.shape
    flat = x.reshape(B, C, H * W)
    return np.stack([flat[:, :, table[k]] for k in range(table.shape[0])], axis=1)


def cross_merge(ys, table, family):
    """ys: (B, K, D, L) -> (B, D, L) with the reference's association order.

    std      csms6s.py:39-42   (ys0 + flip(ys2)) + T(ys1 + flip(ys3))
    dil/win  Dilation.py:86-96 zeros, then scatter_add k = 0..3 in t order
    line     csms6s.py:196-201 y_h (scatter_add k=4..7 from zeros) + y_rc (std fold)
    """
    B, K, D, L = ys.shape
    dt = ys.dtype

    def scatter(ks):
        out = np.zeros((B, D, L), dtype=dt)
        for k in ks:
            # sequential in t: np.add.at keeps index order => same order as CPU scatter_add_
            np.add.at(out, (slice(None), slice(None), table[k]), ys[:, k])
        return out

    def std_fold():
        H = int(round(L ** 0.5))
        a = ys[:, 0] + ys[:, 2][..., ::-1]
        b = ys[:, 1] + ys[:, 3][..., ::-1]
        bt = b.reshape(B, D, H, H).transpose(0, 1, 3, 2).reshape(B, D, L)
        return a + bt

    if family == "std":
        return std_fold()
    if family in ("dilation", "window"):
        return scatter(range(4))
    if family == "line":
        return scatter(range(4, 8)) + std_fold()
    if family in ABLATION or family == "custom":
        # the exact transpose, added sequentially over k then t.  The reference groups these sums
        # differently per family (e.g. y_h + y_rc, csms6s.py:333-338): same value up to fp32 association
        return scatter(range(K))
    raise ValueError(family)
