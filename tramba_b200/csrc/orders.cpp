// Host-side scan-order table generators (no GPU).  B200-side replacement of the reference's
// Python generators: Models/SS2D/Dilation.py:3-45, Models/SS2D/Window.py:3-35,
// Models/SS2D/SpiralLine.py:3-82 and the standard unfold of Models/SS2D/csms6s.py:13-31.
// All tables are flat pixel indices row*W + col of a square map; closed forms where the order
// is a permutation, an integer Bresenham walk for the helix lines.
// The ablation families follow Models/SS2D/csms6s.py:219-737, Spiral.py:3-82
// and Hilbert.py:9-331; each is the table T with  CrossScan_X(x)[b,k,c,t] = x[b,c].flat[T[k][t]].
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/tramba_b200.h"

namespace {

void std_table(int H, int W, int32_t* out) {
    const int L = H * W;
    for (int t = 0; t < L; ++t) {
        out[t] = t;                            // k0 row-major
        out[L + t] = (t % H) * W + (t / H);    // k1 column-major (csms6s.py:20)
    }
    for (int t = 0; t < L; ++t) {
        out[2 * L + t] = out[L - 1 - t];       // k2 = flip(k0)  (csms6s.py:21)
        out[3 * L + t] = out[L + (L - 1 - t)]; // k3 = flip(k1)
    }
}

// Dilation.py:16-43: four raster sequences (row-major, col-major, and their reversals), each
// regrouped by phase (position mod rate).  flat = c0*H + c1 (Dilation.py:66).
void dilation_table(int H, int W, int rate, int32_t* out) {
    const int L = H * W;
    std::vector<int32_t> seq(L);
    for (int k = 0; k < 4; ++k) {
        for (int p = 0; p < L; ++p) {
            const int q = (k >= 2) ? (L - 1 - p) : p;  // coords[2/3] = flip(coords[0/1])
            const int i = q / W, j = q % W;
            seq[p] = (k & 1) ? (j * H + i) : (i * H + j);
        }
        int w = 0;
        for (int r = 0; r < rate; ++r)
            for (int p = r; p < L; p += rate) out[k * L + w++] = seq[p];
    }
}

// Window.py:19-34: windows visited row-major; horizontal list walks a window row-major,
// vertical list uses (j + x, i + y) with y outer / x inner; then [h, rev h, v, rev v].
void window_table(int H, int W, int ws, int32_t* out) {
    const int L = H * W;
    int w = 0;
    for (int i = 0; i < H; i += ws)
        for (int j = 0; j < W; j += ws) {
            int base = w;
            for (int x = 0; x < ws; ++x)
                for (int y = 0; y < ws; ++y) out[base++] = (i + x) * H + (j + y);
            base = w;
            for (int y = 0; y < ws; ++y)
                for (int x = 0; x < ws; ++x) out[2 * L + base++] = (j + x) * H + (i + y);
            w += ws * ws;
        }
    for (int t = 0; t < L; ++t) {
        out[L + t] = out[L - 1 - t];
        out[3 * L + t] = out[2 * L + (L - 1 - t)];
    }
}

// SpiralLine.py:3-24 (both tests evaluated on every step).
void bresenham(int x0, int y0, int x1, int y1, std::vector<int>& xs, std::vector<int>& ys) {
    xs.clear();
    ys.clear();
    const int dx = std::abs(x1 - x0), dy = std::abs(y1 - y0);
    const int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1;
    int err = dx - dy;
    for (;;) {
        xs.push_back(x0);
        ys.push_back(y0);
        if (x0 == x1 && y0 == y1) break;
        const int e2 = 2 * err;
        if (e2 > -dy) {
            err -= dy;
            x0 += sx;
        }
        if (e2 < dx) {
            err += dx;
            y0 += sy;
        }
    }
}

// SpiralLine.py:27-82: families E, E-flip (each line reversed), O, O-flip.  Even H: E = even
// left-edge rows then even bottom-edge columns, O = the odd ones.  Odd H: E's columns start at
// 2 (SpiralLine.py:47-50) and O gains the anti-diagonal (SpiralLine.py:65-68).
// flat = x + y*H with x the column (SpiralLine.py:103).  Slots 4..7 of the K = 8 table.
int line_table(int H, int W, int32_t* out) {
    const int L = H * W;
    std::vector<int> xs, ys;
    int wE = 0, wO = 0;
    auto emit = [&](int fam, int& w) {
        const int n = static_cast<int>(xs.size());
        if (w + n > L) return false;
        for (int i = 0; i < n; ++i) {
            out[(4 + fam) * L + w + i] = xs[i] + ys[i] * H;
            out[(5 + fam) * L + w + i] = xs[n - 1 - i] + ys[n - 1 - i] * H;
        }
        w += n;
        return true;
    };
    for (int r = 0; r < H; r += 2) {
        bresenham(0, r, H - 1, W - 1 - r, xs, ys);
        if (!emit(0, wE)) return TRAMBA_E_ORDER;
    }
    for (int c = (H % 2 == 0) ? 0 : 2; c < W; c += 2) {
        bresenham(c, W - 1, H - 1 - c, 0, xs, ys);
        if (!emit(0, wE)) return TRAMBA_E_ORDER;
    }
    for (int r = 1; r < H; r += 2) {
        bresenham(0, r, H - 1, W - 1 - r, xs, ys);
        if (!emit(2, wO)) return TRAMBA_E_ORDER;
    }
    if (H % 2 != 0) {
        bresenham(0, W - 1, H - 1, 0, xs, ys);
        if (!emit(2, wO)) return TRAMBA_E_ORDER;
    }
    for (int c = 1; c < W; c += 2) {
        bresenham(c, W - 1, H - 1 - c, 0, xs, ys);
        if (!emit(2, wO)) return TRAMBA_E_ORDER;
    }
    return (wE == L && wO == L) ? TRAMBA_OK : TRAMBA_E_ORDER;
}

// out[dst] = reverse of out[src] along t (torch.flip(..., dims=[-1])).
void flip_row(int L, int32_t* out, int dst, int src) {
    for (int t = 0; t < L; ++t) out[dst * L + t] = out[src * L + (L - 1 - t)];
}

// Spiral.py:3-60.  The walks emit (row, col); the forward gather uses `y*H + x` with (x, y) = (row, col)
// (Spiral.py:68-71), i.e. the TRANSPOSED pixel col*H + row.  Slot `kc` clockwise from the top-left corner,
// slot `kc + 1` counter-clockwise from the bottom-right corner.
void spiral_tables(int H, int W, int32_t* out, int kc) {
    const int L = H * W;
    int32_t* cw = out + kc * L;
    int32_t* ccw = out + (kc + 1) * L;
    int w = 0;
    auto put = [&](int32_t* dst, int r, int c) { dst[w++] = c * H + r; };
    int top = 0, bottom = H - 1, left = 0, right = W - 1;
    while (top <= bottom && left <= right) {
        for (int i = left; i <= right; ++i) put(cw, top, i);
        ++top;
        for (int i = top; i <= bottom; ++i) put(cw, i, right);
        --right;
        if (top <= bottom) {
            for (int i = right; i >= left; --i) put(cw, bottom, i);
            --bottom;
        }
        if (left <= right) {
            for (int i = bottom; i >= top; --i) put(cw, i, left);
            ++left;
        }
    }
    w = 0, top = 0, bottom = H - 1, left = 0, right = W - 1;
    while (top <= bottom && left <= right) {
        for (int i = right; i >= left; --i) put(ccw, bottom, i);
        --bottom;
        for (int i = bottom; i >= top; --i) put(ccw, i, left);
        ++left;
        if (top <= bottom) {
            for (int i = left; i <= right; ++i) put(ccw, top, i);
            ++top;
        }
        if (left <= right) {
            for (int i = top; i <= bottom; ++i) put(ccw, i, right);
            --right;
        }
    }
}

// csms6s.py:478-497: gather(3, (h +/- w) % W) then transpose + flatten, i.e. token t = w*H + h reads
// x[h, (h + w) % W] (diagonal, slot kd) and x[h, (w - h) mod W] (anti-diagonal, slot kd + 1).
void diagonal_tables(int H, int W, int32_t* out, int kd) {
    const int L = H * W;
    for (int t = 0; t < L; ++t) {
        const int w = t / H, h = t % H;
        out[kd * L + t] = h * W + (h + w) % W;
        out[(kd + 1) * L + t] = h * W + ((w - h) % W + W) % W;
    }
}

// Hilbert.py:9-331 ("spacefill" of an arbitrary rectangle), restricted to the cases a square of side
// 2^m or 3*2^m reaches: halving even-even blocks (Hilbert.py:187-200), splitting odd-odd 3x3 blocks into
// 2+1 strips (:242-257) and the <= 6-cell leaves (:33-64).  Coordinates are the reference's 1-based
// doubles (block corners sit on half-integers after the /2 terms); anything outside the restricted set
// marks the walk as failed and the caller reports TRAMBA_E_ORDER.
struct HilbertWalk {
    int H;
    std::vector<int32_t>* out;
    bool ok = true;
    static int iabs(int v) { return v < 0 ? -v : v; }
    static int half_even(int v) { return static_cast<int>(std::nearbyint(v / 2.0)); }  // Python round()
    void emit(double x, double y) {                                                    // Hilbert.py:7-9
        out->push_back(static_cast<int>(y - 1) * H + static_cast<int>(x - 1));
    }
    // leaf: walk n cells along (ax, ay) at the half-offset (ox, oy), forwards or backwards
    void strip(double x0, double y0, int ax, int ay, double ox, double oy, bool backwards) {
        const int n = iabs(ax + ay);
        const double ddx = static_cast<double>(ax) / n, ddy = static_cast<double>(ay) / n;
        for (int s = 0; s < n; ++s) {
            const int ii = backwards ? n - 1 - s : s;
            emit(x0 + ii * ddx + (ox + ddx - 1) / 2, y0 + ii * ddy + (oy + ddy - 1) / 2);
        }
    }
    void go(int x0, int y0, int xl, int yl, int xr, int yr, char dir) {
        if (!ok) return;
        const int nl = iabs(xl + yl), nr = iabs(xr + yr);
        if (nl * nr <= 6) {
            if (nl == 1) return strip(x0, y0, xr, yr, xl, yl, false);                  // :34-39
            if (nr == 1) return strip(x0, y0, xl, yl, xr, yr, false);                  // :41-46
            if (dir == 'l') {                                                          // :48-55
                strip(x0, y0, xr, yr, xl / 2.0, yl / 2.0, false);
                return strip(x0, y0, xr, yr, xl + xl / 2.0, yl + yl / 2.0, true);
            }
            if (dir == 'r') {                                                          // :57-64
                strip(x0, y0, xl, yl, xr / 2.0, yr / 2.0, false);
                return strip(x0, y0, xl, yl, xr + xr / 2.0, yr + yr / 2.0, true);
            }
            ok = false;
            return;
        }
        // squares only: neither side is "much longer" than the other (:102,142), and dir is l / r
        if (2 * nl > 3 * nr || 2 * nr > 3 * nl || dir == 'm') {
            ok = false;
            return;
        }
        int xl2 = half_even(xl), yl2 = half_even(yl), xr2 = half_even(xr), yr2 = half_even(yr);
        const bool le = nl % 2 == 0, re = nr % 2 == 0;
        char mid;
        if (le && re) {
            if (iabs(xl2 + yl2 + xr2 + yr2) % 2 != 0) {                                // ee-oo / oo-ee: not reached
                ok = false;
                return;
            }
            mid = 0;
        } else if (!le && !re) {                                                       // odd-odd -> eo-eo (:244-247)
            if (xl2 % 2 != 0) xl2 = xl - xl2;
            if (yl2 % 2 != 0) yl2 = yl - yl2;
            if (xr2 % 2 != 0) xr2 = xr - xr2;
            if (yr2 % 2 != 0) yr2 = yr - yr2;
            mid = 'm';
        } else {
            ok = false;
            return;
        }
        if (dir == 'l') {
            go(x0, y0, xl2, yl2, xr2, yr2, 'r');
            if (mid) {                                                                 // :250-251
                go(x0 + xr2, y0 + yr2, xl2, yl2, xr - xr2, yr - yr2, 'm');
                go(x0 + xr + xl2, y0 + yr + yl2, xr2 - xr, yr2 - yr, xl - xl2, yl - yl2, 'm');
            } else {                                                                   // :193-194
                go(x0 + xr2, y0 + yr2, xl2, yl2, xr - xr2, yr - yr2, 'l');
                go(x0 + xr2 + xl2, y0 + yr2 + yl2, xl - xl2, yl - yl2, xr - xr2, yr - yr2, 'l');
            }
            go(x0 + xr2 + xl, y0 + yr2 + yl, xl2 - xl, yl2 - yl, -xr2, -yr2, 'r');
        } else {
            go(x0, y0, xl2, yl2, xr2, yr2, 'l');
            if (mid) {                                                                 // :255-256
                go(x0 + xl2, y0 + yl2, xl - xl2, yl - yl2, xr2, yr2, 'm');
                go(x0 + xl + xr2, y0 + yl + yr2, xr - xr2, yr - yr2, xl2 - xl, yl2 - yl, 'm');
            } else {                                                                   // :198-199
                go(x0 + xl2, y0 + yl2, xl - xl2, yl - yl2, xr2, yr2, 'r');
                go(x0 + xr2 + xl2, y0 + yr2 + yl2, xl - xl2, yl - yl2, xr - xr2, yr - yr2, 'r');
            }
            go(x0 + xl2 + xr, y0 + yl2 + yr, -xl2, -yl2, xr2 - xr, yr2 - yr, 'l');
        }
    }
};

// csms6s.py:374-405: k0 = x.flat[hilbert], k1 = flip(x, rows).flat[hilbert], k2/k3 = their reversals.
// get_hilber_spacefill(1, 1, H, H) -> spacefill(ll=1, tt=1, ww=H, hh=H) -> go(1, 1, H, 0, 0, H, "l").
int hilbert_table(int H, int W, int32_t* out) {
    const int L = H * W;
    std::vector<int32_t> walk;
    walk.reserve(L);
    HilbertWalk hw{H, &walk};
    hw.go(1, 1, W, 0, 0, H, 'l');
    if (!hw.ok || static_cast<int>(walk.size()) != L) return TRAMBA_E_ORDER;
    std::vector<char> seen(L, 0);
    for (int t = 0; t < L; ++t) {
        const int p = walk[t];
        if (p < 0 || p >= L || seen[p]) return TRAMBA_E_ORDER;
        seen[p] = 1;
        out[t] = p;
        out[L + t] = (H - 1 - p / W) * W + p % W;   // torch.flip(input, [2]) (csms6s.py:401)
    }
    flip_row(L, out, 2, 0);
    flip_row(L, out, 3, 1);
    return TRAMBA_OK;
}

}  // namespace

extern "C" int tramba_order_num_directions(int order) {
    switch (order) {
        case TRAMBA_ORDER_STD:
        case TRAMBA_ORDER_DILATION:
        case TRAMBA_ORDER_WINDOW:
        case TRAMBA_ORDER_LINE4:
        case TRAMBA_ORDER_HILBERT:
        case TRAMBA_ORDER_AB2:
        case TRAMBA_ORDER_AB1: return 4;
        case TRAMBA_ORDER_LINE:
        case TRAMBA_ORDER_SPIRAL:
        case TRAMBA_ORDER_DIAGONAL:
        case TRAMBA_ORDER_DS: return 8;
        default: return TRAMBA_E_ORDER;
    }
}

extern "C" int tramba_order_table(int order, int H, int W, int param, int32_t* idx_out) {
    if (!idx_out) return TRAMBA_E_NULL;
    if (H <= 0 || W <= 0 || H != W) return TRAMBA_E_SHAPE;  // square maps only (reference: row*H+col)
    if (static_cast<long long>(H) * W > (1ll << 24)) return TRAMBA_E_LIMIT;
    const int L = H * W;
    switch (order) {
        case TRAMBA_ORDER_STD: std_table(H, W, idx_out); return TRAMBA_OK;
        case TRAMBA_ORDER_DILATION:
            if (param <= 0) return TRAMBA_E_ORDER;
            dilation_table(H, W, param, idx_out);
            return TRAMBA_OK;
        case TRAMBA_ORDER_WINDOW:
            // Window.py:15 asserts 0 < ws < H; the generator also silently needs H % ws == 0
            if (param <= 0 || param >= H || H % param != 0) return TRAMBA_E_ORDER;
            window_table(H, W, param, idx_out);
            return TRAMBA_OK;
        case TRAMBA_ORDER_LINE:
            std_table(H, W, idx_out);
            return line_table(H, W, idx_out);
        case TRAMBA_ORDER_LINE4: {  // csms6s.py:219-265: the four helix lists alone
            std::vector<int32_t> t8(static_cast<size_t>(8) * L);
            const int rc = line_table(H, W, t8.data());
            if (rc != TRAMBA_OK) return rc;
            std::copy(t8.begin() + static_cast<size_t>(4) * L, t8.end(), idx_out);
            return TRAMBA_OK;
        }
        case TRAMBA_ORDER_SPIRAL:  // csms6s.py:316-339: std k0-3, cw, ccw, flip(cw), flip(ccw)
            std_table(H, W, idx_out);
            spiral_tables(H, W, idx_out, 4);
            flip_row(L, idx_out, 6, 4);
            flip_row(L, idx_out, 7, 5);
            return TRAMBA_OK;
        case TRAMBA_ORDER_HILBERT: return hilbert_table(H, W, idx_out);
        case TRAMBA_ORDER_DIAGONAL:  // csms6s.py:531-549: std k0-3, diag, anti-diag, their flips
            std_table(H, W, idx_out);
            diagonal_tables(H, W, idx_out, 4);
            flip_row(L, idx_out, 6, 4);
            flip_row(L, idx_out, 7, 5);
            return TRAMBA_OK;
        case TRAMBA_ORDER_DS: {  // csms6s.py:619-631: helix lists in slots 0-3, diagonals in 4-7
            const int rc = line_table(H, W, idx_out);  // fills slots 4-7
            if (rc != TRAMBA_OK) return rc;
            std::copy(idx_out + static_cast<size_t>(4) * L, idx_out + static_cast<size_t>(8) * L, idx_out);
            diagonal_tables(H, W, idx_out, 4);
            flip_row(L, idx_out, 6, 4);
            flip_row(L, idx_out, 7, 5);
            return TRAMBA_OK;
        }
        case TRAMBA_ORDER_AB2:  // csms6s.py:678-685: row-major twice, then both reversed
        case TRAMBA_ORDER_AB1:  // csms6s.py:713-718: row-major four times
            for (int k = 0; k < 4; ++k)
                for (int t = 0; t < L; ++t)
                    idx_out[k * L + t] = (order == TRAMBA_ORDER_AB2 && k >= 2) ? L - 1 - t : t;
            return TRAMBA_OK;
        default: return TRAMBA_E_ORDER;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory permutation schedule of a scan order (consumed by the fused backward, ss2d_fused_bwd.cu).
//
// The kernel stages an image-ordered plane with one linear TMA bulk copy and re-orders it into token order inside
// shared memory (and scatters the token-ordered du back into the image plane).  Which (pixel -> token) move a lane
// performs in which step is free, so it is chosen here, once per (order, H): a proper EDGE COLOURING of the bipartite
// multigraph {source bank = pixel mod 32} x {destination bank = token mod 32} whose edges are the L tokens.  Every
// colour class is one warp step in which all 32 source banks and all 32 destination banks are distinct: the
// permutation runs without bank conflicts for 4-byte elements in either direction (and with at most 2-way conflicts
// for 2-byte elements), for ANY order table -- column-major, windows, dilation phases, Bresenham lines with
// duplicates.  Koenig: Delta colours suffice (Delta = largest bank degree = L / 32 for permutations).
//
// Entry = (2 * pixel) | (2 * token) << 16 (byte offsets for 2-byte elements; << 1 for 4-byte), 0xFFFFFFFF = no move.
// Layout: blocks of 64 entries = two steps interleaved (entry 2 * lane + c of a block = step c of that block), since a
// thread fetches two consecutive entries with one 8-byte load.  `len` (entries per direction, a multiple of
// max(64, 2 * L / 18)) is common to all directions of the order.
// kind_out has 2K entries.  [k]: 0 = identity (token t = pixel t: no permutation pass), 1 = reversed identity,
//          2 = permutation, 3 = table with repeated pixels (the merge side needs atomic adds and a zeroed plane);
//          [K + k]: entries direction k uses (<= len, a multiple of the granule; == L for a permutation).
namespace {

int colour_tokens(const int32_t* idx, int L, std::vector<int>& colour) {
    constexpr int NB = 32;
    std::vector<int> degL(NB, 0), degR(NB, 0);
    for (int t = 0; t < L; ++t) ++degL[idx[t] & (NB - 1)], ++degR[t & (NB - 1)];
    int delta = 0;
    for (int b = 0; b < NB; ++b) delta = std::max(delta, std::max(degL[b], degR[b]));
    std::vector<int> atL(static_cast<size_t>(NB) * delta, -1), atR(static_cast<size_t>(NB) * delta, -1);
    colour.assign(L, -1);
    auto slotL = [&](int u, int c) -> int& { return atL[static_cast<size_t>(u) * delta + c]; };
    auto slotR = [&](int v, int c) -> int& { return atR[static_cast<size_t>(v) * delta + c]; };
    std::vector<int> path;
    for (int e = 0; e < L; ++e) {
        const int u = idx[e] & (NB - 1), v = e & (NB - 1);
        int a = 0, b = 0;
        while (slotL(u, a) != -1) ++a;
        while (slotR(v, b) != -1) ++b;
        if (a != b) {
            // free colour a at v: swap a <-> b along the alternating path that starts with v's a-edge
            path.clear();
            int side = 1, node = v, want = a;  // side 1 = right vertex
            for (;;) {
                const int f = side ? slotR(node, want) : slotL(node, want);
                if (f == -1) break;
                path.push_back(f);
                node = side ? (idx[f] & (NB - 1)) : (f & (NB - 1));
                side ^= 1;
                want = (want == a) ? b : a;
            }
            for (int f : path) slotL(idx[f] & (NB - 1), colour[f]) = -1, slotR(f & (NB - 1), colour[f]) = -1;
            for (int f : path) {
                colour[f] = (colour[f] == a) ? b : a;
                slotL(idx[f] & (NB - 1), colour[f]) = f, slotR(f & (NB - 1), colour[f]) = f;
            }
        }
        colour[e] = a;
        slotL(u, a) = e, slotR(v, a) = e;
    }
    return delta;
}

int table_kind(const int32_t* idx, int L) {
    bool ident = true, rev = true;
    std::vector<uint8_t> seen(L, 0);
    bool dup = false;
    for (int t = 0; t < L; ++t) {
        ident = ident && idx[t] == t;
        rev = rev && idx[t] == L - 1 - t;
        if (seen[idx[t]]) dup = true;
        seen[idx[t]] = 1;
    }
    return ident ? 0 : rev ? 1 : dup ? 3 : 2;
}

size_t sched_granule(int L) { return static_cast<size_t>(std::max(64, 2 * (L / 18))); }

}  // namespace

extern "C" size_t tramba_order_sched_len(int order, int H, int W, int param) {
    const int K = tramba_order_num_directions(order);
    if (K < 0 || H <= 0 || W <= 0 || static_cast<long long>(H) * W > 32768) return 0;
    const int L = H * W;
    std::vector<int32_t> t(static_cast<size_t>(K) * L);
    if (tramba_order_table(order, H, W, param, t.data()) != TRAMBA_OK) return 0;
    int delta = 0;
    for (int k = 0; k < K; ++k) {
        std::vector<int> deg(32, 0);
        for (int i = 0; i < L; ++i) ++deg[t[static_cast<size_t>(k) * L + i] & 31];
        for (int b = 0; b < 32; ++b) delta = std::max(delta, deg[b]);
    }
    delta = std::max(delta, (L + 31) / 32);
    const size_t g = sched_granule(L);
    return (static_cast<size_t>(delta) * 32 + g - 1) / g * g;
}

extern "C" int tramba_order_sched(int order, int H, int W, int param, uint32_t* sched_out, int32_t* kind_out) {
    if (!sched_out || !kind_out) return TRAMBA_E_NULL;
    const int K = tramba_order_num_directions(order);
    if (K < 0) return K;
    const size_t len = tramba_order_sched_len(order, H, W, param);
    if (len == 0) return TRAMBA_E_LIMIT;
    const int L = H * W;
    std::vector<int32_t> t(static_cast<size_t>(K) * L);
    const int rc = tramba_order_table(order, H, W, param, t.data());
    if (rc != TRAMBA_OK) return rc;
    std::vector<int> colour, fill;
    for (int k = 0; k < K; ++k) {
        const int32_t* idx = t.data() + static_cast<size_t>(k) * L;
        uint32_t* out = sched_out + static_cast<size_t>(k) * len;
        std::fill(out, out + len, 0xFFFFFFFFu);
        kind_out[k] = table_kind(idx, L);
        const int steps = colour_tokens(idx, L, colour);
        if (static_cast<size_t>(steps) * 32 > len) return TRAMBA_E_LIMIT;
        // entries this direction really uses (the row stride `len` is the order's maximum): a permutation needs L
        const size_t g = sched_granule(L);
        kind_out[K + k] = static_cast<int32_t>((static_cast<size_t>(steps) * 32 + g - 1) / g * g);
        fill.assign(steps, 0);
        for (int e = 0; e < L; ++e) {
            const int s = colour[e], lane = fill[s]++;
            const size_t pos = static_cast<size_t>(s >> 1) * 64 + 2 * lane + (s & 1);
            out[pos] = (static_cast<uint32_t>(idx[e]) * 2u) | (static_cast<uint32_t>(e) * 2u) << 16;
        }
    }
    return TRAMBA_OK;
}

// DCT_2D.py:37-45: W[v][j] = cos(pi*(0.5+j)*v/n)/sqrt(n) * (v ? sqrt(2) : 1), in double, stored fp32.
extern "C" void tramba_dct_basis(int n, float* w_out) {
    const double pi = 3.141592653589793;  // == math.pi
    for (int v = 0; v < n; ++v)
        for (int j = 0; j < n; ++j) {
            double b = std::cos(pi * (0.5 + j) * v / n) / std::sqrt(static_cast<double>(n));
            if (v != 0) b = b * std::sqrt(2.0);
            w_out[v * n + j] = static_cast<float>(b);
        }
}
