// Host-side scan-order table generators (no GPU).  B200-side replacement of the reference's
// Python generators: Models/SS2D/Dilation.py:3-45, Models/SS2D/Window.py:3-35,
// Models/SS2D/SpiralLine.py:3-82 and the standard unfold of Models/SS2D/csms6s.py:13-31.
// All tables are flat pixel indices row*W + col of a square map; closed forms where the order
// is a permutation, an integer Bresenham walk for the helix lines.
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

}  // namespace

extern "C" int tramba_order_num_directions(int order) {
    switch (order) {
        case TRAMBA_ORDER_STD:
        case TRAMBA_ORDER_DILATION:
        case TRAMBA_ORDER_WINDOW: return 4;
        case TRAMBA_ORDER_LINE: return 8;
        default: return TRAMBA_E_ORDER;
    }
}

extern "C" int tramba_order_table(int order, int H, int W, int param, int32_t* idx_out) {
    if (!idx_out) return TRAMBA_E_NULL;
    if (H <= 0 || W <= 0 || H != W) return TRAMBA_E_SHAPE;  // square maps only (reference: row*H+col)
    if (static_cast<long long>(H) * W > (1ll << 24)) return TRAMBA_E_LIMIT;
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
        default: return TRAMBA_E_ORDER;
    }
}

// DCT_2D.py:37-45: W[v][j] = cos(pi*(0.5+j)*v/n)/sqrt(n) * (v ? sqrt(2) : 1), in double, stored fp32.
#include <cmath>
extern "C" void tramba_dct_basis(int n, float* w_out) {
    const double pi = 3.141592653589793;  // == math.pi
    for (int v = 0; v < n; ++v)
        for (int j = 0; j < n; ++j) {
            double b = std::cos(pi * (0.5 + j) * v / n) / std::sqrt(static_cast<double>(n));
            if (v != 0) b = b * std::sqrt(2.0);
            w_out[v * n + j] = static_cast<float>(b);
        }
}
