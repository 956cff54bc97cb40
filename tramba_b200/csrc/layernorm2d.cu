// LayerNorm2d (+ optional GELU epilogue) for sm_100a: layer norm over the CHANNEL axis of a
// channel-first (B, C, L) map in one pass over HBM, forward and backward.
//
// Replaces LayerNorm2d.forward (Models/modules.py:22-27: permute(0,2,3,1) -> F.layer_norm ->
// permute(0,3,1,2), i.e. two full transposed copies around the norm) and, with act = GELU, the
// out_norm -> .to(dtype) -> out_act chain that follows every cross-merge (Models/vmamba.py:271,
// 287).  SURVEY 8(f) item 2: the op on the merged y, 6-8 map-sized passes in the reference.
//
// Decomposition: CTA = one batch image x one tile of 32*PPL consecutive pixels x ALL channels.
// lane <-> PPL adjacent pixels (the contiguous axis: every warp access is one coalesced
// 64/128/256-byte row segment), warps stride over the channels.  Pass 1 accumulates the
// per-pixel moments (shifted by the pixel's channel-0 value, so E[d^2] - E[d]^2 does not cancel),
// a shared-memory reduce over the warps yields mean / rstd, pass 2 re-reads the tile (it is still
// in L1/L2: one HBM read), normalises, applies the affine map and the activation and writes the
// output in the requested dtype.  The backward has the same shape: pass 1 gathers the two per-pixel
// sums of the LN adjoint and the per-channel sums for dweight / dbias (warp-reduced over the tile's
// pixels, written as per-tile partials -- no atomics, the caller sums them), pass 2 emits dx.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace tramba {

namespace {

constexpr int kMaxWarps = 16;

template <typename T, int PPL>
__device__ __forceinline__ void ld_px(const T* __restrict__ p, float (&v)[PPL]) {
    if constexpr (PPL == 1) {
        v[0] = to_f32(*p);
    } else if constexpr (sizeof(T) == 4) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x, v[1] = t.y;
    } else {
        const uint32_t t = *reinterpret_cast<const uint32_t*>(p);
        if constexpr (DTypeOf<T>::value == TRAMBA_BF16) {
            v[0] = __uint_as_float(t << 16), v[1] = __uint_as_float(t & 0xffff0000u);
        } else {
            const __half2 h = *reinterpret_cast<const __half2*>(&t);
            v[0] = __low2float(h), v[1] = __high2float(h);
        }
    }
}
template <typename T, int PPL>
__device__ __forceinline__ void st_px(T* __restrict__ p, const float (&v)[PPL]) {
    if constexpr (PPL == 1) {
        *p = from_f32<T>(v[0]);
    } else if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else if constexpr (DTypeOf<T>::value == TRAMBA_BF16) {
        *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(v[0], v[1]);
    } else {
        *reinterpret_cast<__half2*>(p) = __floats2half2_rn(v[0], v[1]);
    }
}

// exact-form GELU (nn.GELU default, approximate='none': z * Phi(z)) and its derivative.  Phi through
// Abramowitz-Stegun 7.1.26, erf(x) = 1 - (a1 t + ... + a5 t^5) exp(-x^2), t = 1/(1 + p x) (|error| <= 1.5e-7,
// far inside the fp32 tolerance): one ex2 + one rcp + 7 FMAs instead of libdevice's erff, and the
// exponential exp(-z^2/2) it needs IS the Gaussian density of the derivative.
__device__ __forceinline__ void gelu_parts(float z, float& cdf, float& ez) {
    const float x = fabsf(z) * 0.70710678118654752f;
    const float t = rcpf(fmaf(0.3275911f, x, 1.0f));
    ez = ex2f(-0.5f * kLog2e * z * z);                       // exp(-z^2 / 2) = exp(-x^2)
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float q = 0.5f * p * t * ez;                       // (1 - erf(x)) / 2 = Phi(-|z|)
    cdf = z >= 0.f ? 1.0f - q : q;
}
__device__ __forceinline__ float gelu_f(float z) {
    float cdf, ez;
    gelu_parts(z, cdf, ez);
    return z * cdf;
}
__device__ __forceinline__ float gelu_df(float z) {
    float cdf, ez;
    gelu_parts(z, cdf, ez);
    return fmaf(z * 0.3989422804014327f, ez, cdf);
}

template <typename in_t, typename out_t, int PPL, int ACT>
__global__ void __launch_bounds__(512) ln2d_fwd_kernel(const in_t* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bs,
                                out_t* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int C,
                                int L, float eps) {
    __shared__ float red[2][kMaxWarps][32 * PPL];
    __shared__ float stat[2][32 * PPL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int b = blockIdx.y;
    const int l0 = (blockIdx.x * 32 + lane) * PPL;
    const bool ok = l0 < L;  // L % PPL == 0: a lane's pixels are all in or all out
    const in_t* xb = x + static_cast<size_t>(b) * C * L + l0;

    float k0[PPL], s1[PPL], s2[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) k0[i] = s1[i] = s2[i] = 0.f;
    if (ok) {
        ld_px<in_t, PPL>(xb, k0);
#pragma unroll 4
        for (int c = warp; c < C; c += nw) {
            float v[PPL];
            ld_px<in_t, PPL>(xb + static_cast<size_t>(c) * L, v);
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
                const float d = v[i] - k0[i];
                s1[i] += d;
                s2[i] = fmaf(d, d, s2[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PPL; ++i) red[0][warp][lane * PPL + i] = s1[i], red[1][warp][lane * PPL + i] = s2[i];
    __syncthreads();
    if (threadIdx.x < 32 * PPL) {
        float a = 0.f, q = 0.f;
        for (int ww = 0; ww < nw; ++ww) a += red[0][ww][threadIdx.x], q += red[1][ww][threadIdx.x];
        const float inv = 1.0f / static_cast<float>(C);
        const float m = a * inv;                      // mean of the shifted values
        const float var = fmaxf(fmaf(-m, m, q * inv), 0.f);
        stat[0][threadIdx.x] = m;
        stat[1][threadIdx.x] = rsqrtf(var + eps);
    }
    __syncthreads();
    if (!ok) return;
    float mu[PPL], rs[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) mu[i] = stat[0][lane * PPL + i] + k0[i], rs[i] = stat[1][lane * PPL + i];
    if (warp == 0) {
        if (mean) st_px<float, PPL>(mean + static_cast<size_t>(b) * L + l0, mu);
        if (rstd) st_px<float, PPL>(rstd + static_cast<size_t>(b) * L + l0, rs);
    }
    out_t* yb = y + static_cast<size_t>(b) * C * L + l0;
#pragma unroll 4
    for (int c = warp; c < C; c += nw) {
        float v[PPL];
        ld_px<in_t, PPL>(xb + static_cast<size_t>(c) * L, v);
        const float wc = __ldg(w + c), bc = bs ? __ldg(bs + c) : 0.f;
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            const float z = fmaf((v[i] - mu[i]) * rs[i], wc, bc);
            v[i] = ACT ? gelu_f(z) : z;
        }
        st_px<out_t, PPL>(yb + static_cast<size_t>(c) * L, v);
    }
}

// dx = rstd * (g*w - mean_c(g*w) - xhat * mean_c(g*w*xhat)),  g = gy * act'(xhat*w + b)
// dweight_part[tile][c] = sum_{pixels of the tile} g * xhat ; dbias_part[tile][c] = sum g
template <typename in_t, typename g_t, int PPL, int ACT>
__global__ void __launch_bounds__(512) ln2d_bwd_kernel(const g_t* __restrict__ gy, const in_t* __restrict__ x, const float* __restrict__ w,
                                const float* __restrict__ bs, const float* __restrict__ mean,
                                const float* __restrict__ rstd, in_t* __restrict__ gx, float* __restrict__ gw_part,
                                float* __restrict__ gb_part, int C, int L) {
    __shared__ float red[2][kMaxWarps][32 * PPL];
    __shared__ float stat[2][32 * PPL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int b = blockIdx.y;
    const int l0 = (blockIdx.x * 32 + lane) * PPL;
    const bool ok = l0 < L;
    const size_t base = static_cast<size_t>(b) * C * L + l0;
    const in_t* xb = x + base;
    const g_t* gb = gy + base;
    float mu[PPL], rs[PPL], s1[PPL], s2[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) mu[i] = 0.f, rs[i] = 0.f, s1[i] = 0.f, s2[i] = 0.f;
    if (ok) {
        ld_px<float, PPL>(mean + static_cast<size_t>(b) * L + l0, mu);
        ld_px<float, PPL>(rstd + static_cast<size_t>(b) * L + l0, rs);
    }
    float* pw = gw_part + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * C;
    float* pb = gb_part ? gb_part + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * C : nullptr;
#pragma unroll 2
    for (int c = warp; c < C; c += nw) {
        float v[PPL], g[PPL];
#pragma unroll
        for (int i = 0; i < PPL; ++i) v[i] = 0.f, g[i] = 0.f;
        if (ok) {
            ld_px<in_t, PPL>(xb + static_cast<size_t>(c) * L, v);
            ld_px<g_t, PPL>(gb + static_cast<size_t>(c) * L, g);
        }
        const float wc = __ldg(w + c), bc = bs ? __ldg(bs + c) : 0.f;
        float cw = 0.f, cb = 0.f;
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            const float xh = (v[i] - mu[i]) * rs[i];
            float gi = g[i];
            if (ACT) gi *= gelu_df(fmaf(xh, wc, bc));
            cw = fmaf(gi, xh, cw);
            cb += gi;
            const float gw = gi * wc;
            s1[i] += gw;
            s2[i] = fmaf(gw, xh, s2[i]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            cw += __shfl_xor_sync(0xffffffffu, cw, off);
            cb += __shfl_xor_sync(0xffffffffu, cb, off);
        }
        if (lane == 0) {
            pw[c] = cw;
            if (pb) pb[c] = cb;
        }
    }
#pragma unroll
    for (int i = 0; i < PPL; ++i) red[0][warp][lane * PPL + i] = s1[i], red[1][warp][lane * PPL + i] = s2[i];
    __syncthreads();
    if (threadIdx.x < 32 * PPL) {
        float a = 0.f, q = 0.f;
        for (int ww = 0; ww < nw; ++ww) a += red[0][ww][threadIdx.x], q += red[1][ww][threadIdx.x];
        const float inv = 1.0f / static_cast<float>(C);
        stat[0][threadIdx.x] = a * inv;
        stat[1][threadIdx.x] = q * inv;
    }
    __syncthreads();
    if (!ok) return;
    float m1[PPL], m2[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) m1[i] = stat[0][lane * PPL + i], m2[i] = stat[1][lane * PPL + i];
    in_t* ob = gx + base;
#pragma unroll 2
    for (int c = warp; c < C; c += nw) {
        float v[PPL], g[PPL];
        ld_px<in_t, PPL>(xb + static_cast<size_t>(c) * L, v);
        ld_px<g_t, PPL>(gb + static_cast<size_t>(c) * L, g);
        const float wc = __ldg(w + c), bc = bs ? __ldg(bs + c) : 0.f;
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            const float xh = (v[i] - mu[i]) * rs[i];
            float gi = g[i];
            if (ACT) gi *= gelu_df(fmaf(xh, wc, bc));
            v[i] = rs[i] * (fmaf(gi, wc, -m1[i]) - xh * m2[i]);
        }
        st_px<in_t, PPL>(ob + static_cast<size_t>(c) * L, v);
    }
}


// ------------------------------------------------------------------------------------------------
// Register-resident variant (the one the model's shapes take).  A thread-block CLUSTER of CS CTAs
// shares one pixel tile (64 pixels, two per lane) and splits the channels; every thread keeps its
// <= ROWS channel rows of the tile in registers (ROWS independent loads in flight per thread), the
// per-pixel sums are reduced over the warps in shared memory and over the cluster's CTAs through
// distributed shared memory, and the output is produced from the registers: exactly one HBM read
// and one write, and CS x more CTAs than pixel tiles alone would give (small maps with many
// channels -- 24x24x1024 -- would otherwise leave most SMs idle).
// ------------------------------------------------------------------------------------------------
constexpr int kRows = 16;   // channel rows per thread held in registers

constexpr int kMaxWP = 4;   // 64-pixel sub-tiles per CTA (warps side by side along the pixel axis)

// Reduce red[2][nw][64] over the WC warps that share a pixel sub-tile (warp = wc*WP + wp), then over
// the cluster's CTAs through distributed shared memory: tot[a*WP*64 + wp*64 + i].
__device__ __forceinline__ void tile_sums(cg::cluster_group& cluster, float (*red)[kMaxWarps][64], float* part,
                                          float* tot, int cs, int WP, int WC) {
    __syncthreads();
    const int n = 2 * WP * 64;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int a = idx / (WP * 64), r = idx - a * WP * 64, wp = r >> 6, i = r & 63;
        float s = 0.f;
        for (int wc = 0; wc < WC; ++wc) s += red[a][wc * WP + wp][i];
        part[idx] = s;
    }
    cluster.sync();                                    // every CTA's partials are written
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < cs; ++r) s += cluster.map_shared_rank(part, r)[idx];
        tot[idx] = s;
    }
    cluster.sync();                                    // all remote reads done (and tot visible to the CTA)
}

// Row bookkeeping shared by the two kernels.  Every thread walks kRows channel rows c = c0 + i*stride;
// rows past the tensor (i >= nvalid, warp-uniform) and lanes past the image (!ok) are handled WITHOUT
// branches: their loads are redirected to a valid address (the row / pixel is clamped) and their
// contributions are zeroed by a select, so the unrolled row loops are straight-line code.  The affine
// parameters of the warp's rows live one per lane (lanes 0..15: weight of row i, lanes 16..31: bias of row
// i - 16) and are broadcast with a shuffle where needed: one global load per thread instead of three per row.
struct LnRows {
    int c0, stride, nvalid;
    float wb;   // this lane's weight (lane < 16) or bias (lane >= 16) of row (lane & 15)
};
__device__ __forceinline__ LnRows ln_rows(const float* __restrict__ w, const float* __restrict__ bs, int C, int c0,
                                          int stride, int lane) {
    LnRows r;
    r.c0 = c0, r.stride = stride;
    r.nvalid = c0 < C ? min(kRows, (C - c0 + stride - 1) / stride) : 0;
    const int i = lane & 15, c = c0 + i * stride;
    r.wb = 0.f;
    if (i < r.nvalid) r.wb = lane < 16 ? __ldg(w + c) : (bs ? __ldg(bs + c) : 0.f);
    return r;
}

template <typename in_t, typename out_t, int ACT>
__global__ void __launch_bounds__(512) ln2d_fwd_reg_kernel(const in_t* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bs, out_t* __restrict__ y,
                                                           float* __restrict__ mean, float* __restrict__ rstd, int C,
                                                           int L, float eps, int WP) {
    __shared__ float red[2][kMaxWarps][64];
    __shared__ float part[2 * kMaxWP * 64];
    __shared__ float tot[2 * kMaxWP * 64];
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int WC = nw / WP, wp = warp % WP, wc = warp / WP;
    const int cs = gridDim.y;
    const LnRows R = ln_rows(w, bs, C, blockIdx.y * WC + wc, cs * WC, lane);
    const int b = blockIdx.z;
    const int l0 = ((blockIdx.x * WP + wp) * 32 + lane) * 2;
    const bool ok = l0 < L;
    const size_t pix = static_cast<size_t>(b) * C * L + (ok ? l0 : 0);
    const size_t rstep = static_cast<size_t>(R.stride) * L;
    const int last = max(R.nvalid - 1, 0);
    const in_t* xr = x + pix + static_cast<size_t>(min(R.c0, C - 1)) * L;
    float v[kRows][2], k0[2];
    ld_px<in_t, 2>(x + pix, k0);                                  // channel 0 of the pixel: the shift
#pragma unroll
    for (int i = 0; i < kRows; ++i) ld_px<in_t, 2>(xr + static_cast<size_t>(min(i, last)) * rstep, v[i]);
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        const float m = (i < R.nvalid && ok) ? 1.f : 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float d = (v[i][j] - k0[j]) * m;
            s1[j] += d;
            s2[j] = fmaf(d, d, s2[j]);
        }
    }
    red[0][warp][lane * 2] = s1[0], red[0][warp][lane * 2 + 1] = s1[1];
    red[1][warp][lane * 2] = s2[0], red[1][warp][lane * 2 + 1] = s2[1];
    tile_sums(cluster, red, part, tot, cs, WP, WC);
    const float inv = 1.0f / static_cast<float>(C);
    float mu[2], rs[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float m = tot[wp * 64 + lane * 2 + j] * inv;
        const float var = fmaxf(fmaf(-m, m, tot[WP * 64 + wp * 64 + lane * 2 + j] * inv), 0.f);
        mu[j] = m + k0[j];
        rs[j] = rsqrtf(var + eps);
    }
    if (ok && blockIdx.y == 0 && wc == 0) {
        if (mean) st_px<float, 2>(mean + static_cast<size_t>(b) * L + l0, mu);
        if (rstd) st_px<float, 2>(rstd + static_cast<size_t>(b) * L + l0, rs);
    }
    out_t* yr = y + pix + static_cast<size_t>(min(R.c0, C - 1)) * L;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        const float wc_ = __shfl_sync(0xffffffffu, R.wb, i), bc_ = __shfl_sync(0xffffffffu, R.wb, 16 + i);
        float o[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float z = fmaf((v[i][j] - mu[j]) * rs[j], wc_, bc_);
            o[j] = ACT ? gelu_f(z) : z;
        }
        if (i < R.nvalid && ok) st_px<out_t, 2>(yr + static_cast<size_t>(i) * rstep, o);
    }
}

// Sum each of 16 per-lane values over the 32 lanes with a transposed butterfly: 8 + 4 + 2 + 1 + 1 = 16
// shuffles instead of 16 x 5.  On return lane l holds the warp-wide total of value (l & 15).
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int h = 8; h >= 1; h >>= 1) {
        const bool up = lane & h;
#pragma unroll
        for (int j = 0; j < h; ++j) {
            const float keep = up ? v[j + h] : v[j];
            const float give = up ? v[j] : v[j + h];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, give, h);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

template <typename in_t, typename g_t, int ACT>
__global__ void __launch_bounds__(512) ln2d_bwd_reg_kernel(const g_t* __restrict__ gy, const in_t* __restrict__ x,
                                                           const float* __restrict__ w, const float* __restrict__ bs,
                                                           const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, in_t* __restrict__ gx,
                                                           float* __restrict__ gw_part, float* __restrict__ gb_part,
                                                           int C, int L, int WP) {
    __shared__ float red[2][kMaxWarps][64];
    __shared__ float part[2 * kMaxWP * 64];
    __shared__ float tot[2 * kMaxWP * 64];
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int WC = nw / WP, wp = warp % WP, wc = warp / WP;
    const int cs = gridDim.y;
    const LnRows R = ln_rows(w, bs, C, blockIdx.y * WC + wc, cs * WC, lane);
    const int b = blockIdx.z;
    const int l0 = ((blockIdx.x * WP + wp) * 32 + lane) * 2;
    const bool ok = l0 < L;
    const size_t pix = static_cast<size_t>(b) * C * L + (ok ? l0 : 0);
    const size_t rstep = static_cast<size_t>(R.stride) * L;
    const size_t row0 = pix + static_cast<size_t>(min(R.c0, C - 1)) * L;
    const int last = max(R.nvalid - 1, 0);
    float xh[kRows][2], g[kRows][2], mu[2], rs[2];
    ld_px<float, 2>(mean + static_cast<size_t>(b) * L + (ok ? l0 : 0), mu);
    ld_px<float, 2>(rstd + static_cast<size_t>(b) * L + (ok ? l0 : 0), rs);
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        const size_t o = row0 + static_cast<size_t>(min(i, last)) * rstep;
        ld_px<in_t, 2>(x + o, xh[i]);
        ld_px<g_t, 2>(gy + o, g[i]);
    }
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
    const size_t prow = static_cast<size_t>(b) * gridDim.x * WP + blockIdx.x * WP + wp;   // one row per 64-pixel tile
#pragma unroll
    for (int half = 0; half < 2; ++half) {               // rows in two batches of 8: 16 channel sums per batch
        float cwb[16];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) {
            const int i = half * 8 + ii;
            const float m = (i < R.nvalid && ok) ? 1.f : 0.f;
            const float wc_ = __shfl_sync(0xffffffffu, R.wb, i), bc_ = __shfl_sync(0xffffffffu, R.wb, 16 + i);
            float cw = 0.f, cb = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float h = (xh[i][j] - mu[j]) * rs[j];
                float gi = g[i][j] * m;
                if (ACT) gi *= gelu_df(fmaf(h, wc_, bc_));
                xh[i][j] = h, g[i][j] = gi;               // keep xhat and the activated gradient
                cw = fmaf(gi, h, cw);
                cb += gi;
                const float gw = gi * wc_;
                s1[j] += gw;
                s2[j] = fmaf(gw, h, s2[j]);
            }
            cwb[ii] = cw, cwb[8 + ii] = cb;
        }
        // lane l ends up with: l & 15 < 8 -> dweight partial of row half*8 + (l & 7); else dbias partial
        const float t = warp_sum16(cwb, lane);
        const int i = half * 8 + (lane & 7);
        if (lane < 16 && i < R.nvalid) {
            const int c = R.c0 + i * R.stride;
            if (lane < 8) gw_part[prow * C + c] = t;
            else if (gb_part) gb_part[prow * C + c] = t;
        }
    }
    red[0][warp][lane * 2] = s1[0], red[0][warp][lane * 2 + 1] = s1[1];
    red[1][warp][lane * 2] = s2[0], red[1][warp][lane * 2 + 1] = s2[1];
    tile_sums(cluster, red, part, tot, cs, WP, WC);
    const float inv = 1.0f / static_cast<float>(C);
    const float* t1 = tot + wp * 64 + lane * 2;
    const float* t2 = tot + WP * 64 + wp * 64 + lane * 2;
    const float m1[2] = {t1[0] * inv, t1[1] * inv};
    const float m2[2] = {t2[0] * inv, t2[1] * inv};
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
        const float wc_ = __shfl_sync(0xffffffffu, R.wb, i);
        float o[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) o[j] = rs[j] * (fmaf(g[i][j], wc_, -m1[j]) - xh[i][j] * m2[j]);
        if (i < R.nvalid && ok) st_px<in_t, 2>(gx + row0 + static_cast<size_t>(i) * rstep, o);
    }
}

struct LnPlan {
    int ppl, threads, tiles;
    int cs;   // > 0: register-resident cluster kernels with cs CTAs per pixel tile; 0: streaming kernels
    int wp;   // 64-pixel sub-tiles per CTA (register-resident kernels)
};
LnPlan ln_plan(int B, int C, int L) {
    const int sms = device_sms();
    LnPlan p;
    p.cs = 0, p.wp = 1;
    if (L % 2 == 0) {
        // register-resident plan: 8 warps per CTA arranged as WP 64-pixel sub-tiles (side by side: one
        // channel row of the CTA is WP*64 contiguous pixels) x WC channel groups; a cluster of cs CTAs splits
        // the channels further so that a thread holds at most kRows rows and the machine is full.  Measured:
        // cluster syncs cost more than narrow rows, so the SMALLEST cluster wins (cs <= 8, the portable limit).
        const int t64 = (L + 63) / 64;
        auto pow2_ge = [](long long v) { int r = 1; while (r < v) r *= 2; return r; };
        const long long want = 3LL * sms / 2;          // CTAs for a full machine
        int best_cs = 1 << 30, best_wp = 0, best_nw = 8;
        for (int wp = kMaxWP; wp >= 1; wp /= 2) {
            int nw = 8, wc = nw / wp;
            int cs_min = pow2_ge((C + kRows * wc - 1) / (kRows * wc));
            if (cs_min > 8 && wp == 1) {   // very wide: 16 warps
                nw = 16, wc = 16;
                cs_min = pow2_ge((C + kRows * wc - 1) / (kRows * wc));
            }
            if (cs_min > 8) continue;
            const long long base = static_cast<long long>(B) * ((t64 + wp - 1) / wp);
            int cs = cs_min;
            while (cs < 8 && base * cs < want) cs *= 2;
            while (cs > cs_min && C < cs * wc) cs /= 2;
            if (cs < best_cs) best_cs = cs, best_wp = wp, best_nw = nw;   // smallest cluster; ties keep the wider tile
        }
        if (best_wp) {
            p.cs = best_cs, p.wp = best_wp, p.ppl = 2, p.threads = best_nw * 32;
            p.tiles = ((t64 + best_wp - 1) / best_wp) * best_wp;   // rows of the dweight / dbias partials per image
            return p;
        }
    }
    // two pixels per lane (wider rows) once that still leaves >= 2 CTAs per SM
    p.ppl = (L % 2 == 0 && static_cast<long long>(B) * ((L + 63) / 64) >= 2LL * sms) ? 2 : 1;
    p.tiles = (L + 32 * p.ppl - 1) / (32 * p.ppl);
    const long long ctas = static_cast<long long>(B) * p.tiles;
    // few CTAs and many channels: more warps per CTA
    p.threads = C >= 512 ? 512 : 256;
    if (ctas >= 8LL * sms && p.threads > 256) p.threads = 256;
    return p;
}

template <typename in_t, typename out_t>
cudaError_t fwd_t(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, int B, int C, int L,
                  float eps, int act, cudaStream_t st) {
    const LnPlan p = ln_plan(B, C, L);
    const in_t* xi = static_cast<const in_t*>(x);
    out_t* yo = static_cast<out_t*>(y);
    if (p.cs > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(p.tiles / p.wp, p.cs, B);
        cfg.blockDim = dim3(p.threads);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = p.cs, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        cudaError_t e = act ? cudaLaunchKernelEx(&cfg, ln2d_fwd_reg_kernel<in_t, out_t, 1>, xi, w, b, yo, mean, rstd, C, L, eps, p.wp)
                            : cudaLaunchKernelEx(&cfg, ln2d_fwd_reg_kernel<in_t, out_t, 0>, xi, w, b, yo, mean, rstd, C, L, eps, p.wp);
        note_launch();
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    const dim3 grid(p.tiles, B);
#define LN_GO(PPL, ACT) ln2d_fwd_kernel<in_t, out_t, PPL, ACT><<<grid, p.threads, 0, st>>>(xi, w, b, yo, mean, rstd, C, L, eps)
    if (p.ppl == 2) {
        if (act) LN_GO(2, 1); else LN_GO(2, 0);
    } else {
        if (act) LN_GO(1, 1); else LN_GO(1, 0);
    }
#undef LN_GO
    note_launch();
    return cudaGetLastError();
}

template <typename in_t, typename g_t>
cudaError_t bwd_t(const void* gy, const void* x, const float* w, const float* b, const float* mean, const float* rstd,
                  void* gx, float* gw_part, float* gb_part, int B, int C, int L, int act, cudaStream_t st) {
    const LnPlan p = ln_plan(B, C, L);
    const g_t* gi = static_cast<const g_t*>(gy);
    const in_t* xi = static_cast<const in_t*>(x);
    in_t* go = static_cast<in_t*>(gx);
    if (p.cs > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(p.tiles / p.wp, p.cs, B);
        cfg.blockDim = dim3(p.threads);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1, attr[0].val.clusterDim.y = p.cs, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        cudaError_t e = act ? cudaLaunchKernelEx(&cfg, ln2d_bwd_reg_kernel<in_t, g_t, 1>, gi, xi, w, b, mean, rstd, go, gw_part, gb_part, C, L, p.wp)
                            : cudaLaunchKernelEx(&cfg, ln2d_bwd_reg_kernel<in_t, g_t, 0>, gi, xi, w, b, mean, rstd, go, gw_part, gb_part, C, L, p.wp);
        note_launch();
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    const dim3 grid(p.tiles, B);
#define LN_GO(PPL, ACT) \
    ln2d_bwd_kernel<in_t, g_t, PPL, ACT><<<grid, p.threads, 0, st>>>(gi, xi, w, b, mean, rstd, go, gw_part, gb_part, C, L)
    if (p.ppl == 2) {
        if (act) LN_GO(2, 1); else LN_GO(2, 0);
    } else {
        if (act) LN_GO(1, 1); else LN_GO(1, 0);
    }
#undef LN_GO
    note_launch();
    return cudaGetLastError();
}

}  // namespace

int layernorm2d_tiles(int B, int C, int L) { return ln_plan(B, C, L).tiles; }

cudaError_t launch_layernorm2d_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd,
                                   int B, int C, int L, float eps, int act, int in_dtype, int out_dtype,
                                   cudaStream_t st) {
    if (in_dtype == TRAMBA_F32 && out_dtype == TRAMBA_F32)
        return fwd_t<float, float>(x, w, b, y, mean, rstd, B, C, L, eps, act, st);
    if (in_dtype == TRAMBA_F32 && out_dtype == TRAMBA_BF16)
        return fwd_t<float, __nv_bfloat16>(x, w, b, y, mean, rstd, B, C, L, eps, act, st);
    if (in_dtype == TRAMBA_BF16 && out_dtype == TRAMBA_BF16)
        return fwd_t<__nv_bfloat16, __nv_bfloat16>(x, w, b, y, mean, rstd, B, C, L, eps, act, st);
    if (in_dtype == TRAMBA_BF16 && out_dtype == TRAMBA_F32)
        return fwd_t<__nv_bfloat16, float>(x, w, b, y, mean, rstd, B, C, L, eps, act, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_layernorm2d_bwd(const void* gy, const void* x, const float* w, const float* b, const float* mean,
                                   const float* rstd, void* gx, float* gw_part, float* gb_part, int B, int C, int L,
                                   int act, int in_dtype, int g_dtype, cudaStream_t st) {
    if (in_dtype == TRAMBA_F32 && g_dtype == TRAMBA_F32)
        return bwd_t<float, float>(gy, x, w, b, mean, rstd, gx, gw_part, gb_part, B, C, L, act, st);
    if (in_dtype == TRAMBA_F32 && g_dtype == TRAMBA_BF16)
        return bwd_t<float, __nv_bfloat16>(gy, x, w, b, mean, rstd, gx, gw_part, gb_part, B, C, L, act, st);
    if (in_dtype == TRAMBA_BF16 && g_dtype == TRAMBA_BF16)
        return bwd_t<__nv_bfloat16, __nv_bfloat16>(gy, x, w, b, mean, rstd, gx, gw_part, gb_part, B, C, L, act, st);
    if (in_dtype == TRAMBA_BF16 && g_dtype == TRAMBA_F32)
        return bwd_t<__nv_bfloat16, float>(gy, x, w, b, mean, rstd, gx, gw_part, gb_part, B, C, L, act, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
