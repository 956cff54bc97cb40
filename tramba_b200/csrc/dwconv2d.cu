// Depthwise k x k convolution (+ bias, + optional SiLU) on channel-first planes for sm_100a,
// forward and backward (dgrad + wgrad + dbias) -- the producer of the SS2D core's input:
//     x = act(conv2d(in_proj(x)))          Models/vmamba.py:281-285 (d_conv = 3, groups = d_inner)
// SURVEY 8(f) item 3.  cuDNN runs these as grouped direct convolutions that take 0.6-1 ms per call
// at 96x96; a depthwise conv is plane-local (like every other stage of the path), so one CTA
// stages whole planes in shared memory (fp32, zero halo) and each thread slides a 4-wide x RH-tall
// register tile down the plane: one shared-memory read per input value, k*k FMAs per output.
//
//   forward : y = act(conv(x, w) + b)
//   backward: g  = dy * act'(conv(x, w) + b)      (pre-activation recomputed, never stored)
//             dx = conv(g, flip(w)) ; dw[c] = sum_{b,h,w} g * shifted x ; db[c] = sum g
//             dw / db leave as per-(b, c) partial sums, one row of k*k + 1 floats per plane (the caller adds
//             over b; no atomics)
#include <algorithm>

#include "common.cuh"

namespace tramba {

namespace {

constexpr int TW = 4;       // outputs per thread along W
constexpr int RH = 8;       // outputs per thread along H
constexpr int kHaloL = 4;   // left halo in shared memory (>= k/2, keeps 16-byte alignment)

struct DwGeom {
    int H, W;
    int tw, th;        // thread tiles per plane along W / H
    int rows, ld;      // shared plane: rows = th*RH + k - 1, ld = kHaloL + tw*TW + 4 floats
    int tpp;           // threads per plane, padded: multiple of 32, or a power of two below 32
    int ppc;           // planes per CTA iteration
    int planes;        // B * C
    int C;
    unsigned mg_per, mg_cw;   // ceil(2^32 / d) for d = H*(W/4) and d = W/4: e / d == umulhi(e, mg) for e*d < 2^32
};

__device__ __forceinline__ int fast_div(int e, unsigned mg, int d) {
    return d == 1 ? e : static_cast<int>(__umulhi(static_cast<unsigned>(e), mg));
}

// Stage `np` consecutive global planes (contiguous in memory) into their shared planes: interior
// element (h, w) of plane pl lands at dst[pl*psz + (h + pad)*ld + kHaloL + w]; the halo stays zero
// (cleared once per CTA).  One flat, fully coalesced sweep over all planes, U 16/8-byte loads in flight per
// thread before the first shared store (U = 1 for the register-bound 7x7 kernels).
template <typename T, int U>
__device__ __forceinline__ void stage_planes(float* __restrict__ dst, const T* __restrict__ src, int np, int psz,
                                             const DwGeom& g, int pad, int tid, int nthr) {
    const int W = g.W, H = g.H;
    if ((W & 3) == 0) {
        const int cw = W >> 2, per = H * cw, total = np * per;
        for (int e0 = tid; e0 < total; e0 += U * nthr) {
            float v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * nthr;
                if (e < total) load_vec<4, T>(src + static_cast<size_t>(e) * 4, v[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * nthr;
                if (e < total) {
                    const int pl = fast_div(e, g.mg_per, per), r = e - pl * per;
                    const int h = fast_div(r, g.mg_cw, cw), c = r - h * cw;
                    *reinterpret_cast<float4*>(dst + pl * psz + (h + pad) * g.ld + kHaloL + c * 4) =
                        make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                }
            }
        }
    } else {
        const int per = H * W, total = np * per;
        for (int e = tid; e < total; e += nthr) {
            const int pl = e / per, r = e - pl * per;
            const int h = r / W, w = r - h * W;
            dst[pl * psz + (h + pad) * g.ld + kHaloL + w] = to_f32(src[e]);
        }
    }
}

__device__ __forceinline__ float silu_f(float z) { return z * rcpf(1.0f + ex2f(-z * kLog2e)); }
__device__ __forceinline__ float silu_df(float z) {
    const float s = rcpf(1.0f + ex2f(-z * kLog2e));
    return s * fmaf(z, 1.0f - s, 1.0f);
}

// acc[o][j] (o < RH, j < TW) += sum_{i,jj} wt[i*K + jj] * plane[(h0 + o + i) * ld + c0 + j + jj]
// where plane row 0 = image row -pad and c0 = kHaloL - pad + w0 (the tile's first tap column).
template <int K>
__device__ __forceinline__ void conv_tile(float (&acc)[RH][TW], const float* __restrict__ plane, int ld, int h0,
                                          int w0, const float (&wt)[K * K]) {
    constexpr int PAD = K / 2;
    // aligned window: columns [w0, w0 + 12) of the padded row cover taps w0 + 4 - PAD .. w0 + 7 + PAD
    const float* base = plane + h0 * ld + w0;
#pragma unroll
    for (int r = 0; r < RH + K - 1; ++r) {
        float in[12];
        const float4 a = *reinterpret_cast<const float4*>(base + r * ld);
        const float4 b = *reinterpret_cast<const float4*>(base + r * ld + 4);
        const float4 c = *reinterpret_cast<const float4*>(base + r * ld + 8);
        in[0] = a.x, in[1] = a.y, in[2] = a.z, in[3] = a.w, in[4] = b.x, in[5] = b.y, in[6] = b.z, in[7] = b.w;
        in[8] = c.x, in[9] = c.y, in[10] = c.z, in[11] = c.w;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int o = r - i;
            if (o >= 0 && o < RH) {
#pragma unroll
                for (int jj = 0; jj < K; ++jj)
#pragma unroll
                    for (int j = 0; j < TW; ++j) acc[o][j] = fmaf(wt[i * K + jj], in[kHaloL - PAD + j + jj], acc[o][j]);
            }
        }
    }
}

template <typename T>
__device__ __forceinline__ void store_row(T* __restrict__ dst, const float (&v)[TW], int w0, int W) {
    if (w0 + TW <= W && (W & 3) == 0) {
        store_vec<4, T>(dst + w0, v);
    } else {
#pragma unroll
        for (int j = 0; j < TW; ++j)
            if (w0 + j < W) dst[w0 + j] = from_f32<T>(v[j]);
    }
}

template <typename T, int K, int ACT>
__global__ void __launch_bounds__(320, 2) dwconv_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wgt,
                                                         const float* __restrict__ bias, T* __restrict__ y,
                                                         const DwGeom g) {
    extern __shared__ __align__(16) float sm[];
    constexpr int PAD = K / 2;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tpp = g.tpp;                       // threads per plane (padded; t >= tw*th idles)
    const int psz = g.rows * g.ld;
    float* swt = sm + g.ppc * psz;               // [ppc][K*K + 1]
    for (int e = tid; e < g.ppc * psz; e += nthr) sm[e] = 0.f;   // halos stay zero for the CTA's life
    const int q = tid / tpp, t = tid - q * tpp;
    const int ty = t / g.tw, tx = t - ty * g.tw;
    const int h0 = ty * RH, w0 = tx * TW;
    const int groups = (g.planes + g.ppc - 1) / g.ppc;
    for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        __syncthreads();                          // previous iteration's readers are done
        const int p0 = grp * g.ppc;
        const int np = min(g.ppc, g.planes - p0);
        stage_planes<T, (K == 7 ? 1 : 4)>(sm, x + static_cast<size_t>(p0) * g.H * g.W, np, psz, g, PAD, tid, nthr);
        for (int e = tid; e < np * (K * K + 1); e += nthr) {
            const int pl = e / (K * K + 1), i = e - pl * (K * K + 1);
            const int c = (p0 + pl) % g.C;
            swt[e] = i < K * K ? __ldg(wgt + c * K * K + i) : (bias ? __ldg(bias + c) : 0.f);
        }
        __syncthreads();
        if (q < np && t < g.tw * g.th) {
            float wt[K * K];
#pragma unroll
            for (int i = 0; i < K * K; ++i) wt[i] = swt[q * (K * K + 1) + i];
            const float bc = swt[q * (K * K + 1) + K * K];
            float acc[RH][TW];
#pragma unroll
            for (int o = 0; o < RH; ++o)
#pragma unroll
                for (int j = 0; j < TW; ++j) acc[o][j] = bc;
            conv_tile<K>(acc, sm + q * psz, g.ld, h0, w0, wt);
            T* yp = y + static_cast<size_t>(p0 + q) * g.H * g.W;
#pragma unroll
            for (int o = 0; o < RH; ++o) {
                if (h0 + o < g.H) {
                    if (ACT) {
#pragma unroll
                        for (int j = 0; j < TW; ++j) acc[o][j] = silu_f(acc[o][j]);
                    }
                    store_row<T>(yp + static_cast<size_t>(h0 + o) * g.W, acc[o], w0, g.W);
                }
            }
        }
    }
}

// Backward.  Shared: plane X (input, zero halo) and plane G (dy, scaled in place to g = dy * act'(pre); zero halo).
template <typename T, int K, int ACT>
__global__ void __launch_bounds__(320, 2) dwconv_bwd_kernel(const T* __restrict__ x, const float* __restrict__ wgt,
                                                         const float* __restrict__ bias, const T* __restrict__ dy,
                                                         T* __restrict__ dx, float* __restrict__ dw_part,
                                                         const DwGeom g) {
    extern __shared__ __align__(16) float sm[];
    constexpr int PAD = K / 2;
    constexpr int KK = K * K;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int tpp = g.tpp;
    const int psz = g.rows * g.ld;
    float* sX = sm;                              // [ppc][psz]
    float* sG = sm + g.ppc * psz;                // [ppc][psz]
    float* swt = sG + g.ppc * psz;               // [ppc][KK + 1]
    float* sred = swt + g.ppc * (KK + 1);        // [nwarps][KK + 1]
    for (int e = tid; e < 2 * g.ppc * psz; e += nthr) sm[e] = 0.f;
    const int q = tid / tpp, t = tid - q * tpp;
    const int ty = t / g.tw, tx = t - ty * g.tw;
    const int h0 = ty * RH, w0 = tx * TW;
    const int nwarps = nthr >> 5;
    const int groups = (g.planes + g.ppc - 1) / g.ppc;
    for (int grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        __syncthreads();
        const int p0 = grp * g.ppc;
        const int np = min(g.ppc, g.planes - p0);
        stage_planes<T, (K == 7 ? 1 : 4)>(sX, x + static_cast<size_t>(p0) * g.H * g.W, np, psz, g, PAD, tid, nthr);
        stage_planes<T, (K == 7 ? 1 : 4)>(sG, dy + static_cast<size_t>(p0) * g.H * g.W, np, psz, g, PAD, tid, nthr);
        for (int e = tid; e < np * (KK + 1); e += nthr) {
            const int pl = e / (KK + 1), i = e - pl * (KK + 1);
            const int c = (p0 + pl) % g.C;
            swt[e] = i < KK ? __ldg(wgt + c * KK + i) : (bias ? __ldg(bias + c) : 0.f);
        }
        __syncthreads();
        const bool live = q < np && t < g.tw * g.th;
        float gt[RH][TW];                         // this thread's tile of g
#pragma unroll
        for (int o = 0; o < RH; ++o)
#pragma unroll
            for (int j = 0; j < TW; ++j) gt[o][j] = 0.f;
        if (ACT) {
            // phase 1: recompute the pre-activation on the tile, g = dy * silu'(pre) -> shared plane G
            if (live) {
                float wt[KK];
#pragma unroll
                for (int i = 0; i < KK; ++i) wt[i] = swt[q * (KK + 1) + i];
                const float bc = swt[q * (KK + 1) + KK];
                float acc[RH][TW];
#pragma unroll
                for (int o = 0; o < RH; ++o)
#pragma unroll
                    for (int j = 0; j < TW; ++j) acc[o][j] = bc;
                conv_tile<K>(acc, sX + q * psz, g.ld, h0, w0, wt);
#pragma unroll
                for (int o = 0; o < RH; ++o) {
                    if (h0 + o < g.H) {   // dy was staged into plane G with x: scale it in place
                        float4* gp = reinterpret_cast<float4*>(sG + q * psz + (h0 + o + PAD) * g.ld + kHaloL + w0);
                        const float4 d = *gp;
                        gt[o][0] = d.x * silu_df(acc[o][0]), gt[o][1] = d.y * silu_df(acc[o][1]);
                        gt[o][2] = d.z * silu_df(acc[o][2]), gt[o][3] = d.w * silu_df(acc[o][3]);
                        *gp = make_float4(gt[o][0], gt[o][1], gt[o][2], gt[o][3]);
                    }
                }
            }
            __syncthreads();
        } else if (live) {
#pragma unroll
            for (int o = 0; o < RH; ++o) {
                const float4 v = *reinterpret_cast<const float4*>(sG + q * psz + (h0 + o + PAD) * g.ld + kHaloL + w0);
                gt[o][0] = v.x, gt[o][1] = v.y, gt[o][2] = v.z, gt[o][3] = v.w;
            }
        }
        // phase 2a: wgrad partials  dw[i][jj] += sum_{o,j} g[o][j] * x[h0+o+i-PAD][w0+j+jj-PAD]; db += sum g
        {
            float dwv[KK], dbv = 0.f;
#pragma unroll
            for (int i = 0; i < KK; ++i) dwv[i] = 0.f;
            if (live) {
                const float* base = sX + q * psz + h0 * g.ld + w0;
#pragma unroll
                for (int r = 0; r < RH + K - 1; ++r) {
                    float in[12];
                    const float4 a = *reinterpret_cast<const float4*>(base + r * g.ld);
                    const float4 b = *reinterpret_cast<const float4*>(base + r * g.ld + 4);
                    const float4 c = *reinterpret_cast<const float4*>(base + r * g.ld + 8);
                    in[0] = a.x, in[1] = a.y, in[2] = a.z, in[3] = a.w, in[4] = b.x, in[5] = b.y, in[6] = b.z;
                    in[7] = b.w, in[8] = c.x, in[9] = c.y, in[10] = c.z, in[11] = c.w;
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const int o = r - i;
                        if (o >= 0 && o < RH) {
#pragma unroll
                            for (int jj = 0; jj < K; ++jj)
#pragma unroll
                                for (int j = 0; j < TW; ++j)
                                    dwv[i * K + jj] = fmaf(gt[o][j], in[kHaloL - PAD + j + jj], dwv[i * K + jj]);
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < RH; ++o)
#pragma unroll
                    for (int j = 0; j < TW; ++j) dbv += gt[o][j];
            }
            // reduce over the threads of a plane: tpp is a multiple of 32 (whole warps per plane) or a power
            // of two below 32 (a segmented xor-reduce of width tpp is exact)
            const int width = tpp < 32 ? tpp : 32;
#pragma unroll
            for (int i = 0; i < KK; ++i) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                    if (off < width) dwv[i] += __shfl_xor_sync(0xffffffffu, dwv[i], off, 32);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
                if (off < width) dbv += __shfl_xor_sync(0xffffffffu, dbv, off, 32);
            if (tpp >= 32) {
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < KK; ++i) sred[warp * (KK + 1) + i] = dwv[i];
                    sred[warp * (KK + 1) + KK] = dbv;
                }
                __syncthreads();
                // plane pl owns warps [pl*tpp/32, (pl+1)*tpp/32)
                const int wpp = tpp >> 5;
                for (int e = tid; e < np * (KK + 1); e += nthr) {
                    const int pl = e / (KK + 1), i = e - pl * (KK + 1);
                    float s = 0.f;
                    for (int ww = 0; ww < wpp; ++ww) s += sred[(pl * wpp + ww) * (KK + 1) + i];
                    dw_part[static_cast<size_t>(p0 + pl) * (KK + 1) + i] = s;
                }
            } else if (q < np && t == 0) {
#pragma unroll
                for (int i = 0; i < KK; ++i) dw_part[static_cast<size_t>(p0 + q) * (KK + 1) + i] = dwv[i];
                dw_part[static_cast<size_t>(p0 + q) * (KK + 1) + KK] = dbv;
            }
        }
        // phase 2b: dgrad  dx = conv(g, flip(w))
        if (live) {
            float wf[KK];
#pragma unroll
            for (int i = 0; i < KK; ++i) wf[i] = swt[q * (KK + 1) + (KK - 1 - i)];
            float acc[RH][TW];
#pragma unroll
            for (int o = 0; o < RH; ++o)
#pragma unroll
                for (int j = 0; j < TW; ++j) acc[o][j] = 0.f;
            conv_tile<K>(acc, sG + q * psz, g.ld, h0, w0, wf);
            T* dxp = dx + static_cast<size_t>(p0 + q) * g.H * g.W;
#pragma unroll
            for (int o = 0; o < RH; ++o)
                if (h0 + o < g.H) store_row<T>(dxp + static_cast<size_t>(h0 + o) * g.W, acc[o], w0, g.W);
        }
    }
}

struct DwPlan {
    DwGeom g;
    int threads;
    size_t smem_fwd, smem_bwd;
    bool ok;
};

DwPlan dw_plan(int B, int C, int H, int W, int K) {
    DwPlan p{};
    DwGeom& g = p.g;
    g.H = H, g.W = W, g.C = C, g.planes = B * C;
    g.tw = (W + TW - 1) / TW;
    g.th = (H + RH - 1) / RH;
    {
        const unsigned long long cw = W / 4 > 0 ? W / 4 : 1, per = static_cast<unsigned long long>(H) * cw;
        g.mg_cw = static_cast<unsigned>(((1ull << 32) + cw - 1) / cw);
        g.mg_per = static_cast<unsigned>(((1ull << 32) + per - 1) / per);
    }
    g.rows = g.th * RH + K - 1;
    g.ld = kHaloL + g.tw * TW + 4;
    int tpp = g.tw * g.th;
    p.ok = false;
    if (tpp > 320) return p;                       // planes up to ~100 x 100
    // pad the threads of a plane to whole warps (or to a power of two below 32) so that the wgrad
    // reduction never mixes planes inside a shuffle segment; planes per CTA fill ~256 threads
    if (tpp >= 32) tpp = (tpp + 31) / 32 * 32;
    else { int v = 1; while (v < tpp) v <<= 1; tpp = v; }
    g.tpp = tpp;
    int ppc = 256 / tpp;
    if (ppc > 16) ppc = 16;
    if (ppc < 1) ppc = 1;
    const size_t cap = 200 * 1024;
    auto smem_f = [&](int n) { return (static_cast<size_t>(n) * g.rows * g.ld + n * (K * K + 1)) * 4; };
    auto smem_b = [&](int n) {
        return (2ull * n * g.rows * g.ld + n * (K * K + 1) + ((n * tpp + 31) / 32 + 1) * (K * K + 1)) * 4;
    };
    while (ppc > 1 && smem_b(ppc) > cap) --ppc;
    if (smem_b(ppc) > cap) return p;
    g.ppc = ppc;
    p.threads = ((ppc * tpp + 31) / 32) * 32;
    if (p.threads > 320) return p;
    p.smem_fwd = smem_f(ppc);
    p.smem_bwd = smem_b(ppc);
    p.ok = true;
    return p;
}

int sm_count() {
    const int sms = device_sms();
    return sms;
}

template <typename T, int K>
cudaError_t fwd_k(const void* x, const float* w, const float* b, void* y, const DwPlan& p, int act, cudaStream_t st) {
    const int groups = (p.g.planes + p.g.ppc - 1) / p.g.ppc;
    // co-resident CTAs per SM: shared memory, then registers (<= 102 per thread under the launch bounds)
    const int per_sm = static_cast<int>(std::max<size_t>(
        1, std::min<size_t>({6, (200 * 1024) / (p.smem_fwd + 1024), static_cast<size_t>(65536 / (104 * p.threads))})));
    const int grid = std::min(groups, sm_count() * per_sm);
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<grid, p.threads, p.smem_fwd, st>>>(static_cast<const T*>(x), w, b, static_cast<T*>(y), p.g);
        return cudaGetLastError();
    };
    cudaError_t e = act ? go(dwconv_fwd_kernel<T, K, 1>) : go(dwconv_fwd_kernel<T, K, 0>);
    note_launch();
    return e;
}

template <typename T, int K>
cudaError_t bwd_k(const void* x, const float* w, const float* b, const void* dy, void* dx, float* dwp,
                  const DwPlan& p, int act, cudaStream_t st) {
    const int groups = (p.g.planes + p.g.ppc - 1) / p.g.ppc;
    const int per_sm = static_cast<int>(std::max<size_t>(
        1, std::min<size_t>({6, (200 * 1024) / (p.smem_bwd + 1024), static_cast<size_t>(65536 / (104 * p.threads))})));
    const int grid = std::min(groups, sm_count() * per_sm);
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<grid, p.threads, p.smem_bwd, st>>>(static_cast<const T*>(x), w, b, static_cast<const T*>(dy),
                                                  static_cast<T*>(dx), dwp, p.g);
        return cudaGetLastError();
    };
    cudaError_t e = act ? go(dwconv_bwd_kernel<T, K, 1>) : go(dwconv_bwd_kernel<T, K, 0>);
    note_launch();
    return e;
}

}  // namespace

bool dwconv2d_supported(int B, int C, int H, int W, int K) {
    if (K != 3 && K != 5 && K != 7) return false;
    return dw_plan(B, C, H, W, K).ok;
}

cudaError_t launch_dwconv2d_fwd(const void* x, const float* w, const float* b, void* y, int B, int C, int H, int W,
                                int K, int act, int dtype, cudaStream_t st) {
    const DwPlan p = dw_plan(B, C, H, W, K);
    if (!p.ok) return cudaErrorInvalidConfiguration;
#define DW_GO(T)                                                        \
    switch (K) {                                                        \
        case 3: return fwd_k<T, 3>(x, w, b, y, p, act, st);             \
        case 5: return fwd_k<T, 5>(x, w, b, y, p, act, st);             \
        case 7: return fwd_k<T, 7>(x, w, b, y, p, act, st);             \
        default: return cudaErrorInvalidValue;                          \
    }
    if (dtype == TRAMBA_F32) { DW_GO(float) }
    if (dtype == TRAMBA_BF16) { DW_GO(__nv_bfloat16) }
#undef DW_GO
    return cudaErrorInvalidValue;
}

cudaError_t launch_dwconv2d_bwd(const void* x, const float* w, const float* b, const void* dy, void* dx, float* dwp,
                                int B, int C, int H, int W, int K, int act, int dtype, cudaStream_t st) {
    const DwPlan p = dw_plan(B, C, H, W, K);
    if (!p.ok) return cudaErrorInvalidConfiguration;
#define DW_GO(T)                                                                    \
    switch (K) {                                                                    \
        case 3: return bwd_k<T, 3>(x, w, b, dy, dx, dwp, p, act, st);          \
        case 5: return bwd_k<T, 5>(x, w, b, dy, dx, dwp, p, act, st);          \
        case 7: return bwd_k<T, 7>(x, w, b, dy, dx, dwp, p, act, st);          \
        default: return cudaErrorInvalidValue;                                      \
    }
    if (dtype == TRAMBA_F32) { DW_GO(float) }
    if (dtype == TRAMBA_BF16) { DW_GO(__nv_bfloat16) }
#undef DW_GO
    return cudaErrorInvalidValue;
}

}  // namespace tramba
