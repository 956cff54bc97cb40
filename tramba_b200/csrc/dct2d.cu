// DCT2D high/low split (forward + adjoint), one CTA per image plane, everything in shared
// memory.  Replaces DCT2D.forward (Models/DCT_2D.py:12-29) whose two axis layers
// (DCT_2D.py:47-53, 71-77) launch 2*H full-tensor multiply+reduce kernels; here a plane is
// read once, transformed in two in-SMEM contractions Y = Wy . X . Wx^T, and only the two
// quadrants the caller keeps are ever formed / written.
//
// This is the SIMT (fp32 FFMA) formulation: exact fp32 products, used as the numerical
// reference for the tensor-core path and as the path for sizes the tcgen05 tiles do not cover.
#include "common.cuh"

namespace tramba {

namespace {

constexpr int kDctThreads = 256;

// forward: Z = X . Wx^T (all columns), then
//   low[u][v]  = sum_h Wy[u][h]      Z[h][v]       u < H/2, v < W/2
//   high[u][v] = sum_h Wy[H/2+u][h]  Z[h][W/2+v]
__global__ void __launch_bounds__(kDctThreads) dct2d_fwd_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ wy,
                                                               const float* __restrict__ wx,
                                                               float* __restrict__ high, float* __restrict__ low,
                                                               int H, int W) {
    extern __shared__ float sm[];
    float* Xs = sm;            // [H][W]
    float* Zs = Xs + H * W;    // [H][W]
    float* WxT = Zs + H * W;   // [W][W]  WxT[w][v] = Wx[v][w]
    float* Wys = WxT + W * W;  // [H][H]  row-major
    const int tid = threadIdx.x;
    const float* xp = x + static_cast<long long>(blockIdx.x) * H * W;
    for (int e = tid; e < H * W; e += kDctThreads) Xs[e] = xp[e];
    for (int e = tid; e < W * W; e += kDctThreads) {
        const int v = e / W, w = e - v * W;
        WxT[w * W + v] = wx[e];
    }
    for (int e = tid; e < H * H; e += kDctThreads) Wys[e] = wy[e];
    __syncthreads();
    // stage 1: 2 rows x 1 column per thread step
    for (int e = tid; e < (H / 2) * W; e += kDctThreads) {
        const int h = (e / W) * 2, v = e % W;
        float a0 = 0.f, a1 = 0.f;
        for (int w = 0; w < W; ++w) {
            const float b = WxT[w * W + v];
            a0 = fmaf(Xs[h * W + w], b, a0);
            a1 = fmaf(Xs[(h + 1) * W + w], b, a1);
        }
        Zs[h * W + v] = a0;
        Zs[(h + 1) * W + v] = a1;
    }
    __syncthreads();
    const int Hh = H / 2, Wh = W / 2;
    float* lp = low + static_cast<long long>(blockIdx.x) * Hh * Wh;
    float* hp = high + static_cast<long long>(blockIdx.x) * Hh * Wh;
    for (int e = tid; e < Hh * Wh; e += kDctThreads) {
        const int u = e / Wh, v = e - u * Wh;
        float al = 0.f, ah = 0.f;
        for (int h = 0; h < H; ++h) {
            al = fmaf(Wys[u * H + h], Zs[h * W + v], al);
            ah = fmaf(Wys[(Hh + u) * H + h], Zs[h * W + Wh + v], ah);
        }
        lp[e] = al;
        hp[e] = ah;
    }
}

// adjoint: T = dY . Wx (dY zero outside the two quadrants), gx = Wy^T . T
__global__ void __launch_bounds__(kDctThreads) dct2d_bwd_kernel(const float* __restrict__ ghigh,
                                                               const float* __restrict__ glow,
                                                               const float* __restrict__ wy,
                                                               const float* __restrict__ wx,
                                                               float* __restrict__ gx, int H, int W) {
    extern __shared__ float sm[];
    const int Hh = H / 2, Wh = W / 2;
    float* Gl = sm;             // [Hh][Wh]
    float* Gh = Gl + Hh * Wh;   // [Hh][Wh]
    float* Ts = Gh + Hh * Wh;   // [H][W]
    float* Wxs = Ts + H * W;    // [W][W] row-major
    float* Wys = Wxs + W * W;   // [H][H] row-major
    const int tid = threadIdx.x;
    const float* lp = glow + static_cast<long long>(blockIdx.x) * Hh * Wh;
    const float* hp = ghigh + static_cast<long long>(blockIdx.x) * Hh * Wh;
    for (int e = tid; e < Hh * Wh; e += kDctThreads) Gl[e] = lp[e], Gh[e] = hp[e];
    for (int e = tid; e < W * W; e += kDctThreads) Wxs[e] = wx[e];
    for (int e = tid; e < H * H; e += kDctThreads) Wys[e] = wy[e];
    __syncthreads();
    for (int e = tid; e < H * W; e += kDctThreads) {
        const int u = e / W, w = e - u * W;
        float a = 0.f;
        if (u < Hh) {
            for (int v = 0; v < Wh; ++v) a = fmaf(Gl[u * Wh + v], Wxs[v * W + w], a);
        } else {
            for (int v = 0; v < Wh; ++v) a = fmaf(Gh[(u - Hh) * Wh + v], Wxs[(Wh + v) * W + w], a);
        }
        Ts[e] = a;
    }
    __syncthreads();
    float* gp = gx + static_cast<long long>(blockIdx.x) * H * W;
    for (int e = tid; e < H * W; e += kDctThreads) {
        const int h = e / W, w = e - h * W;
        float a = 0.f;
        for (int u = 0; u < H; ++u) a = fmaf(Wys[u * H + h], Ts[u * W + w], a);
        gp[e] = a;
    }
}

}  // namespace

cudaError_t launch_dct2d_fwd(const float* x, const float* wy, const float* wx, float* high, float* low,
                             int planes, int H, int W, cudaStream_t st) {
    const size_t smem = (2ull * H * W + static_cast<size_t>(W) * W + static_cast<size_t>(H) * H) * 4;
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(dct2d_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    dct2d_fwd_kernel<<<planes, kDctThreads, smem, st>>>(x, wy, wx, high, low, H, W);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_dct2d_bwd(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                             int planes, int H, int W, cudaStream_t st) {
    const size_t smem =
        (2ull * (H / 2) * (W / 2) + static_cast<size_t>(H) * W + static_cast<size_t>(W) * W + static_cast<size_t>(H) * H) * 4;
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(dct2d_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    dct2d_bwd_kernel<<<planes, kDctThreads, smem, st>>>(ghigh, glow, wy, wx, gx, H, W);
    note_launch();
    return cudaGetLastError();
}

}  // namespace tramba
