// Cross-scan (gather by scan order) and cross-merge (its exact transpose) for sm_100a.
// Replaces the reference's CrossScan*/CrossMerge* autograd functions
// (Models/SS2D/csms6s.py:13-216) and their index_select / scatter_add_ helpers
// (Dilation.py:48-96, Window.py:38-86, SpiralLine.py:85-133).
//
// Both kernels keep HBM traffic at the algorithmic minimum -- each element of x / xs / ys / y
// crosses HBM once, in fully coalesced 16-byte accesses -- by doing the permutation inside
// shared memory: an image plane (or one direction's token row) is staged once, and the
// irregular side of the access (column-major, window, dilated, Bresenham) is an LDS.  Planes
// are stored skewed (one pad word per 32) so that the power-of-two strides of column-major
// orders spread over all banks.  The merge is a gather through the inverse (CSR) table, so it
// is deterministic and needs no atomics even where the helix order hits a pixel many times.
#include "common.cuh"

namespace tramba {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int skew(int e) { return e + (e >> 5); }
inline int skewed_len(int L) { return L + (L >> 5) + 1; }

// Stage `np` consecutive rows of length L (global, contiguous) into skewed shared planes.
template <typename T, bool VEC, int NTH>
__device__ __forceinline__ void stage_rows(float* __restrict__ sp, int Lp, const T* __restrict__ g,
                                           long long row_stride, int np, int L) {
    for (int pl = 0; pl < np; ++pl) {
        const T* src = g + pl * row_stride;
        float* dst = sp + pl * Lp;
        if constexpr (VEC) {
            for (int e = threadIdx.x * 4; e < L; e += NTH * 4) {
                float v[4];
                load_vec<4>(src + e, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[skew(e + j)] = v[j];
            }
        } else {
            for (int e = threadIdx.x; e < L; e += NTH) dst[skew(e)] = to_f32(src[e]);
        }
    }
}

// ------------------------------------------------------------------------------- scan
// NTH threads per CTA: 256, or 1024 where a CTA's planes leave room for only one or two CTAs per SM (192 x 192 maps:
// 152 KB per plane) -- the gather is latency-bound and needs the warps.
template <typename T, bool VEC, int NTH>
__global__ void __launch_bounds__(NTH) cross_scan_kernel(const T* __restrict__ x, T* __restrict__ xs,
                                                              const int32_t* __restrict__ table, int K, int C,
                                                              int L, int planes, int ppb) {
    extern __shared__ float sp[];
    const int Lp = L + (L >> 5) + 1;
    const int plane0 = blockIdx.x * ppb;
    const int np = min(ppb, planes - plane0);
    stage_rows<T, VEC, NTH>(sp, Lp, x + static_cast<long long>(plane0) * L, L, np, L);
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const int32_t* tk = table + static_cast<long long>(k) * L;
        if constexpr (VEC) {
            for (int t = threadIdx.x * 4; t < L; t += NTH * 4) {
                const int4 id = *reinterpret_cast<const int4*>(tk + t);
                const int p0 = skew(id.x), p1 = skew(id.y), p2 = skew(id.z), p3 = skew(id.w);
                for (int pl = 0; pl < np; ++pl) {
                    const int plane = plane0 + pl;
                    const int b = plane / C, c = plane - b * C;
                    const float* s = sp + pl * Lp;
                    const float v[4] = {s[p0], s[p1], s[p2], s[p3]};
                    store_vec<4>(xs + ((static_cast<long long>(b) * K + k) * C + c) * L + t, v);
                }
            }
        } else {
            for (int t = threadIdx.x; t < L; t += NTH) {
                const int p0 = skew(tk[t]);
                for (int pl = 0; pl < np; ++pl) {
                    const int plane = plane0 + pl;
                    const int b = plane / C, c = plane - b * C;
                    xs[((static_cast<long long>(b) * K + k) * C + c) * L + t] = from_f32<T>(sp[pl * Lp + p0]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------- merge
struct MergeCfg {
    int klo, khi;  // directions accumulated sequentially from zero through the CSR table
    int has_std;   // directions 0..3 folded as (ys0 + flip ys2) + T(ys1 + flip ys3)
    int perm;      // sequential directions are permutations (one token per pixel)
};

// Staged merge: per direction, the token rows of `np` planes are staged in shared memory and
// every thread gathers the tokens of its own pixels (p = tid + i*blockDim) from them.
template <typename T, bool VEC, int NTH>
__global__ void __launch_bounds__(NTH) cross_merge_staged_kernel(const T* __restrict__ ys, T* __restrict__ y,
                                                                      const int32_t* __restrict__ plan, MergeCfg cfg,
                                                                      int K, int D, int L, int planes, int ppb) {
    extern __shared__ float sm[];
    const int Lp = L + (L >> 5) + 1;
    float* row = sm;                                       // [ppb][Lp] skewed token rows
    float* acc = row + static_cast<size_t>(ppb) * Lp;      // [ppb][L]  sequential part (if any)
    float* acc1 = acc + (cfg.khi > cfg.klo ? static_cast<size_t>(ppb) * L : 0);  // [ppb][L] std T-part
    const int32_t* ent = plan + static_cast<long long>(K) * L;
    const int32_t* off = plan + 2ll * K * L;
    const int plane0 = blockIdx.x * ppb;
    const int np = min(ppb, planes - plane0);
    const int b0 = plane0 / D, d0 = plane0 - b0 * D;  // host guarantees a block never straddles b
    const long long kstride = static_cast<long long>(D) * L;
    const T* ysb = ys + (static_cast<long long>(b0) * K * D + d0) * L;  // (b0, k=0, d0, 0)

    for (int k = cfg.klo; k < cfg.khi; ++k) {
        stage_rows<T, VEC, NTH>(row, Lp, ysb + k * kstride, L, np, L);
        __syncthreads();
        const int32_t* ek = ent + static_cast<long long>(k) * L;
        const int32_t* ok = off + static_cast<long long>(k) * (L + 1);
        for (int p = threadIdx.x; p < L; p += NTH) {
            int j0 = p, j1 = p + 1;
            if (!cfg.perm) j0 = ok[p], j1 = ok[p + 1];
            for (int pl = 0; pl < np; ++pl) {
                float s = (k == cfg.klo) ? 0.f : acc[pl * L + p];
                for (int j = j0; j < j1; ++j) s += row[pl * Lp + skew(ek[j])];
                acc[pl * L + p] = s;
            }
        }
        __syncthreads();
    }
    if (cfg.has_std) {
        // T-part: ys1 gathered through inverse(k1), plus ys3 through inverse(k3)
        for (int kk = 1; kk <= 3; kk += 2) {
            stage_rows<T, VEC, NTH>(row, Lp, ysb + kk * kstride, L, np, L);
            __syncthreads();
            const int32_t* ek = ent + static_cast<long long>(kk) * L;
            for (int p = threadIdx.x; p < L; p += NTH) {
                const int sp_ = skew(ek[p]);
                for (int pl = 0; pl < np; ++pl) {
                    const float v = row[pl * Lp + sp_];
                    acc1[pl * L + p] = (kk == 1) ? v : acc1[pl * L + p] + v;
                }
            }
            __syncthreads();
        }
    }
    for (int pl = 0; pl < np; ++pl) {
        const T* y0 = ysb + static_cast<long long>(pl) * L;  // k = 0
        const T* y2 = y0 + 2 * kstride;                      // k = 2
        T* yo = y + (static_cast<long long>(plane0) + pl) * L;
        for (int p = threadIdx.x; p < L; p += NTH) {
            float o;
            if (cfg.has_std) {
                const float s = (to_f32(y0[p]) + to_f32(y2[L - 1 - p])) + acc1[pl * L + p];
                o = (cfg.khi > cfg.klo) ? acc[pl * L + p] + s : s;
            } else {
                o = acc[pl * L + p];
            }
            yo[p] = from_f32<T>(o);
        }
    }
}

// Unstaged merge: any plane size; gathers straight from global / L2.
template <typename T>
__global__ void __launch_bounds__(kThreads) cross_merge_direct_kernel(const T* __restrict__ ys, T* __restrict__ y,
                                                                      const int32_t* __restrict__ plan, MergeCfg cfg,
                                                                      int K, int D, int L, long long total) {
    const int32_t* ent = plan + static_cast<long long>(K) * L;
    const int32_t* off = plan + 2ll * K * L;
    for (long long e = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * kThreads) {
        const long long plane = e / L;
        const int p = static_cast<int>(e - plane * L);
        const int b = static_cast<int>(plane / D), d = static_cast<int>(plane - static_cast<long long>(b) * D);
        const long long kstride = static_cast<long long>(D) * L;
        const T* base = ys + (static_cast<long long>(b) * K * D + d) * L;
        float acc = 0.f;
        // (four directions' loads in flight at once for the permutation orders was measured SLOWER at 192 x 192:
        // 1510 against 1320 us -- the kernel is bound by the L2 sector traffic of its strided reads, not by latency)
        for (int k = cfg.klo; k < cfg.khi; ++k) {
            const int32_t* ek = ent + static_cast<long long>(k) * L;
            int j0 = p, j1 = p + 1;
            if (!cfg.perm) j0 = off[static_cast<long long>(k) * (L + 1) + p], j1 = off[static_cast<long long>(k) * (L + 1) + p + 1];
            for (int j = j0; j < j1; ++j) acc += to_f32(base[k * kstride + ek[j]]);
        }
        float o = acc;
        if (cfg.has_std) {
            const float t13 = to_f32(base[kstride + ent[static_cast<long long>(L) + p]]) +
                              to_f32(base[3 * kstride + ent[3ll * L + p]]);
            const float s = (to_f32(base[p]) + to_f32(base[2 * kstride + (L - 1 - p)])) + t13;
            o = (cfg.khi > cfg.klo) ? acc + s : s;
        }
        y[e] = from_f32<T>(o);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T>
cudaError_t scan_t(const T* x, T* xs, const int32_t* plan, int K, int B, int C, int L, cudaStream_t st) {
    const int planes = B * C;
    const int Lp = skewed_len(L);
    const size_t plane_bytes = static_cast<size_t>(Lp) * 4;
    if (plane_bytes > 220 * 1024) return cudaErrorInvalidConfiguration;
    int ppb = static_cast<int>((48 * 1024) / plane_bytes);
    ppb = ppb < 1 ? 1 : (ppb > 16 ? 16 : ppb);
    const size_t smem = plane_bytes * ppb;
    const bool vec = (L % 4 == 0) && aligned16(x) && aligned16(xs) && aligned16(plan);
    const bool wide = smem > 72 * 1024;  // at most two CTAs per SM: 1024 threads each
    auto kern = wide ? (vec ? cross_scan_kernel<T, true, 1024> : cross_scan_kernel<T, false, 1024>)
                     : (vec ? cross_scan_kernel<T, true, kThreads> : cross_scan_kernel<T, false, kThreads>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    const int grid = (planes + ppb - 1) / ppb;
    kern<<<grid, wide ? 1024 : kThreads, smem, st>>>(x, xs, plan, K, C, L, planes, ppb);
    note_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t merge_t(const T* ys, T* y, const int32_t* plan, int order, int K, int B, int D, int L,
                    cudaStream_t st) {
    MergeCfg cfg;
    cfg.has_std = (order == TRAMBA_ORDER_STD || order == TRAMBA_ORDER_LINE);
    cfg.klo = (order == TRAMBA_ORDER_LINE) ? 4 : 0;
    cfg.khi = (order == TRAMBA_ORDER_STD) ? 0 : K;
    // every direction a permutation (one token per pixel): the CSR offsets need not be read
    cfg.perm = (order == TRAMBA_ORDER_DILATION || order == TRAMBA_ORDER_WINDOW || order == TRAMBA_ORDER_SPIRAL ||
                order == TRAMBA_ORDER_HILBERT || order == TRAMBA_ORDER_DIAGONAL || order == TRAMBA_ORDER_AB2 ||
                order == TRAMBA_ORDER_AB1);
    const int planes = B * D;
    const int Lp = skewed_len(L);
    const int nacc = (cfg.khi > cfg.klo ? 1 : 0) + (cfg.has_std ? 1 : 0);
    const size_t per_plane = static_cast<size_t>(Lp) * 4 + static_cast<size_t>(nacc) * L * 4;
    if (per_plane <= 200 * 1024) {
        int ppb = static_cast<int>((72 * 1024) / per_plane);
        ppb = ppb < 1 ? 1 : (ppb > 16 ? 16 : ppb);
        while (D % ppb != 0) --ppb;  // a block never straddles a batch index
        const size_t smem = per_plane * ppb;
        const bool vec = (L % 4 == 0) && aligned16(ys);
        const bool wide = smem > 72 * 1024;  // at most two CTAs per SM: 1024 threads each
        auto kern = wide ? (vec ? cross_merge_staged_kernel<T, true, 1024> : cross_merge_staged_kernel<T, false, 1024>)
                         : (vec ? cross_merge_staged_kernel<T, true, kThreads> : cross_merge_staged_kernel<T, false, kThreads>);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        const int grid = planes / ppb;
        kern<<<grid, wide ? 1024 : kThreads, smem, st>>>(ys, y, plan, cfg, K, D, L, planes, ppb);
    } else {
        const long long total = static_cast<long long>(planes) * L;
        long long grid = (total + kThreads - 1) / kThreads;
        if (grid > 148 * 32) grid = 148 * 32;
        cross_merge_direct_kernel<T><<<static_cast<int>(grid), kThreads, 0, st>>>(ys, y, plan, cfg, K, D, L, total);
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_cross_scan(const void* x, void* xs, const int32_t* plan, int K, int B, int C, int L,
                              int dtype, cudaStream_t st) {
    if (dtype == TRAMBA_F32)
        return scan_t(static_cast<const float*>(x), static_cast<float*>(xs), plan, K, B, C, L, st);
    return scan_t(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(xs), plan, K, B, C, L, st);
}

cudaError_t launch_cross_merge(const void* ys, void* y, const int32_t* plan, int order, int K, int B, int D,
                               int L, int dtype, cudaStream_t st) {
    if (dtype == TRAMBA_F32)
        return merge_t(static_cast<const float*>(ys), static_cast<float*>(y), plan, order, K, B, D, L, st);
    return merge_t(static_cast<const __nv_bfloat16*>(ys), static_cast<__nv_bfloat16*>(y), plan, order, K, B, D,
                   L, st);
}

}  // namespace tramba
