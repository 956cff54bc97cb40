// Per-direction gather of the projected per-token channels into scan order, and its adjoint.
//
// The fused SS2D core evaluates x_proj ONCE on the image-ordered map (it is a per-token linear map, so
// it commutes with the cross-scan gather; Models/vmamba.py:233-236).  What the scan consumes is that
// result in SCAN order, split into the dt-rank channels, B and C:
//     src (B, K, R + 2N, L) image order  ->  dts_r (B, K, R, L), Bs (B, K, N, L), Cs (B, K, N, L)
//     out[b, k, c, t] = src[b, k, c, idx[k][t]]
// i.e. the reference's xs -> x_dbl -> torch.split (vmamba.py:231-235) restricted to the R + 2N projected
// channels (1/25 of a cross-scan of the feature map).  One kernel replaces gather + split + three
// .contiguous() copies; the adjoint (one kernel) replaces cat + zero-fill + scatter_add_: it is a
// GATHER through the inverse (CSR) table of the order plan, so duplicate hits of the helix lines are
// added deterministically in token order, without atomics.
#include <type_traits>

#include "common.cuh"

namespace tramba {

namespace {


// row = (b, k, c); the row is staged in shared memory once (coalesced), the permuted side is an LDS
// (block size chosen by the launcher: these one-row-per-CTA kernels are latency-bound, long rows get 1024 threads)
template <typename T>
__global__ void __launch_bounds__(1024) proj_gather_kernel(const T* __restrict__ src, T* __restrict__ dts,
                                                               T* __restrict__ Bs, T* __restrict__ Cs,
                                                               const int32_t* __restrict__ table, int K, int R, int N,
                                                               int L) {
    extern __shared__ float srow[];
    const int CX = R + 2 * N;
    const int row = blockIdx.x;               // (b * K + k) * CX + c
    const int c = row % CX, bk = row / CX, k = bk % K;
    const T* s = src + static_cast<size_t>(row) * L;
    const int nth = blockDim.x;
    for (int e = threadIdx.x; e < L; e += nth) srow[e] = to_f32(s[e]);
    __syncthreads();
    T* out = c < R ? dts + (static_cast<size_t>(bk) * R + c) * L
                   : c < R + N ? Bs + (static_cast<size_t>(bk) * N + (c - R)) * L
                               : Cs + (static_cast<size_t>(bk) * N + (c - R - N)) * L;
    const int32_t* tk = table + static_cast<size_t>(k) * L;
    for (int t = threadIdx.x; t < L; t += nth) out[t] = from_f32<T>(srow[__ldg(tk + t)]);
}

// adjoint: gsrc[b, k, c, p] = sum_{t : idx[k][t] == p} g[b, k, c, t], tokens of a pixel in ascending order
template <typename T, typename G>
__global__ void __launch_bounds__(1024) proj_scatter_kernel(const T* __restrict__ gdts, const G* __restrict__ gBs,
                                                                const G* __restrict__ gCs, T* __restrict__ gsrc,
                                                                const int32_t* __restrict__ ent,
                                                                const int32_t* __restrict__ off, int K, int R, int N,
                                                                int L) {
    extern __shared__ float srow[];
    const int CX = R + 2 * N;
    const int row = blockIdx.x;
    const int c = row % CX, bk = row / CX, k = bk % K;
    const int nth = blockDim.x;
    if (c < R) {
        const T* g = gdts + (static_cast<size_t>(bk) * R + c) * L;
        for (int e = threadIdx.x; e < L; e += nth) srow[e] = to_f32(g[e]);
    } else {
        const G* g = c < R + N ? gBs + (static_cast<size_t>(bk) * N + (c - R)) * L
                               : gCs + (static_cast<size_t>(bk) * N + (c - R - N)) * L;
        for (int e = threadIdx.x; e < L; e += nth) srow[e] = to_f32(g[e]);
    }
    __syncthreads();
    const int32_t* ek = ent + static_cast<size_t>(k) * L;
    const int32_t* ok = off + static_cast<size_t>(k) * (L + 1);
    T* out = gsrc + static_cast<size_t>(row) * L;
    for (int p = threadIdx.x; p < L; p += nth) {
        const int e0 = __ldg(ok + p), e1 = __ldg(ok + p + 1);
        float acc = 0.f;
        for (int e = e0; e < e1; ++e) acc += srow[__ldg(ek + e)];
        out[p] = from_f32<T>(acc);
    }
}

// out[n] = sum_b in[b][n] (fp32 accumulation): the reduction over the image batch of the x_proj / dt_proj weight
// gradients (dW = sum_b dY_b X_b^T is evaluated as one batched GEMM) and over the per-tile / per-plane partials of
// the LayerNorm2d and depthwise-conv weight gradients.  Block = 32 column groups x 8 row slices: a thread adds the rows
// b = slice, slice + 8, ... of its V columns (16-byte loads), the 8 slices are folded in shared memory.  Columns that
// are not a multiple of V (or rows that are not 16-byte aligned) take the scalar path.
template <typename T>
__global__ void __launch_bounds__(256) batch_sum_kernel(const T* __restrict__ in, T* __restrict__ out, int B, long long n) {
    constexpr int V = 16 / sizeof(T);
    __shared__ float part[8][32][V + 1];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    in += static_cast<long long>(blockIdx.y) * B * n, out += static_cast<long long>(blockIdx.y) * n;  // group
    const long long i0 = (static_cast<long long>(blockIdx.x) * 32 + lane) * V;
    const bool vec = (n % V) == 0;
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    if (i0 < n) {
        if (vec) {
            for (int b = slice; b < B; b += 8) {
                const uint4 q = *reinterpret_cast<const uint4*>(in + static_cast<long long>(b) * n + i0);
                const T* h = reinterpret_cast<const T*>(&q);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += to_f32(h[v]);
            }
        } else {
            for (int b = slice; b < B; b += 8)
                for (int v = 0; v < V && i0 + v < n; ++v) acc[v] += to_f32(in[static_cast<long long>(b) * n + i0 + v]);
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) part[slice][lane][v] = acc[v];
    __syncthreads();
    if (slice == 0 && i0 < n) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float t = 0.f;
#pragma unroll
            for (int s8 = 0; s8 < 8; ++s8) t += part[s8][lane][v];
            acc[v] = t;
        }
        if (vec) {
            uint4 o;
            T* h = reinterpret_cast<T*>(&o);
#pragma unroll
            for (int v = 0; v < V; ++v) h[v] = from_f32<T>(acc[v]);
            *reinterpret_cast<uint4*>(out + i0) = o;
        } else {
            for (int v = 0; v < V && i0 + v < n; ++v) out[i0 + v] = from_f32<T>(acc[v]);
        }
    }
}

// The same reduction for TALL inputs (hundreds of partial rows, a few hundred columns: the per-tile partials of the
// LayerNorm2d / depthwise-conv weight gradients).  The vector kernel above would run on 2-4 CTAs; here a CTA owns 32
// columns and its 32 warps each add every 32nd row (128-byte coalesced row segments), then fold in shared memory.
template <typename T>
__global__ void __launch_bounds__(1024) batch_sum_tall_kernel(const T* __restrict__ in, T* __restrict__ out, int B, long long n) {
    __shared__ float part[32][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    in += static_cast<long long>(blockIdx.y) * B * n, out += static_cast<long long>(blockIdx.y) * n;  // group
    const long long i = static_cast<long long>(blockIdx.x) * 32 + lane;
    float acc0 = 0.f, acc1 = 0.f;
    if (i < n) {
        int b = slice;
        for (; b + 32 < B; b += 64)
            acc0 += to_f32(in[static_cast<long long>(b) * n + i]), acc1 += to_f32(in[static_cast<long long>(b + 32) * n + i]);
        if (b < B) acc0 += to_f32(in[static_cast<long long>(b) * n + i]);
    }
    part[slice][lane] = acc0 + acc1;
    __syncthreads();
    if (slice == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int s32 = 0; s32 < 32; ++s32) t += part[s32][lane];
        out[i] = from_f32<T>(t);
    }
}

}  // namespace

cudaError_t launch_batch_sum(const void* in, void* out, int G, int B, long long n, int dtype, cudaStream_t st) {
    const int V = dtype == TRAMBA_F32 ? 4 : 8;
    const long long groups = (n + V - 1) / V;
    const dim3 grid(static_cast<unsigned>((groups + 31) / 32), static_cast<unsigned>(G));
    if (B >= 64 && static_cast<long long>(grid.x) * grid.y < 64) {  // tall and narrow: more, deeper CTAs
        const dim3 gt(static_cast<unsigned>((n + 31) / 32), static_cast<unsigned>(G));
        if (dtype == TRAMBA_F32)
            batch_sum_tall_kernel<float><<<gt, 1024, 0, st>>>(static_cast<const float*>(in), static_cast<float*>(out), B, n);
        else if (dtype == TRAMBA_BF16)
            batch_sum_tall_kernel<__nv_bfloat16><<<gt, 1024, 0, st>>>(static_cast<const __nv_bfloat16*>(in),
                                                                      static_cast<__nv_bfloat16*>(out), B, n);
        else
            return cudaErrorInvalidValue;
        note_launch();
        return cudaGetLastError();
    }
    if (dtype == TRAMBA_F32)
        batch_sum_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(in), static_cast<float*>(out), B, n);
    else if (dtype == TRAMBA_BF16)
        batch_sum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(in),
                                                              static_cast<__nv_bfloat16*>(out), B, n);
    else
        return cudaErrorInvalidValue;
    note_launch();
    return cudaGetLastError();
}

// threads per one-row CTA: ~8 elements per thread, 256 .. 1024
static int row_threads(int L) {
    int t = (L / 8 + 31) / 32 * 32;
    return t < 256 ? 256 : (t > 1024 ? 1024 : t);
}

cudaError_t launch_proj_gather(const void* src, void* dts, void* Bs, void* Cs, const int32_t* plan, int K, int B, int R,
                               int N, int L, int dtype, cudaStream_t st) {
    const int rows = B * K * (R + 2 * N);
    const size_t smem = static_cast<size_t>(L) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    auto go = [&](auto kern, auto* tp) {
        using T = std::remove_pointer_t<decltype(tp)>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<rows, row_threads(L), smem, st>>>(static_cast<const T*>(src), static_cast<T*>(dts), static_cast<T*>(Bs),
                                           static_cast<T*>(Cs), plan, K, R, N, L);
        return cudaGetLastError();
    };
    cudaError_t e = cudaErrorInvalidValue;
    if (dtype == TRAMBA_F32) e = go(proj_gather_kernel<float>, static_cast<float*>(nullptr));
    if (dtype == TRAMBA_BF16) e = go(proj_gather_kernel<__nv_bfloat16>, static_cast<__nv_bfloat16*>(nullptr));
    note_launch();
    return e;
}

// gdts has `dtype`; gBs / gCs have `bc_dtype` (the fused core returns them in fp32)
cudaError_t launch_proj_scatter(const void* gdts, const void* gBs, const void* gCs, void* gsrc, const int32_t* plan,
                                int K, int B, int R, int N, int L, int dtype, int bc_dtype, cudaStream_t st) {
    const int rows = B * K * (R + 2 * N);
    const size_t smem = static_cast<size_t>(L) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
    const int32_t* ent = plan + static_cast<size_t>(K) * L;
    const int32_t* off = plan + 2 * static_cast<size_t>(K) * L;
    auto go = [&](auto kern, auto* tp, auto* gp) {
        using T = std::remove_pointer_t<decltype(tp)>;
        using G = std::remove_pointer_t<decltype(gp)>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<rows, row_threads(L), smem, st>>>(static_cast<const T*>(gdts), static_cast<const G*>(gBs),
                                           static_cast<const G*>(gCs), static_cast<T*>(gsrc), ent, off, K, R, N, L);
        return cudaGetLastError();
    };
    using bf = __nv_bfloat16;
    cudaError_t e = cudaErrorInvalidValue;
    if (dtype == TRAMBA_F32 && bc_dtype == TRAMBA_F32)
        e = go(proj_scatter_kernel<float, float>, static_cast<float*>(nullptr), static_cast<float*>(nullptr));
    else if (dtype == TRAMBA_BF16 && bc_dtype == TRAMBA_F32)
        e = go(proj_scatter_kernel<bf, float>, static_cast<bf*>(nullptr), static_cast<float*>(nullptr));
    else if (dtype == TRAMBA_BF16 && bc_dtype == TRAMBA_BF16)
        e = go(proj_scatter_kernel<bf, bf>, static_cast<bf*>(nullptr), static_cast<bf*>(nullptr));
    note_launch();
    return e;
}

}  // namespace tramba
