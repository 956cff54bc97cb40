// Shared device helpers for the sm_100a kernels: mbarrier / TMA bulk-copy PTX wrappers,
// fast transcendental helpers and dtype conversion.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tramba_b200.h"

namespace tramba {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------ launch accounting
void note_launch();  // capi.cu

// ------------------------------------------------------------------ per-device launch configuration
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a process that drives several GPUs
// (threads per device, nn.DataParallel) must configure every kernel on each of them.  One bit per device ordinal
// and kernel instantiation; `bytes` may grow (kernels whose footprint depends on N).
constexpr int kMaxDevices = 64;
// (kernel entry point, device ordinal) -> largest size configured so far.  Keyed by the function's ADDRESS:
// instantiations of one kernel template share a function-pointer TYPE, so a per-type static would be shared.
int* dyn_smem_slot(const void* kern, int dev);  // capi.cu (mutex-protected map; the slot itself is benign-racy)
template <typename Kern>
inline cudaError_t ensure_dyn_smem(Kern kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int* done = dyn_smem_slot(reinterpret_cast<const void*>(kern), dev & (kMaxDevices - 1));
    if (*done >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) *done = bytes;
    return e;
}
inline int device_sms() {
    static int sms[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev & (kMaxDevices - 1);
    if (!sms[slot]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        sms[slot] = v > 0 ? v : 148;
    }
    return sms[slot];
}

// ------------------------------------------------------------------ division by a run-time constant
// n / d for n < 2^31 as one multiply-high + shift (Granlund-Montgomery round-up), the constant prepared on the host
struct FastDiv {
    uint32_t d, mul, shift;
    __host__ void init(uint32_t d_) {
        d = d_;
        if (d <= 1) {
            mul = 0;
            shift = 0;
            return;
        }
        shift = 0;
        while ((1ull << shift) < d) ++shift;
        mul = static_cast<uint32_t>((((1ull << shift) - d) << 32) / d + 1);
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
        return static_cast<uint32_t>((static_cast<uint64_t>(__umulhi(n, mul)) + n) >> shift);
#else
        return static_cast<uint32_t>(((static_cast<uint64_t>(n) * mul >> 32) + n) >> shift);
#endif
    }
};


// ------------------------------------------------------------------ mbarrier + TMA bulk
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (TMA, 1-D).  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk copy (TMA, 1-D), tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
// shared -> global bulk REDUCTION (TMA, 1-D): dst[i] += src[i] in fp32, performed at L2; same group tracking.
__device__ __forceinline__ void bulk_reduce_add_f32_s2g(float* dst_gmem, const float* src_smem, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// 16-byte vector reduction (fp32 adds performed at L2, no return value); p 16-byte aligned
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ------------------------------------------------------------------ math
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softplus with torch's threshold (x > 20 -> x), branch-free.  e = exp(x) and w = 1 + e are
// returned for the backward's sigmoid.  log1p(e) = log(w) + (e - (w - 1)) / w to first order;
// the correction term restores the bits of e that rounding 1 + e drops, so small deltas keep
// fp32 relative accuracy (1/w ~ 1 wherever the term matters; it is switched off for e >= 1).
__device__ __forceinline__ float softplus20(float x, float& e_out, float& w_out) {
    const float e = ex2f(x * kLog2e);
    const float w = 1.0f + e;
    const float corr = (e < 1.0f) ? (e - (w - 1.0f)) : 0.0f;
    const float sp = fmaf(lg2f(w), kLn2, corr);
    e_out = e;
    w_out = w;
    return (x > 20.0f) ? x : sp;
}
__device__ __forceinline__ float softplus20(float x, float& e_out) {
    float w;
    return softplus20(x, e_out, w);
}
// Same without the log1p correction: absolute error <= 6e-8 (half an ulp of 1 + e), which only shows
// where delta itself is below ~1e-4.  Used when the operands are 16-bit (their own rounding is 4e-3
// relative); 4 issue slots per token cheaper.
__device__ __forceinline__ float softplus20_fast(float x) {
    const float sp = lg2f(1.0f + ex2f(x * kLog2e)) * kLn2;
    return (x > 20.0f) ? x : sp;
}

// ------------------------------------------------------------------ dtype helpers
template <typename T>
struct DTypeOf;
template <>
struct DTypeOf<float> {
    static constexpr int value = TRAMBA_F32;
};
template <>
struct DTypeOf<__nv_bfloat16> {
    static constexpr int value = TRAMBA_BF16;
};
template <>
struct DTypeOf<__half> {
    static constexpr int value = TRAMBA_F16;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) {
    return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) {
    return __float2half_rn(v);
}

// Load / store CNT consecutive elements (CNT % 4 == 0) as fp32 with the widest aligned
// vector access the element type allows (16 B for fp32, 8 B for 16-bit types).
template <int CNT, typename T>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&v)[CNT]) {
    static_assert(CNT % 4 == 0, "CNT % 4");
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < CNT / 4; ++i) {
            float4 q = reinterpret_cast<const float4*>(p)[i];
            v[4 * i] = q.x, v[4 * i + 1] = q.y, v[4 * i + 2] = q.z, v[4 * i + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < CNT / 4; ++i) {
            uint2 q = reinterpret_cast<const uint2*>(p)[i];
            const T* h = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[4 * i + j] = to_f32(h[j]);
        }
    }
}
template <int CNT, typename T>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&v)[CNT]) {
    static_assert(CNT % 4 == 0, "CNT % 4");
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < CNT / 4; ++i)
            reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < CNT / 4; ++i) {
            uint2 q;
            T* h = reinterpret_cast<T*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = from_f32<T>(v[4 * i + j]);
            reinterpret_cast<uint2*>(p)[i] = q;
        }
    }
}

}  // namespace tramba
