// Fused SS2D core (cross-scan -> selective scan -> cross-merge in ONE kernel) for sm_100a --
// shared definitions of the forward and backward kernels.
//
// Replaces the composition at Models/vmamba.py:231-259 (SS2Dv2.forward_corev2):
//     xs = CrossScan(x); ...proj...; ys = SelectiveScanOflex(xs, dts, A, Bs, Cs, Ds, bias); y = CrossMerge(ys)
// without ever materialising the permuted copies xs / ys of the feature map in HBM.
//     x    (B, D, L)      IMAGE order  -- gathered by the kernel (u_t = x[pixel(k, t)])
//     dts  (B, K, D, L)   scan order   -- the dt_proj output, as in the reference (vmamba.py:236)
//     Bs, Cs (B, K, L)    scan order   -- (d_state N = 1, the model's shape)
//     y    (B, D, L)      IMAGE order  -- every token's output is added to its pixel in shared memory
// The scan order enters only through the uint16 table idx[k][t] = skew(pixel of token t in direction
// k) (bit-exact with the reference generators before the skew): the permutation is pure
// shared-memory addressing.  (The host evaluates the tiny x_proj once on the un-permuted map and
// gathers its R+2 channels into scan order -- 1/25 of a cross-scan -- so delta/B/C arrive in scan
// order exactly as the reference produces them.)
//
// Shared-memory planes that are accessed through the table (x, dy, the merged y) are stored
// SKEWED: element p lives at skew(p) = p + 2*(p >> 6), i.e. one pad word per 32 words, so that the
// lane stride of column-major / window / dilated walks (a multiple of 32 words un-skewed) spreads
// over the banks.  They are staged with 4-byte cp.async copies (a bulk copy cannot skew);
// everything read in token order (delta, B, C) and everything written in token order (du, ddelta)
// moves with 1-D TMA bulk copies.
//
// Work decomposition.  A "round" is PP whole image planes (b, d0 .. d0+PP-1) = NT*S tokens per
// direction; thread (q, j) owns the S consecutive tokens [j*S, (j+1)*S) of plane q's sequence.
// The recurrence over a plane's L = TP*S tokens is a three-level scan: sequential fold in the
// thread, shuffle scan over the warp, running combine over the TP/32 warps of the plane.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace tramba {

constexpr int kFusedS = 18;     // tokens per thread: L = 18 * 2^m covers H = 12, 24, 48, 96
constexpr int kFusedNT = 512;   // threads per CTA (one CTA per SM, persistent over rounds)
constexpr int kFusedTok = kFusedS * kFusedNT;  // tokens per round and direction (= PP * L)

struct FusedArgs {
    const void* x;        // (B, D, L) image order
    const void* dts;      // (B, K, D, L) pre-softplus delta, scan order
    const void* Bs;       // (B, K, L) scan order
    const void* Cs;       // (B, K, L) scan order
    const float* A;       // (K*D)
    const float* Dp;      // (K*D) or null
    const float* bias;    // (K*D) or null
    const uint16_t* idx;  // (K, L) skewed pixel of token t in direction k
    float* y;             // fwd out: (B, D, L) merged, fp32, image order
    // backward
    const float* dy;      // (B, D, L) fp32, image order
    void* du;             // (B, K, D, L) in_t, scan order (the caller merges it back: dx = CrossMerge(du))
    void* ddts;           // (B, K, D, L) in_t, scan order
    float* dBs;           // (B, K, L) fp32 (accumulated with red.global.add)
    float* dCs;           // (B, K, L) fp32
    float* dA;            // (B, K*D) partial sums over L (host sums over B)
    float* dD;            // (B, K*D)
    float* dbias;         // (B, K*D)
    int B, K, D, L;
    int TP, PP;           // threads per plane (L / S), planes per round (NT / TP)
    int rounds;           // B * D / PP
    int softplus;
    uint32_t dup_mask;    // bit k set: direction k visits some pixel more than once (helix lines)
    int xb;               // 2 = double-buffered staging, 1 = single
};

// Backward, with the merge of du folded in (ss2d_fused_bwd.cu)
struct FusedBwdArgs {
    const void* x;          // (B, D, L) image order
    const void* dts;        // (B, K, D, L) scan order
    const void* Bs;         // (B, K, L)
    const void* Cs;         // (B, K, L)
    const float* A;         // (K*D)
    const float* Dp;        // (K*D) or null
    const float* bias;      // (K*D) or null
    const uint32_t* sched;  // (K, sched_len) permutation schedule (tramba_order_sched), device
    const float* dy;        // (B, D, L) fp32 image order
    float* dx;              // (B, D, L) fp32 image order, ZEROED by the caller; accumulated with bulk reductions
    void* ddts;             // (B, K, D, L) in_t scan order
    float* dBs;             // (B, K, L) fp32, zeroed by the caller
    float* dCs;
    float* dA;              // (K*D) fp32, zeroed by the caller (red.global.add over the batch)
    float* dD;
    float* dbias;
    int B, K, D;
    int sched_len;          // row stride of sched (entries)
    int sched_used[8];      // entries direction k really uses (<= sched_len; == L for a permutation)
    uint32_t kinds;         // 4 bits per direction: 0 identity, 1 reversed identity, 2 permutation, 3 with repeats
    int softplus;
    int R;                  // rounds per item (set by the launcher)
    FastDiv div_db, div_k, div_r;  // / (D / (R*PP)), / K, / R (set by the launcher)
};
cudaError_t launch_ss2d_fused_bwd(const FusedBwdArgs& a, int L, int in_dtype, cudaStream_t st);
bool ss2d_fused_bwd_fits(int L, int es);

cudaError_t launch_ss2d_fused_fwd(const FusedArgs& a, int in_dtype, cudaStream_t st);
size_t ss2d_fused_fwd_smem(int L, int es, int xb);

// skewed plane length in elements (element p lives at p + 2*(p >> 6); +2 keeps planes word-aligned)
__host__ __device__ inline int fused_skew_len(int L) { return L + 2 * ((L + 63) >> 6) + 2; }
__host__ __device__ inline int fused_skew(int p) { return p + ((p >> 6) << 1); }

// ---- cp.async (4-byte) with mbarrier completion
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

// Same staging with the geometry known at compile time (TPC = threads per plane, n = kFusedS * TPC elements):
// the kFusedS / 2 copies of a thread are fully unrolled and, for TPC a multiple of 32, their source and
// destination strides are constants (vector v = j + i*TPC lands at v + (v >> 5) = j + (j >> 5) + i*(TPC + TPC/32)).
template <int ES, int TPC>
__device__ __forceinline__ void stage_plane_skewed_c(void* dst_plane, const void* __restrict__ src_plane, int j) {
    static_assert(ES == 2 || ES == 4, "bf16 / fp32 planes");
    if constexpr (ES == 2) {
        uint32_t* d = static_cast<uint32_t*>(dst_plane);
        const uint32_t* s = static_cast<const uint32_t*>(src_plane);
        if constexpr ((TPC & 31) == 0) {
            uint32_t* dp = d + j + (j >> 5);
            const uint32_t* sp = s + j;
#pragma unroll
            for (int i = 0; i < kFusedS / 2; ++i) cp_async4(dp + i * (TPC + (TPC >> 5)), sp + i * TPC);
        } else {
#pragma unroll
            for (int i = 0; i < kFusedS / 2; ++i) {
                const int w = j + i * TPC;
                cp_async4(d + w + (w >> 5), s + w);
            }
        }
    } else {
        uint2* d = static_cast<uint2*>(dst_plane);
        const uint2* s = static_cast<const uint2*>(src_plane);
        if constexpr ((TPC & 31) == 0) {
            uint2* dp = d + j + (j >> 5);
            const uint2* sp = s + j;
#pragma unroll
            for (int i = 0; i < kFusedS / 2; ++i) cp_async8(dp + i * (TPC + (TPC >> 5)), sp + i * TPC);
        } else {
#pragma unroll
            for (int i = 0; i < kFusedS / 2; ++i) {
                const int v = j + i * TPC;
                cp_async8(d + v + (v >> 5), s + v);
            }
        }
    }
}

// ---- explicit shared-window accesses at (base pointer + byte offset): the address is formed in 32-bit shared space
// once per base (no generic-address conversion inside the permutation loops); loads first, stores later = the
// caller's instruction-level parallelism
__device__ __forceinline__ uint32_t lds_u16(const void* base, uint32_t off) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(smem_u32(base) + off) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(void* base, uint32_t off, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(base) + off), "h"(static_cast<uint16_t>(v)) : "memory");
}
__device__ __forceinline__ float lds_f32(const void* base, uint32_t off) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(base) + off) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(void* base, uint32_t off, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(base) + off), "f"(v) : "memory");
}

// Walks this thread's entries of a permutation schedule (uint2 = two moves; stride TP between a thread's entries) in
// batches of three loads, the next batch in flight while `move(dense, e0, e1, e2, count)` performs the current one.
// DENSE: every entry is a move and the entry count is a multiple of three (permutations: sched_len == L).
template <bool FROM_SMEM>
__device__ __forceinline__ uint2 sched_entry(const uint2* p) {
    if constexpr (FROM_SMEM) return *p;   // staged in shared memory for the item
    else return __ldg(p);                 // global (L2-resident table)
}
template <int TP, bool DENSE, bool FROM_SMEM = false, typename Move>
__device__ __forceinline__ void walk_schedule(const uint2* __restrict__ sc, int iters, Move&& move) {
    const uint2 none = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    constexpr std::integral_constant<bool, DENSE> dense{};
    int n = 0;
    uint2 c0 = none, c1 = none, c2 = none;
    if (iters >= 3) c0 = sched_entry<FROM_SMEM>(sc), c1 = sched_entry<FROM_SMEM>(sc + TP), c2 = sched_entry<FROM_SMEM>(sc + 2 * TP);
    for (; n + 3 <= iters; n += 3) {
        sc += 3 * TP;
        uint2 n0 = none, n1 = none, n2 = none;
        if (n + 6 <= iters) n0 = sched_entry<FROM_SMEM>(sc), n1 = sched_entry<FROM_SMEM>(sc + TP), n2 = sched_entry<FROM_SMEM>(sc + 2 * TP);
        move(dense, c0, c1, c2, 3);
        c0 = n0, c1 = n1, c2 = n2;
    }
    if (!DENSE && n < iters) {  // 1 or 2 left
        const int cnt = iters - n;
        c0 = sched_entry<FROM_SMEM>(sc);
        c1 = cnt > 1 ? sched_entry<FROM_SMEM>(sc + TP) : none;
        move(dense, c0, c1, none, cnt);
    }
}

// A permutation's schedule has exactly kFusedS / 2 entry pairs per thread: fetch them ALL first (one exposed L2 latency
// per pass instead of one per batch), then perform the moves three pairs at a time.
template <int TP, typename Move>
__device__ __forceinline__ void walk_schedule_dense(const uint2* __restrict__ sc, Move&& move) {
    constexpr std::integral_constant<bool, true> dense{};
    uint2 e[kFusedS / 2];
#pragma unroll
    for (int i = 0; i < kFusedS / 2; ++i) e[i] = __ldg(sc + i * TP);
#pragma unroll
    for (int i = 0; i < kFusedS / 2; i += 3) move(dense, e[i], e[i + 1], e[i + 2], 3);
}

// Affine-map inclusive scan over `width` lanes (width = 32 or a power of two below).
__device__ __forceinline__ void fused_scan(float& P, float& S, int sl, int width) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        if (off < width) {
            const float Pp = __shfl_up_sync(0xffffffffu, P, off, width);
            const float Sp = __shfl_up_sync(0xffffffffu, S, off, width);
            if (sl >= off) {
                S = fmaf(P, Sp, S);
                P = P * Pp;
            }
        }
    }
}
__device__ __forceinline__ void fused_scan_rev(float& P, float& S, int sl, int width) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        if (off < width) {
            const float Pp = __shfl_down_sync(0xffffffffu, P, off, width);
            const float Sp = __shfl_down_sync(0xffffffffu, S, off, width);
            if (sl + off < width) {
                S = fmaf(P, Sp, S);
                P = P * Pp;
            }
        }
    }
}

// S consecutive elements of a token-ordered shared array -> fp32 registers (S even).
template <int CNT>
__device__ __forceinline__ void load_tokens(const float* p, float (&v)[CNT]) {
    const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
    for (int i = 0; i < CNT / 2; ++i) {
        const float2 t = q[i];
        v[2 * i] = t.x, v[2 * i + 1] = t.y;
    }
}
template <int CNT>
__device__ __forceinline__ void load_tokens(const __nv_bfloat16* p, float (&v)[CNT]) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < CNT / 2; ++i) {
        const uint32_t t = q[i];
        v[2 * i] = __uint_as_float(t << 16), v[2 * i + 1] = __uint_as_float(t & 0xffff0000u);
    }
}
// elements 2*i and 2*i+1 of a token-ordered shared array
__device__ __forceinline__ void load_pair(const float* p, int i, float& lo, float& hi) {
    const float2 t = reinterpret_cast<const float2*>(p)[i];
    lo = t.x, hi = t.y;
}
__device__ __forceinline__ void load_pair(const __nv_bfloat16* p, int i, float& lo, float& hi) {
    const uint32_t t = reinterpret_cast<const uint32_t*>(p)[i];
    lo = __uint_as_float(t << 16), hi = __uint_as_float(t & 0xffff0000u);
}
// The same chunk read / written BACKWARDS (token m <-> element CNT-1-m): a direction that walks the image in reverse.
template <bool REV, typename T>
__device__ __forceinline__ void chunk_pair(const T* p, int mp, float& lo, float& hi) {
    if constexpr (!REV) {
        load_pair(p, mp, lo, hi);
    } else {
        load_pair(p, kFusedS / 2 - 1 - mp, hi, lo);
    }
}
template <int CNT>
__device__ __forceinline__ void store_tokens_rev(float* p, const float (&v)[CNT]) {
    float2* q = reinterpret_cast<float2*>(p);
#pragma unroll
    for (int i = 0; i < CNT / 2; ++i) q[i] = make_float2(v[CNT - 1 - 2 * i], v[CNT - 2 - 2 * i]);
}
template <int CNT>
__device__ __forceinline__ void store_tokens(float* p, const float (&v)[CNT]) {
    float2* q = reinterpret_cast<float2*>(p);
#pragma unroll
    for (int i = 0; i < CNT / 2; ++i) q[i] = make_float2(v[2 * i], v[2 * i + 1]);
}
template <int CNT>
__device__ __forceinline__ void store_tokens(__nv_bfloat16* p, const float (&v)[CNT]) {
    __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(p);
#pragma unroll
    for (int i = 0; i < CNT / 2; ++i) q[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
}

}  // namespace tramba
