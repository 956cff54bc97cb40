// Fused SS2D core, forward (sm_100a).  See ss2d_fused.cuh for the data model.
// Replaces CrossScan* -> SelectiveScanOflex -> CrossMerge* of Models/vmamba.py:231-259 for N = 1.
//
// Persistent CTA (512 threads, one per SM) walks "rounds" of PP whole planes (PP * L = 9216 tokens) and loops the K
// directions inside; thread = 18 consecutive tokens of one plane: local fold -> shuffle scan -> combine over the
// plane's warps; every token's output is added to its pixel of the merged plane held in shared memory.  Geometry is a
// template parameter (TPC = threads per plane = L / 18, XB = staging depth): the index arithmetic folds, the skewed
// cp.async staging of x is fully unrolled, the un-skewing copy-out divides by a constant, and the per-direction
// operands (A, D, bias, the thread's 9 table words) are loaded one stage ahead.  (Round 2 measured this
// compile-time-geometry kernel 12 % faster than the run-time-geometry one it replaces, bit-identical outputs, and on
// par or ahead of a TMA-staged direction-pairing variant: profiles/r02_fused_ab_bf16.jsonl)
#include "ss2d_fused.cuh"

#ifndef TRAMBA_PACK
#define TRAMBA_PACK 1
#endif

namespace tramba {

namespace {

constexpr int S = kFusedS;
constexpr int NT = kFusedNT;
constexpr int TOK = kFusedTok;
constexpr int NW = NT / 32;

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// shared-memory layout (byte offsets), identical on host and device
struct FwdLayout {
    size_t x, d, b, c, y, tot, bar, total;
    int lsk;  // skewed plane length, elements
    __host__ __device__ FwdLayout(int L, int es, int xb) {
        lsk = fused_skew_len(L);
        const int pp = TOK / L;
        size_t o = 0;
        x = o, o += align_up(static_cast<size_t>(xb) * pp * lsk * es, 16);
        d = o, o += static_cast<size_t>(xb) * TOK * es;
        b = o, o += align_up(static_cast<size_t>(xb) * L * es, 16);
        c = o, o += align_up(static_cast<size_t>(xb) * L * es, 16);
        y = o, o += align_up(static_cast<size_t>(pp) * lsk * 4, 16);
        tot = o, o += 2 * NW * sizeof(float2);
        bar = o, o += 8 * sizeof(uint64_t);
        total = o;
    }
};

template <typename in_t, bool SOFTPLUS, int TPC, int XB>
__global__ void __launch_bounds__(NT, 1) ss2d_fwd_kernel(const FusedArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ES = sizeof(in_t);
    constexpr int TP = TPC, L = S * TPC, PP = NT / TPC;
    static_assert(TPC >= 1 && (TPC & (TPC - 1)) == 0 && TPC <= NT, "threads per plane: a power of two <= 512");
    const int K = p.K, D = p.D;
    const FwdLayout lay(L, ES, XB);
    const int LSK = lay.lsk;
    in_t* sx = reinterpret_cast<in_t*>(smem + lay.x);    // [XB][PP][LSK] skewed
    in_t* sd = reinterpret_cast<in_t*>(smem + lay.d);    // [XB][TOK] token order
    in_t* sB = reinterpret_cast<in_t*>(smem + lay.b);    // [XB][L]
    in_t* sC = reinterpret_cast<in_t*>(smem + lay.c);    // [XB][L]
    float* sy = reinterpret_cast<float*>(smem + lay.y);  // [PP][LSK] skewed
    float2* stot = reinterpret_cast<float2*>(smem + lay.tot);
    uint64_t* barX = reinterpret_cast<uint64_t*>(smem + lay.bar);  // [2] count NT (cp.async arrivals)
    uint64_t* barD = barX + 2;                                     // [2] count 1 (TMA tx)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = tid / TP;       // plane slot of this thread
    const int j = tid - q * TP;   // segment index inside the plane
    const int width = TP < 32 ? TP : 32;
    const int sl = lane & (width - 1);
    const int wq = TP > 32 ? (j >> 5) : 0;          // warp index inside the plane group
    const int gw0 = TP > 32 ? (q * (TP >> 5)) : 0;  // first warp of the plane group

    const in_t* __restrict__ gx = static_cast<const in_t*>(p.x);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.dts);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bs);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cs);

    const int my_rounds = (p.rounds - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                          static_cast<int>(gridDim.x);
    const int nstages = my_rounds * K;
    if (nstages <= 0) return;

    if (tid == 0) {
        mbar_init(&barX[0], NT), mbar_init(&barX[1], NT), mbar_init(&barD[0], 1), mbar_init(&barD[1], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    auto issue_x = [&](int i) {  // all threads: thread (q, j) stages its own plane
        const long long plane0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP;
        const int buf = i % XB;
        stage_plane_skewed_c<ES, TPC>(sx + (static_cast<size_t>(buf) * PP + q) * LSK, gx + (plane0 + q) * L, j);
        cp_async_arrive(&barX[buf]);
    };
    auto issue_dir = [&](int s) {  // thread 0
        const int i = s / K, k = s - i * K;
        const long long plane0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP;
        const long long b = plane0 / D, d0 = plane0 - b * D;
        const int buf = s % XB;
        const uint32_t lb = L * sizeof(in_t);
        mbar_expect_tx(&barD[buf], TOK * sizeof(in_t) + 2 * lb);
        bulk_g2s(sd + static_cast<size_t>(buf) * TOK, gd + ((b * K + k) * D + d0) * L, TOK * sizeof(in_t), &barD[buf]);
        bulk_g2s(sB + static_cast<size_t>(buf) * L, gB + (b * K + k) * L, lb, &barD[buf]);
        bulk_g2s(sC + static_cast<size_t>(buf) * L, gC + (b * K + k) * L, lb, &barD[buf]);
    };
    issue_x(0);
    if (tid == 0) issue_dir(0);

    // operands of one (round, direction) stage that do not come through shared memory: the row's A / D / bias
    // and the thread's S table entries (packed pairs).  Loaded one stage ahead of their use.
    auto load_stage = [&](int d, int k, float& Ar, float& Dd, float& bias, uint32_t (&pk)[S / 2]) {
        const int row = k * D + d;
        Ar = __ldg(p.A + row);
        Dd = p.Dp ? __ldg(p.Dp + row) : 0.f;
        bias = p.bias ? __ldg(p.bias + row) : 0.f;
        const uint32_t* tp = reinterpret_cast<const uint32_t*>(p.idx + static_cast<size_t>(k) * L + j * S);
#pragma unroll
        for (int m = 0; m < S / 2; ++m) pk[m] = __ldg(tp + m);
    };
    auto round_d = [&](int i) {  // channel of this thread's plane in round i
        const long long plane0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP;
        return static_cast<int>(plane0 % D) + q;
    };
    float nAr, nDd, nbias;
    uint32_t npk[S / 2];
    load_stage(round_d(0), 0, nAr, nDd, nbias, npk);

    int s = 0;
    for (int i = 0; i < my_rounds; ++i) {
        const long long plane0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP;
        const int b = static_cast<int>(plane0 / D);
        const int d = static_cast<int>(plane0 - static_cast<long long>(b) * D) + q;
        const in_t* xq = sx + (static_cast<size_t>(i % XB) * PP + q) * LSK;
        float* yq = sy + static_cast<size_t>(q) * LSK;
        if (XB == 2 && i + 1 < my_rounds) issue_x(i + 1);
        mbar_wait(&barX[i % XB], static_cast<uint32_t>((i / XB) & 1));

        for (int k = 0; k < K; ++k, ++s) {
            const int buf = s % XB;
            if (XB == 2 && tid == 0 && s + 1 < nstages) issue_dir(s + 1);
            const float A2 = nAr * kLog2e, Dd = nDd, bias = nbias;  // loaded during the previous stage
            uint32_t pk[S / 2];
#pragma unroll
            for (int m = 0; m < S / 2; ++m) pk[m] = npk[m];
            if (s + 1 < nstages) {  // stage s + 1 = next direction of this round, or direction 0 of the next round
                if (k + 1 < K) load_stage(d, k + 1, nAr, nDd, nbias, npk);
                else load_stage(round_d(i + 1), 0, nAr, nDd, nbias, npk);
            }
            mbar_wait(&barD[buf], static_cast<uint32_t>((s / XB) & 1));

            // ---- pass 1: local fold from a zero carry-in
            float y0[S], g[S];
            float P = 1.f, Sacc = 0.f;
            {
                float dl[S], Bv[S], Cv[S];
                load_tokens<S>(sd + static_cast<size_t>(buf) * TOK + static_cast<size_t>(q) * L + j * S, dl);
                load_tokens<S>(sB + static_cast<size_t>(buf) * L + j * S, Bv);
                load_tokens<S>(sC + static_cast<size_t>(buf) * L + j * S, Cv);
                if constexpr (TRAMBA_PACK && SOFTPLUS && ES == 2) {
                    // 16-bit operands: the per-token operations that do not depend on the running state are issued for
                    // token PAIRS as packed fp32 instructions (FADD2 / FMUL2 / FFMA2: two tokens per issue slot, same
                    // IEEE rounding as the scalar forms) -- the kernel is bound by issue slots, not by FMA throughput
                    const float2 bias2 = make_float2(bias, bias), A22 = make_float2(A2, A2), Dd2 = make_float2(Dd, Dd);
                    const float2 l2e = make_float2(kLog2e, kLog2e), ln2 = make_float2(kLn2, kLn2), one2 = make_float2(1.f, 1.f);
#pragma unroll
                    for (int mp = 0; mp < S / 2; ++mp) {
                        const float2 u2 = make_float2(to_f32(xq[pk[mp] & 0xffffu]), to_f32(xq[pk[mp] >> 16]));
                        const float2 raw2 = __fadd2_rn(make_float2(dl[2 * mp], dl[2 * mp + 1]), bias2);
                        const float2 t2 = __fmul2_rn(raw2, l2e);
                        const float2 w2 = __fadd2_rn(make_float2(ex2f(t2.x), ex2f(t2.y)), one2);
                        float2 sp2 = __fmul2_rn(make_float2(lg2f(w2.x), lg2f(w2.y)), ln2);
                        sp2.x = raw2.x > 20.f ? raw2.x : sp2.x;
                        sp2.y = raw2.y > 20.f ? raw2.y : sp2.y;
                        const float2 ea = __fmul2_rn(sp2, A22);
                        const float a0 = ex2f(ea.x), a1 = ex2f(ea.y);
                        const float2 Bm = make_float2(Bv[2 * mp], Bv[2 * mp + 1]), Cm = make_float2(Cv[2 * mp], Cv[2 * mp + 1]);
                        const float2 bb2 = __fmul2_rn(__fmul2_rn(sp2, Bm), u2);
                        const float2 du2 = __fmul2_rn(Dd2, u2);
                        const float S0 = fmaf(a0, Sacc, bb2.x);
                        const float P0 = P * a0;
                        Sacc = fmaf(a1, S0, bb2.y);
                        P = P0 * a1;
                        const float2 y2 = __ffma2_rn(Cm, make_float2(S0, Sacc), du2);
                        const float2 g2 = __fmul2_rn(Cm, make_float2(P0, P));
                        y0[2 * mp] = y2.x, y0[2 * mp + 1] = y2.y;
                        g[2 * mp] = g2.x, g[2 * mp + 1] = g2.y;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < S; ++m) {
                        const int px = (m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu);
                        const float u = to_f32(xq[px]);
                        const float raw = dl[m] + bias;
                        float e;
                        const float dsp = !SOFTPLUS ? raw : (ES == 2 ? softplus20_fast(raw) : softplus20(raw, e));
                        const float a = ex2f(dsp * A2);
                        const float bb = dsp * Bv[m] * u;
                        Sacc = fmaf(a, Sacc, bb);
                        P *= a;
                        y0[m] = fmaf(Cv[m], Sacc, Dd * u);
                        g[m] = Cv[m] * P;
                    }
                }
            }
            // ---- carry-in of this thread: scan over the TP threads of the plane
            fused_scan(P, Sacc, sl, width);
            float Pe = __shfl_up_sync(0xffffffffu, P, 1, width);
            float Se = __shfl_up_sync(0xffffffffu, Sacc, 1, width);
            if (sl == 0) Pe = 1.f, Se = 0.f;
            const int par = s & 1;
            if (TP > 32 && lane == 31) stot[par * NW + warp] = make_float2(P, Sacc);
            __syncthreads();
            if (XB == 1) {  // single-buffered staging: everybody is past pass 1, the buffers are free
                if (tid == 0 && s + 1 < nstages) issue_dir(s + 1);
                if (k == K - 1 && i + 1 < my_rounds) issue_x(i + 1);
            }
            float hin = Se;
            if (TP > 32) {
                float hp = 0.f;
                for (int w = 0; w < wq; ++w) {
                    const float2 t = stot[par * NW + gw0 + w];
                    hp = fmaf(t.x, hp, t.y);
                }
                hin = fmaf(Pe, hp, Se);
            }
            // ---- pass 2: add the carry-in term and merge into the (skewed) image plane
            const bool dup = (p.dup_mask >> k) & 1u;
            if (k == 0) {
#pragma unroll
                for (int m = 0; m < S; ++m) {
                    const int px = (m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu);
                    yq[px] = fmaf(g[m], hin, y0[m]);
                }
            } else if (!dup) {
#pragma unroll
                for (int m = 0; m < S; ++m) {
                    const int px = (m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu);
                    yq[px] += fmaf(g[m], hin, y0[m]);
                }
            } else {
#pragma unroll
                for (int m = 0; m < S; ++m) {
                    const int px = (m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu);
                    atomicAdd(yq + px, fmaf(g[m], hin, y0[m]));
                }
            }
        }
        // ---- merged planes of this round -> global (PP planes are contiguous in y), un-skewing
        __syncthreads();
        {
            float* gy = p.y + plane0 * L;
            for (int e = tid * 4; e < TOK; e += NT * 4) {
                const int qq = e / L, pp = e - qq * L;  // 4 consecutive pixels never straddle a pad (4 | 64)
                const float* src = sy + static_cast<size_t>(qq) * LSK + fused_skew(pp);
                *reinterpret_cast<float4*>(gy + e) = make_float4(src[0], src[1], src[2], src[3]);
            }
        }
        // the next round's first scatter into sy happens after its mid-stage barrier, which every
        // thread reaches only after finishing the copy-out above
    }
}

template <typename in_t, bool SOFTPLUS, int TPC, int XB>
static cudaError_t launch_fwd(const FusedArgs& a, cudaStream_t st) {
    auto kern = ss2d_fwd_kernel<in_t, SOFTPLUS, TPC, XB>;
    if (cudaError_t e = ensure_dyn_smem(kern, 227 * 1024); e != cudaSuccess) return e;
    const int sms = device_sms();
    FusedArgs b = a;
    b.xb = XB;
    const size_t smem = FwdLayout(a.L, sizeof(in_t), XB).total;
    const int grid = a.rounds < sms ? a.rounds : sms;
    kern<<<grid, NT, smem, st>>>(b);
    note_launch();
    return cudaGetLastError();
}

template <typename in_t, bool SOFTPLUS, int XB>
static cudaError_t dispatch_tp(const FusedArgs& a, cudaStream_t st) {
    switch (a.TP) {  // L = 18 * TP: 144, 576, 2304, 9216 -- every Tramba-V map at 384^2
        case 8: return launch_fwd<in_t, SOFTPLUS, 8, XB>(a, st);
        case 32: return launch_fwd<in_t, SOFTPLUS, 32, XB>(a, st);
        case 128: return launch_fwd<in_t, SOFTPLUS, 128, XB>(a, st);
        case 512: return launch_fwd<in_t, SOFTPLUS, 512, XB>(a, st);
        default: return cudaErrorInvalidConfiguration;
    }
}
template <typename in_t, bool SOFTPLUS>
static cudaError_t dispatch_xb(const FusedArgs& a, cudaStream_t st) {
    return a.xb == 2 ? dispatch_tp<in_t, SOFTPLUS, 2>(a, st) : dispatch_tp<in_t, SOFTPLUS, 1>(a, st);
}

}  // namespace

size_t ss2d_fused_fwd_smem(int L, int es, int xb) { return FwdLayout(L, es, xb).total; }

// The staging depth is the caller's choice (a.xb, as tramba_ss2d_fused_fwd picks it: 2 if it fits 227 KB, else 1).
cudaError_t launch_ss2d_fused_fwd(const FusedArgs& a, int in_dtype, cudaStream_t st) {
    if (a.L != S * a.TP || a.PP * a.TP != NT) return cudaErrorInvalidConfiguration;
    if (in_dtype == TRAMBA_BF16)
        return a.softplus ? dispatch_xb<__nv_bfloat16, true>(a, st) : dispatch_xb<__nv_bfloat16, false>(a, st);
    if (in_dtype == TRAMBA_F32) return a.softplus ? dispatch_xb<float, true>(a, st) : dispatch_xb<float, false>(a, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
