// Fused SS2D core, backward (sm_100a).  See ss2d_fused.cuh for the data model.
// Replaces CrossMerge*.backward -> SelectiveScanOflex.backward of Models/SS2D/csms6s.py:45-55,914-923
// (and the _Dilation/_Window/_Line variants) for N = 1: the gradient of the merged map is
// gathered through the scan-order table in shared memory, so neither the K-fold copy of dy
// (CrossMerge.backward = a cross-scan) nor xs is ever materialised.
//
// Work item = (b, k, block of RB = R*PP channels): the CTA walks the R rounds of the item, each
// round = PP whole planes of direction k.  Nothing is stored by the forward: pass 1 re-runs the
// forward fold AND the reverse fold of  E_t = a_t*(g_t + E_{t+1}),  g_t = C_t*dy_t,  dh_t = g_t + E_{t+1}
// over the thread's S tokens (both are affine maps sharing the same product of a's), the two
// carries come from a forward and a reverse scan over the plane's threads, pass 2 walks the
// segment right-to-left for dh and left-to-right for h and emits, in scan order,
//     du_t     = D*dy_t + dh_t*delta_t*B_t                      (the caller folds it back: dx = CrossMerge(du))
//     ddelta_t = dh_t*(A*a_t*h_{t-1} + B_t*u_t) * sigmoid(raw)
// through token-ordered shared staging that leaves with one bulk store per round.
// dB_t = sum_d dh_t*delta_t*u_t and dC_t = sum_d dy_t*h_t are accumulated in registers over all
// rounds of the item (same b, k) and added to global memory once per item.  dA, dD, ddelta_bias
// leave as per-(b, row) partial sums (no atomics; the host sums over b).
#include "ss2d_fused.cuh"

namespace tramba {

namespace {

constexpr int S = kFusedS;
constexpr int NT = kFusedNT;
constexpr int TOK = kFusedTok;
constexpr int NW = NT / 32;

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct BwdLayout {
    size_t x, dy, d, b, c, o0, o1, tot, red, bar, total;
    int lsk;
    __host__ __device__ BwdLayout(int L, int es, int xb) {
        lsk = fused_skew_len(L);
        const int pp = TOK / L;
        size_t o = 0;
        x = o, o += align_up(static_cast<size_t>(xb) * pp * lsk * es, 16);
        dy = o, o += align_up(static_cast<size_t>(xb) * pp * lsk * 4, 16);
        d = o, o += static_cast<size_t>(xb) * TOK * es;
        b = o, o += align_up(static_cast<size_t>(L) * es, 16);
        c = o, o += align_up(static_cast<size_t>(L) * es, 16);
        o0 = o, o += static_cast<size_t>(TOK) * es;
        o1 = o, o += static_cast<size_t>(TOK) * es;
        tot = o, o += 2 * 2 * NW * sizeof(float2);
        red = o, o += 3 * NW * sizeof(float);
        bar = o, o += 8 * sizeof(uint64_t);
        total = o;
    }
};

template <typename in_t, bool SOFTPLUS>
__global__ void __launch_bounds__(NT, 1) ss2d_bwd_kernel(const FusedArgs p, const int R) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ES = sizeof(in_t);
    const int L = p.L, K = p.K, D = p.D, TP = p.TP, PP = p.PP, XB = p.xb;
    const BwdLayout lay(L, ES, XB);
    const int LSK = lay.lsk;
    in_t* sx = reinterpret_cast<in_t*>(smem + lay.x);      // [XB][PP][LSK] skewed
    float* sdy = reinterpret_cast<float*>(smem + lay.dy);  // [XB][PP][LSK] skewed
    in_t* sd = reinterpret_cast<in_t*>(smem + lay.d);      // [XB][TOK] token order
    in_t* sB = reinterpret_cast<in_t*>(smem + lay.b);      // [L]
    in_t* sC = reinterpret_cast<in_t*>(smem + lay.c);      // [L]
    in_t* so_dd = reinterpret_cast<in_t*>(smem + lay.o0);  // [TOK] ddelta staging, token order
    in_t* so_du = reinterpret_cast<in_t*>(smem + lay.o1);  // [TOK] du staging, token order
    float2* stot = reinterpret_cast<float2*>(smem + lay.tot);  // [2 parity][2 fwd/rev][NW]
    float* sred = reinterpret_cast<float*>(smem + lay.red);    // [3][NW]
    uint64_t* barX = reinterpret_cast<uint64_t*>(smem + lay.bar);  // [2] count NT: x + dy (cp.async)
    uint64_t* barD = barX + 2;                                     // [2] count 1: delta (TMA)
    uint64_t* barBC = barX + 4;                                    // [1] count 1: item B/C (TMA)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = tid / TP;
    const int j = tid - q * TP;
    const int width = TP < 32 ? TP : 32;
    const int sl = lane & (width - 1);
    const int nwg = TP > 32 ? (TP >> 5) : 1;  // warps per plane group
    const int wq = TP > 32 ? (j >> 5) : 0;
    const int gw0 = TP > 32 ? q * nwg : 0;

    const in_t* __restrict__ gx = static_cast<const in_t*>(p.x);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.dts);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bs);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cs);
    in_t* __restrict__ gdu = static_cast<in_t*>(p.du);
    in_t* __restrict__ gdd = static_cast<in_t*>(p.ddts);

    const int RB = R * PP;  // channels per item
    const int dblocks = D / RB;
    const int items = p.B * K * dblocks;
    const int my_items = (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int nrounds = my_items * R;
    if (nrounds <= 0) return;

    if (tid == 0) {
        mbar_init(&barX[0], NT), mbar_init(&barX[1], NT), mbar_init(&barD[0], 1), mbar_init(&barD[1], 1);
        mbar_init(&barBC[0], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    // item -> (b, k, first channel)
    auto item_coord = [&](int it, int& b, int& k, int& d0) {
        const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        const int db = item % dblocks;
        const int bk = item / dblocks;
        k = bk % K;
        b = bk / K;
        d0 = db * RB;
    };
    auto issue_round = [&](int s) {  // all threads (x, dy of the thread's own plane by cp.async); thread 0 adds delta
        int b, k, d0;
        item_coord(s / R, b, k, d0);
        d0 += (s % R) * PP;
        const int buf = s % XB;
        const long long pl = static_cast<long long>(b) * D + d0 + q;
        stage_plane_skewed<ES>(sx + (static_cast<size_t>(buf) * PP + q) * LSK, gx + pl * L, L, j, TP);
        stage_plane_skewed<4>(sdy + (static_cast<size_t>(buf) * PP + q) * LSK, p.dy + pl * L, L, j, TP);
        cp_async_arrive(&barX[buf]);
        if (tid == 0) {
            mbar_expect_tx(&barD[buf], TOK * sizeof(in_t));
            bulk_g2s(sd + static_cast<size_t>(buf) * TOK, gd + ((static_cast<long long>(b) * K + k) * D + d0) * L,
                     TOK * sizeof(in_t), &barD[buf]);
        }
    };
    auto issue_bc = [&](int it) {  // thread 0
        int b, k, d0;
        item_coord(it, b, k, d0);
        const uint32_t lb = L * sizeof(in_t);
        mbar_expect_tx(&barBC[0], 2 * lb);
        bulk_g2s(sB, gB + (static_cast<long long>(b) * K + k) * L, lb, &barBC[0]);
        bulk_g2s(sC, gC + (static_cast<long long>(b) * K + k) * L, lb, &barBC[0]);
    };
    if (tid == 0) issue_bc(0);
    issue_round(0);

    int s = 0;
    for (int it = 0; it < my_items; ++it) {
        int b, k, d00;
        item_coord(it, b, k, d00);
        uint32_t pk[S / 2];
        {
            const uint32_t* tp = reinterpret_cast<const uint32_t*>(p.idx + static_cast<size_t>(k) * L + j * S);
#pragma unroll
            for (int m = 0; m < S / 2; ++m) pk[m] = __ldg(tp + m);
        }
        float accB[S], accC[S];
#pragma unroll
        for (int m = 0; m < S; ++m) accB[m] = 0.f, accC[m] = 0.f;
        mbar_wait(&barBC[0], static_cast<uint32_t>(it & 1));
        const in_t* Bq = sB + j * S;
        const in_t* Cq = sC + j * S;

        for (int r = 0; r < R; ++r, ++s) {
            const int buf = s % XB;
            if (XB == 2 && s + 1 < nrounds) issue_round(s + 1);
            const int d0 = d00 + r * PP;
            const int d = d0 + q;
            const int row = k * D + d;
            const float Ar = __ldg(p.A + row);
            const float A2 = Ar * kLog2e;
            const float Dd = p.Dp ? __ldg(p.Dp + row) : 0.f;
            const float bias = p.bias ? __ldg(p.bias + row) : 0.f;
            mbar_wait(&barX[buf], static_cast<uint32_t>((s / XB) & 1));
            mbar_wait(&barD[buf], static_cast<uint32_t>((s / XB) & 1));
            const in_t* xq = sx + (static_cast<size_t>(buf) * PP + q) * LSK;
            const float* yq = sdy + (static_cast<size_t>(buf) * PP + q) * LSK;

            // ---- pass 1: forward fold (P, Sf) and reverse fold (P, Sr) of the segment
            float sp[S];
            float P = 1.f, Sf = 0.f, Sr = 0.f;
            {
                load_tokens<S>(sd + static_cast<size_t>(buf) * TOK + static_cast<size_t>(q) * L + j * S, sp);
#pragma unroll
                for (int mp = 0; mp < S / 2; ++mp) {
                    float Bv[2], Cv[2];
                    load_pair(Bq, mp, Bv[0], Bv[1]);
                    load_pair(Cq, mp, Cv[0], Cv[1]);
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int m = 2 * mp + h2;
                        const int px = h2 ? (pk[mp] >> 16) : (pk[mp] & 0xffffu);
                        const float u = to_f32(xq[px]);
                        const float raw = sp[m] + bias;
                        float e;
                        sp[m] = !SOFTPLUS ? raw : (ES == 2 ? softplus20_fast(raw) : softplus20(raw, e));
                        const float a = ex2f(sp[m] * A2);
                        const float bb = sp[m] * Bv[h2] * u;
                        const float g = Cv[h2] * yq[px];
                        Sf = fmaf(a, Sf, bb);
                        P *= a;
                        Sr = fmaf(P, g, Sr);
                    }
                }
            }
            // ---- carries: forward scan (h entering from the left), reverse scan (E entering from the right)
            float Pf = P, Pr = P;
            fused_scan(Pf, Sf, sl, width);
            fused_scan_rev(Pr, Sr, sl, width);
            float Pfe = __shfl_up_sync(0xffffffffu, Pf, 1, width);
            float Sfe = __shfl_up_sync(0xffffffffu, Sf, 1, width);
            float Pre = __shfl_down_sync(0xffffffffu, Pr, 1, width);
            float Sre = __shfl_down_sync(0xffffffffu, Sr, 1, width);
            if (sl == 0) Pfe = 1.f, Sfe = 0.f;
            if (sl == width - 1) Pre = 1.f, Sre = 0.f;
            const int par = s & 1;
            float2* totF = stot + (par * 2 + 0) * NW;
            float2* totR = stot + (par * 2 + 1) * NW;
            if (TP > 32) {
                if (lane == 31) totF[warp] = make_float2(Pf, Sf);
                if (lane == 0) totR[warp] = make_float2(Pr, Sr);
            }
            if (tid == 0 && s > 0) bulk_wait_read<0>();  // previous round's staging planes have left
            __syncthreads();
            float hin = Sfe, Ein = Sre;
            if (TP > 32) {
                float hp = 0.f;
                for (int w = 0; w < wq; ++w) {
                    const float2 t = totF[gw0 + w];
                    hp = fmaf(t.x, hp, t.y);
                }
                hin = fmaf(Pfe, hp, Sfe);
                float ep = 0.f;
                for (int w = nwg - 1; w > wq; --w) {
                    const float2 t = totR[gw0 + w];
                    ep = fmaf(t.x, ep, t.y);
                }
                Ein = fmaf(Pre, ep, Sre);
            }
            // ---- pass 2a: right-to-left, dh_t = g_t + E_{t+1}
            float dh[S];
            {
                float E = Ein;
#pragma unroll
                for (int mp = S / 2 - 1; mp >= 0; --mp) {
                    float Cv[2];
                    load_pair(Cq, mp, Cv[0], Cv[1]);
#pragma unroll
                    for (int h2 = 1; h2 >= 0; --h2) {
                        const int m = 2 * mp + h2;
                        const int px = h2 ? (pk[mp] >> 16) : (pk[mp] & 0xffffu);
                        const float g = Cv[h2] * yq[px];
                        const float a = ex2f(sp[m] * A2);
                        dh[m] = g + E;
                        E = a * dh[m];
                    }
                }
            }
            // ---- pass 2b: left-to-right, h_t; emit du, ddelta; accumulate dB, dC, dA, dD, dbias
            float h = hin, sA = 0.f, sD = 0.f, sb = 0.f;
#pragma unroll
            for (int m = 0; m < S; ++m) {
                const int px = (m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu);
                const float u = to_f32(xq[px]);
                const float dyv = yq[px];
                const float Bm = to_f32(Bq[m]);
                const float a = ex2f(sp[m] * A2);
                const float w = a * h;  // a_t * h_{t-1}
                const float bu = Bm * u;
                h = fmaf(sp[m], bu, w);  // h_t
                const float dhs = dh[m] * sp[m];
                const float du = fmaf(Dd, dyv, dhs * Bm);
                float dd = dh[m] * fmaf(Ar, w, bu);
                if (SOFTPLUS) {
                    // sigmoid(raw) = 1 - exp(-softplus(raw)) = -expm1(-sp); series below 1/8 keeps
                    // fp32 relative accuracy where 1 - exp(.) would cancel
                    const float t = sp[m];
                    if (ES == 2) {  // 16-bit operands: the cancellation error of 1 - exp(-t) (<= 6e-8 absolute) is moot
                        dd *= 1.0f - ex2f(-t * kLog2e);
                    } else {
                        const float ser = t * (1.0f - t * (0.5f - t * (0.16666667f - t * 0.041666668f)));
                        dd *= (t < 0.125f) ? ser : (1.0f - ex2f(-t * kLog2e));
                    }
                }
                accB[m] = fmaf(dhs, u, accB[m]);
                accC[m] = fmaf(dyv, h, accC[m]);
                sA = fmaf(dhs, w, sA);
                sD = fmaf(dyv, u, sD);
                sb += dd;
                sp[m] = dd;  // reuse: sp <- ddelta, dh <- du (token-ordered outputs)
                dh[m] = du;
            }
            store_tokens<S>(so_dd + static_cast<size_t>(q) * L + j * S, sp);
            store_tokens<S>(so_du + static_cast<size_t>(q) * L + j * S, dh);
            // ---- per-row sums over L: warp reduce, then over the plane's warps
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                if (off < width) {
                    sA += __shfl_xor_sync(0xffffffffu, sA, off, width);
                    sD += __shfl_xor_sync(0xffffffffu, sD, off, width);
                    sb += __shfl_xor_sync(0xffffffffu, sb, off, width);
                }
            }
            if (TP > 32 && lane == 0) sred[warp] = sA, sred[NW + warp] = sD, sred[2 * NW + warp] = sb;
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                const long long po = ((static_cast<long long>(b) * K + k) * D + d0) * L;
                bulk_s2g(gdd + po, so_dd, TOK * sizeof(in_t));
                bulk_s2g(gdu + po, so_du, TOK * sizeof(in_t));
                bulk_commit();
            }
            if (XB == 1 && s + 1 < nrounds) issue_round(s + 1);
            if (TP > 32 && j == 0) {
                float tA = 0.f, tD = 0.f, tb = 0.f;
                for (int w = 0; w < nwg; ++w) tA += sred[gw0 + w], tD += sred[NW + gw0 + w], tb += sred[2 * NW + gw0 + w];
                sA = tA, sD = tD, sb = tb;
            }
            if (j == 0) {
                const size_t o = static_cast<size_t>(b) * K * D + row;
                p.dA[o] = sA;
                if (p.dD) p.dD[o] = sD;
                if (p.dbias) p.dbias[o] = sb;
            }
        }
        // ---- item flush: dB / dC of the item's channels (token order, registers) -> global reds
        {
            float* gdB = p.dBs + (static_cast<size_t>(b) * K + k) * L + j * S;
            float* gdC = p.dCs + (static_cast<size_t>(b) * K + k) * L + j * S;
            if (PP == 1) {
#pragma unroll
                for (int m = 0; m < S; m += 2) {
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gdB + m), "f"(accB[m]), "f"(accB[m + 1])
                                 : "memory");
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gdC + m), "f"(accC[m]), "f"(accC[m + 1])
                                 : "memory");
                }
            } else {
                // PP planes hold partial sums for the same tokens: add them up in shared memory first
                if (tid == 0) bulk_wait_read<0>();  // the staging planes are reused as fp32 scratch
                __syncthreads();
                float* f = reinterpret_cast<float*>(smem + lay.o0);  // [2][L] floats <= 2 * TOK * ES bytes (PP >= 4)
                for (int e = tid; e < 2 * L; e += NT) f[e] = 0.f;
                __syncthreads();
#pragma unroll
                for (int m = 0; m < S; ++m) {
                    atomicAdd(f + j * S + m, accB[m]);
                    atomicAdd(f + L + j * S + m, accC[m]);
                }
                __syncthreads();
                float* gB2 = p.dBs + (static_cast<size_t>(b) * K + k) * L;
                float* gC2 = p.dCs + (static_cast<size_t>(b) * K + k) * L;
                for (int e = tid * 2; e < L; e += NT * 2) {
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gB2 + e), "f"(f[e]), "f"(f[e + 1]) : "memory");
                    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(gC2 + e), "f"(f[L + e]), "f"(f[L + e + 1])
                                 : "memory");
                }
            }
            __syncthreads();  // everybody is done with sB / sC (and the scratch) before the next item's loads
            if (tid == 0 && it + 1 < my_items) issue_bc(it + 1);
        }
    }
    if (tid == 0) bulk_wait_all<0>();
}

}  // namespace

size_t ss2d_fused_bwd_smem(int L, int es, int xb) { return BwdLayout(L, es, xb).total; }

// Channels per item: enough rounds that the per-item dB/dC flush is amortised (>= 8 channels).
static int pick_rounds(const FusedArgs& a) {
    int R = 1;
    while (R * a.PP < 8 && a.D % (2 * R * a.PP) == 0) R *= 2;
    return R;
}

template <typename in_t, bool SOFTPLUS>
static cudaError_t launch_t(const FusedArgs& a, cudaStream_t st) {
    const size_t cap = 227 * 1024;
    auto kern = ss2d_bwd_kernel<in_t, SOFTPLUS>;
    if (cudaError_t e = ensure_dyn_smem(kern, 227 * 1024); e != cudaSuccess) return e;
    const int sms = device_sms();
    FusedArgs b = a;
    b.xb = BwdLayout(a.L, sizeof(in_t), 2).total <= cap ? 2 : 1;
    const size_t smem = BwdLayout(a.L, sizeof(in_t), b.xb).total;
    if (smem > cap) return cudaErrorInvalidConfiguration;
    const int R = pick_rounds(a);
    const int items = a.B * a.K * (a.D / (R * a.PP));
    const int grid = items < sms ? items : sms;
    kern<<<grid, NT, smem, st>>>(b, R);
    note_launch();
    return cudaGetLastError();
}

#ifdef TRAMBA_NEXT
cudaError_t launch_ss2d_fused_bwd_cg(const FusedArgs& a, int in_dtype, cudaStream_t st);  // csrc/next/
#endif

cudaError_t launch_ss2d_fused_bwd(const FusedArgs& a, int in_dtype, cudaStream_t st) {
#ifdef TRAMBA_NEXT  // candidate build (libtramba_b200_next.so): compile-time-geometry kernels where they apply
    const cudaError_t e = launch_ss2d_fused_bwd_cg(a, in_dtype, st);
    if (e != cudaErrorNotSupported) return e;
#endif
    if (in_dtype == TRAMBA_F32) return a.softplus ? launch_t<float, true>(a, st) : launch_t<float, false>(a, st);
    if (in_dtype == TRAMBA_BF16)
        return a.softplus ? launch_t<__nv_bfloat16, true>(a, st) : launch_t<__nv_bfloat16, false>(a, st);
    return cudaErrorInvalidValue;
}

bool ss2d_fused_bwd_fits(int L, int es, uint32_t dup_mask, int K) {
    (void)dup_mask, (void)K;
    return BwdLayout(L, es, 1).total <= 227 * 1024;
}

}  // namespace tramba
