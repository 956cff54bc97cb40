// Fused SS2D core, forward, on TMA-staged planes with DIRECTION PAIRING (sm_100a).  See ss2d_fused.cuh.
// Replaces CrossScan* -> SelectiveScanOflex -> CrossMerge* of Models/vmamba.py:231-259 for N = 1.
//
// Every scan order of the reference comes in pairs (k, k') with token t of k' = token L-1-t of k (std: (0,2), (1,3);
// dilation: (0,2), (1,3), Dilation.py:23-24; window: (0,1), (2,3), Window.py:28-34): both directions of a pair visit
// the SAME pixels, back to front.  A thread that owns pixels [jS, (j+1)S) of the pair's token order is segment j of
// direction k and segment TP-1-j of direction k': it reads its u values ONCE, runs k with the carry scan going up the
// lanes and k' with it going down, sums the two outputs of every pixel in registers and merges them into the image
// plane ONCE.  Compared with the first-generation kernel (ss2d_fused_fwd.cu): half the gathers, half the merge
// read-modify-writes; planes staged by 1-D TMA bulk copies instead of per-thread cp.async; the scan order applied by a
// shared -> shared permutation pass on the host-made conflict-free schedule (tramba_order_sched) instead of
// table-driven gathers / scatters with the lanes 18 tokens apart; geometry as template parameter.
// Helix line directions (SpiralLine.py:27-82; per-line reversal, repeated pixels) have no partner: single steps, merged
// with shared-memory atomics as before.
//
// Item = PP whole planes (b, d0 .. d0+PP-1); the K directions are walked as the step list of FusedFwdPairArgs.
// Per step: [gather image -> token order] -> wait delta/B/C (TMA, double buffered over the steps) -> local fold ->
// scan over the plane's threads -> outputs -> [sum into the pair's registers] -> [merge: identity orders add their own
// chunk in place; others write token order and scatter-add along the schedule].  After the last step the merged fp32
// planes leave with one bulk store.
#include "ss2d_fused.cuh"

namespace tramba {

namespace {

constexpr int S = kFusedS;
constexpr int NT = kFusedNT;
constexpr int TOK = kFusedTok;
constexpr int NW = NT / 32;

constexpr size_t al16(size_t v) { return (v + 15) / 16 * 16; }

template <int ES, int TP>
struct PairLayout {
    static constexpr int L = TP * S;
    static constexpr bool DB = ES == 2;  // 16-bit operands: delta / B / C double buffered over the steps
    static constexpr size_t XB = static_cast<size_t>(TOK) * ES, YB = static_cast<size_t>(TOK) * 4;
    static constexpr size_t LB = al16(static_cast<size_t>(L) * ES);
    static constexpr size_t x = 0;                    // image-ordered x planes (TMA)
    static constexpr size_t xt = XB;                  // token-ordered x planes
    static constexpr size_t d0 = 2 * XB, d1 = DB ? d0 + XB : d0;
    static constexpr size_t b0 = d1 + XB, c0 = b0 + LB;
    static constexpr size_t b1 = DB ? c0 + LB : b0, c1 = DB ? b1 + LB : c0;
    static constexpr size_t y = c1 + LB;              // merged output planes, fp32, image order
    static constexpr size_t yt = ES == 4 ? xt : y + YB;  // token-ordered outputs of a group (fp32 operands: over xt)
    static constexpr size_t tot = (ES == 4 ? y + YB : yt + YB);
    static constexpr size_t bar = al16(tot + 2 * NW * sizeof(float2));
    static constexpr size_t total = bar + 8 * sizeof(uint64_t);
};

template <typename in_t, bool SOFTPLUS, int TP>
__global__ void __launch_bounds__(NT, 1) ss2d_fwd_pair_kernel(const FusedFwdPairArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ES = sizeof(in_t);
    constexpr int L = TP * S, PP = NT / TP;
    using Lay = PairLayout<ES, TP>;
    constexpr bool DB = Lay::DB;
    in_t* const sX = reinterpret_cast<in_t*>(smem + Lay::x);
    in_t* const sXt = reinterpret_cast<in_t*>(smem + Lay::xt);
    float* const sY = reinterpret_cast<float*>(smem + Lay::y);
    float* const sYt = reinterpret_cast<float*>(smem + Lay::yt);
    float2* const stot = reinterpret_cast<float2*>(smem + Lay::tot);  // [2 parity][NW]
    uint64_t* const barX = reinterpret_cast<uint64_t*>(smem + Lay::bar);  // [1]
    uint64_t* const barD = barX + 2;                                      // [2]
    auto sD = [&](int buf) { return reinterpret_cast<in_t*>(smem + (buf ? Lay::d1 : Lay::d0)); };
    auto sB = [&](int buf) { return reinterpret_cast<in_t*>(smem + (buf ? Lay::b1 : Lay::b0)); };
    auto sC = [&](int buf) { return reinterpret_cast<in_t*>(smem + (buf ? Lay::c1 : Lay::c0)); };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = tid / TP;
    const int j = tid - q * TP;
    constexpr int width = TP < 32 ? TP : 32;
    const int sl = lane & (width - 1);
    constexpr int nwg = TP > 32 ? (TP >> 5) : 1;
    const int wq = TP > 32 ? (j >> 5) : 0;
    const int gw0 = TP > 32 ? q * nwg : 0;

    const in_t* __restrict__ gx = static_cast<const in_t*>(p.x);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.dts);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bs);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cs);

    const int K = p.K, D = p.D, NS = p.nsteps;
    const int rounds = p.B * D / PP;
    const int my_rounds = (rounds - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                          static_cast<int>(gridDim.x);
    const int nstages = my_rounds * NS;
    if (nstages <= 0) return;

    if (tid == 0) {
        mbar_init(&barX[0], 1), mbar_init(&barD[0], 1), mbar_init(&barD[1], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    auto plane0_of = [&](int i) { return (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP; };
    auto issue_x = [&](int i) {  // thread 0
        mbar_expect_tx(&barX[0], TOK * ES);
        bulk_g2s(sX, gx + plane0_of(i) * L, TOK * ES, &barX[0]);
    };
    auto issue_dir = [&](int sg) {  // thread 0: global step sg = i * NS + st
        const int i = sg / NS, st = sg - i * NS;
        const int k = p.step_k[st];
        const uint32_t pl0 = static_cast<uint32_t>(plane0_of(i));
        const uint32_t b = p.div_d.div(pl0);
        const uint32_t d0 = pl0 - b * static_cast<uint32_t>(D);
        const int buf = DB ? (sg & 1) : 0;
        constexpr uint32_t lb = L * ES;
        mbar_expect_tx(&barD[buf], TOK * ES + 2 * lb);
        bulk_g2s(sD(buf), gd + ((static_cast<long long>(b) * K + k) * D + d0) * L, TOK * ES, &barD[buf]);
        bulk_g2s(sB(buf), gB + (static_cast<long long>(b) * K + k) * L, lb, &barD[buf]);
        bulk_g2s(sC(buf), gC + (static_cast<long long>(b) * K + k) * L, lb, &barD[buf]);
    };
    if (tid == 0) {
        issue_x(0);
        issue_dir(0);
    }

    uint32_t phD0 = 0, phD1 = 0;
    int sg = 0;
    for (int i = 0; i < my_rounds; ++i) {
        const uint32_t pl0 = static_cast<uint32_t>(plane0_of(i));
        const int b = static_cast<int>(p.div_d.div(pl0));
        const int d = static_cast<int>(pl0) - b * D + q;
        mbar_wait(&barX[0], static_cast<uint32_t>(i & 1));
        float ysum[S];  // outputs of the current group, in the base direction's token order
        bool first_merge = true;

        // One step.  REV (compile time): the step's tokens are the base order's tokens back to front: this thread is
        // segment TP-1-j of the direction, its m-th token is element S-1-m of its data chunk, and the carries travel
        // DOWN the lanes.
        auto step = [&](auto rev_tag, const int st) __attribute__((always_inline)) {
            constexpr bool REV = decltype(rev_tag)::value;
            const int k = p.step_k[st];
            const int kind = p.step_kind[st];
            const bool perm = kind >= 2;
            const int buf = DB ? (sg & 1) : 0;
            if (DB && tid == 0 && sg + 1 < nstages) issue_dir(sg + 1);  // the other buffers: free since the last barrier
            const uint2* __restrict__ sched =
                reinterpret_cast<const uint2*>(p.sched + static_cast<size_t>(p.step_base[st]) * p.sched_len) + j;
            const int sched_used = p.sched_used[p.step_base[st]];
            const int sched_iters = sched_used / (2 * TP);
            const bool dense = kind == 2 && sched_used == L;
            uint32_t pix[S / 2];  // kind 3: pixels of this thread's tokens (two 16-bit indices per word), fetched early
            if (kind == 3) {
                const uint32_t* tp = reinterpret_cast<const uint32_t*>(p.idx + static_cast<size_t>(p.step_base[st]) * L + j * S);
#pragma unroll
                for (int m = 0; m < S / 2; ++m) pix[m] = __ldg(tp + m);
            }
            if (p.step_gather[st]) {
                // ---- image order -> token order of the group's base direction
                const in_t* xi = sX + static_cast<size_t>(q) * L;
                in_t* xo = sXt + static_cast<size_t>(q) * L;
                auto move = [&](auto dense_tag, const uint2 ea, const uint2 eb, const uint2 ec, int cnt)
                                __attribute__((always_inline)) {
                    constexpr bool DENSE = decltype(dense_tag)::value;
                    const uint32_t ee[6] = {ea.x, ea.y, eb.x, eb.y, ec.x, ec.y};
                    uint32_t xv[6];
#pragma unroll
                    for (int n = 0; n < 6; ++n) {
                        const uint32_t e = ee[n];
                        const uint32_t src2 = (DENSE ? e : (e == 0xFFFFFFFFu ? 0u : e)) & 0xFFFFu;
                        if (DENSE || (n >> 1) < cnt) {
                            if constexpr (ES == 2) xv[n] = lds_u16(xi, src2);
                            else xv[n] = __float_as_uint(lds_f32(xi, 2 * src2));
                        }
                    }
#pragma unroll
                    for (int n = 0; n < 6; ++n) {
                        const uint32_t e = ee[n];
                        const uint32_t dst2 = e >> 16;
                        if (DENSE || ((n >> 1) < cnt && e != 0xFFFFFFFFu)) {
                            if constexpr (ES == 2) sts_u16(xo, dst2, xv[n]);
                            else sts_f32(xo, 2 * dst2, __uint_as_float(xv[n]));
                        }
                    }
                };
                if (dense) walk_schedule<TP, true>(sched, sched_iters, move);
                else walk_schedule<TP, false>(sched, sched_iters, move);
                __syncthreads();  // token-ordered x complete
                if (p.step_xfree[st] && tid == 0 && i + 1 < my_rounds) issue_x(i + 1);
            }
            const int row = k * D + d;
            const float A2 = __ldg(p.A + row) * kLog2e;
            const float Dd = p.Dp ? __ldg(p.Dp + row) : 0.f;
            const float bias = p.bias ? __ldg(p.bias + row) : 0.f;
            if (buf == 0) {
                mbar_wait(&barD[0], phD0);
                phD0 ^= 1;
            } else {
                mbar_wait(&barD[1], phD1);
                phD1 ^= 1;
            }
            const int jt = REV ? TP - 1 - j : j;  // this thread's segment of the step's direction
            const in_t* xq = (perm ? sXt : sX) + static_cast<size_t>(q) * L + j * S;
            const in_t* dq = sD(buf) + static_cast<size_t>(q) * L + jt * S;
            const in_t* Bq = sB(buf) + jt * S;
            const in_t* Cq = sC(buf) + jt * S;

            // ---- pass 1: local fold from a zero carry-in
            float y0[S], g[S];
            float P = 1.f, Sacc = 0.f;
#pragma unroll
            for (int mp = 0; mp < S / 2; ++mp) {
                float dv[2], Bv[2], Cv[2], uv[2];
                load_pair(dq, mp, dv[0], dv[1]);
                load_pair(Bq, mp, Bv[0], Bv[1]);
                load_pair(Cq, mp, Cv[0], Cv[1]);
                chunk_pair<REV>(xq, mp, uv[0], uv[1]);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int m = 2 * mp + h2;
                    const float raw = dv[h2] + bias;
                    float e;
                    const float dsp = !SOFTPLUS ? raw : (ES == 2 ? softplus20_fast(raw) : softplus20(raw, e));
                    const float a = ex2f(dsp * A2);
                    const float bb = dsp * Bv[h2] * uv[h2];
                    Sacc = fmaf(a, Sacc, bb);
                    P *= a;
                    y0[m] = fmaf(Cv[h2], Sacc, Dd * uv[h2]);
                    g[m] = Cv[h2] * P;
                }
            }
            // ---- carry-in: scan over the plane's threads, up the lanes (or down them for a reversed step)
            float Pe, Se;
            if constexpr (!REV) {
                fused_scan(P, Sacc, sl, width);
                Pe = __shfl_up_sync(0xffffffffu, P, 1, width);
                Se = __shfl_up_sync(0xffffffffu, Sacc, 1, width);
                if (sl == 0) Pe = 1.f, Se = 0.f;
            } else {
                fused_scan_rev(P, Sacc, sl, width);
                Pe = __shfl_down_sync(0xffffffffu, P, 1, width);
                Se = __shfl_down_sync(0xffffffffu, Sacc, 1, width);
                if (sl == width - 1) Pe = 1.f, Se = 0.f;
            }
            const int par = sg & 1;
            if (TP > 32 && lane == (REV ? 0 : 31)) stot[par * NW + warp] = make_float2(P, Sacc);
            if (tid == 0 && sg > 0 && (sg - i * NS) == 0) bulk_wait_read<0>();  // previous item's planes have left sY
            __syncthreads();
            if (!p.step_gather[st] && p.step_xfree[st] && tid == 0 && i + 1 < my_rounds) issue_x(i + 1);
            if (!DB && tid == 0 && sg + 1 < nstages) issue_dir(sg + 1);  // single-buffered: everybody is past pass 1
            float hin = Se;
            if (TP > 32) {
                float hp = 0.f;
                if constexpr (!REV) {
                    for (int w = 0; w < wq; ++w) {
                        const float2 t = stot[par * NW + gw0 + w];
                        hp = fmaf(t.x, hp, t.y);
                    }
                } else {
                    for (int w = nwg - 1; w > wq; --w) {
                        const float2 t = stot[par * NW + gw0 + w];
                        hp = fmaf(t.x, hp, t.y);
                    }
                }
                hin = fmaf(Pe, hp, Se);
            }
            // ---- pass 2: add the carry-in term; sum into the group's registers (base token order)
#pragma unroll
            for (int m = 0; m < S; ++m) {
                const float yv = fmaf(g[m], hin, y0[m]);
                if constexpr (!REV) ysum[m] = yv;
                else ysum[S - 1 - m] += yv;
            }
            if (p.step_merge[st]) {
                if (!perm) {  // identity order: this thread's own pixels, in place
                    float* yq = sY + static_cast<size_t>(q) * L + j * S;
                    if (first_merge) {
                        store_tokens<S>(yq, ysum);
                    } else {
#pragma unroll
                        for (int mp = 0; mp < S / 2; ++mp) {
                            float lo, hi;
                            load_pair(yq, mp, lo, hi);
                            reinterpret_cast<float2*>(yq)[mp] = make_float2(lo + ysum[2 * mp], hi + ysum[2 * mp + 1]);
                        }
                    }
                } else if (kind == 3) {
                    // repeated pixels (helix lines): every token is added to its pixel straight from the registers
                    // with a shared-memory atomic; the pixel of each token comes from the plain order table
                    float* yo = sY + static_cast<size_t>(q) * L;
                    if (first_merge) {  // (no order of the reference starts with such a group)
                        float4* z = reinterpret_cast<float4*>(sY);
                        for (int e = tid; e < TOK / 4; e += NT) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncthreads();
                    }
#pragma unroll
                    for (int mp = 0; mp < S / 2; ++mp) {
                        atomicAdd(yo + (pix[mp] & 0xFFFFu), ysum[2 * mp]);
                        atomicAdd(yo + (pix[mp] >> 16), ysum[2 * mp + 1]);
                    }
                } else {
                    store_tokens<S>(sYt + static_cast<size_t>(q) * L + j * S, ysum);
                    __syncthreads();  // token-ordered outputs complete
                    const float* yi = sYt + static_cast<size_t>(q) * L;
                    float* yo = sY + static_cast<size_t>(q) * L;
                    {
                        const bool init = first_merge;
                        auto move = [&](auto dense_tag, const uint2 ea, const uint2 eb, const uint2 ec, int cnt)
                                        __attribute__((always_inline)) {
                            constexpr bool DENSE = decltype(dense_tag)::value;
                            const uint32_t ee[6] = {ea.x, ea.y, eb.x, eb.y, ec.x, ec.y};
                            float v[6], o[6];
#pragma unroll
                            for (int n = 0; n < 6; ++n) {
                                const uint32_t e = DENSE ? ee[n] : (ee[n] == 0xFFFFFFFFu ? 0u : ee[n]);
                                if (DENSE || (n >> 1) < cnt) {
                                    v[n] = lds_f32(yi, 2 * (e >> 16));
                                    o[n] = init ? 0.f : lds_f32(yo, 2 * (e & 0xFFFFu));
                                }
                            }
#pragma unroll
                            for (int n = 0; n < 6; ++n) {
                                const uint32_t e = ee[n];
                                if (DENSE || ((n >> 1) < cnt && e != 0xFFFFFFFFu)) sts_f32(yo, 2 * (e & 0xFFFFu), o[n] + v[n]);
                            }
                        };
                        if (dense) walk_schedule<TP, true>(sched, sched_iters, move);
                        else walk_schedule<TP, false>(sched, sched_iters, move);
                    }
                    // 16-bit operands: the next group's writes to sYt / sY come after barriers of the next step.
                    // fp32 operands: sYt IS the token-ordered x buffer the next gather overwrites -> close the readers
                    if constexpr (ES == 4) __syncthreads();
                }
                first_merge = false;
            }
            ++sg;
        };
        for (int st = 0; st < NS; ++st) {
            if (p.step_rev[st]) step(std::true_type{}, st);
            else step(std::false_type{}, st);
        }
        // ---- merged planes of this item -> global (PP planes are contiguous in y)
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(p.y + static_cast<long long>(pl0) * L, sY, TOK * 4);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all<0>();
}

template <typename in_t, bool SOFTPLUS, int TP>
cudaError_t launch_tp(const FusedFwdPairArgs& a, cudaStream_t st) {
    using Lay = PairLayout<sizeof(in_t), TP>;
    if (Lay::total > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = ss2d_fwd_pair_kernel<in_t, SOFTPLUS, TP>;
    if (cudaError_t e = ensure_dyn_smem(kern, static_cast<int>(Lay::total)); e != cudaSuccess) return e;
    const int sms = device_sms();
    constexpr int PP = NT / TP;
    FusedFwdPairArgs b = a;
    b.div_d.init(static_cast<uint32_t>(a.D));
    const int rounds = a.B * a.D / PP;
    const int grid = rounds < sms ? rounds : sms;
    kern<<<grid, NT, Lay::total, st>>>(b);
    note_launch();
    return cudaGetLastError();
}

template <typename in_t, bool SOFTPLUS>
cudaError_t launch_t(const FusedFwdPairArgs& a, int TP, cudaStream_t st) {
    switch (TP) {
        case 512: return launch_tp<in_t, SOFTPLUS, 512>(a, st);
        case 128: return launch_tp<in_t, SOFTPLUS, 128>(a, st);
        case 32: return launch_tp<in_t, SOFTPLUS, 32>(a, st);
        case 8: return launch_tp<in_t, SOFTPLUS, 8>(a, st);
        default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace

bool ss2d_fused_fwd_pair_fits(int L, int es) {
    const int TP = L / S;
    if (TP * S != L) return false;
    auto ok = [&](auto lay) { return decltype(lay)::total <= 227 * 1024; };
    switch (TP) {
        case 512: return es == 2 ? ok(PairLayout<2, 512>{}) : ok(PairLayout<4, 512>{});
        case 128: return es == 2 ? ok(PairLayout<2, 128>{}) : ok(PairLayout<4, 128>{});
        case 32: return es == 2 ? ok(PairLayout<2, 32>{}) : ok(PairLayout<4, 32>{});
        case 8: return es == 2 ? ok(PairLayout<2, 8>{}) : ok(PairLayout<4, 8>{});
        default: return false;
    }
}

cudaError_t launch_ss2d_fused_fwd_pair(const FusedFwdPairArgs& a, int L, int in_dtype, cudaStream_t st) {
    const int TP = L / S;
    if (in_dtype == TRAMBA_F32) return a.softplus ? launch_t<float, true>(a, TP, st) : launch_t<float, false>(a, TP, st);
    if (in_dtype == TRAMBA_BF16)
        return a.softplus ? launch_t<__nv_bfloat16, true>(a, TP, st) : launch_t<__nv_bfloat16, false>(a, TP, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
