// Selective-scan backward kernel (sm_100a).  See sscan.cuh for the data model.
// Replaces selective_scan_cuda_oflex.bwd (reference call site Models/SS2D/csms6s.py:920-922).
//
// No states are stored by the forward: every chunk is re-run forward from its checkpointed
// carry-in (one float per state per chunk), then the reverse recurrence
//     dh_t = C_t*dout_t + a_{t+1}*dh_{t+1}
// is evaluated with the same lane / warp / chunk scan hierarchy, walking the chunks of a row
// from last to first.  With E_t := a_t*dh_t the reverse scan uses only a token's own a_t:
//     E_t = a_t*(g_t + E_{t+1}),  dh_t = g_t + E_{t+1},  g_t = C_t*dout_t.
// dB/dC (sums over the channels of a group) are pre-reduced in registers over RW rows that a
// warp walks chunk-by-chunk together, then added to global memory with vector reductions.
#include "sscan.cuh"

#ifndef TRAMBA_SSCAN_BWD_MINB
#define TRAMBA_SSCAN_BWD_MINB 3
#endif

namespace tramba {

namespace {

constexpr int kBwdStages = 2;

template <int TSEG, typename in_t, typename go_t>
struct BwdSmem {
    static constexpr int CH = 32 * TSEG;
    static constexpr int IN_ARR = CH * sizeof(in_t);
    static constexpr int GO_ARR = CH * sizeof(go_t);
    static constexpr int STAGE = 2 * IN_ARR + GO_ARR;
    static constexpr int OFF_OUT = kBwdStages * STAGE;
    static constexpr int OFF_BAR = OFF_OUT + 2 * 2 * IN_ARR;
    static constexpr int OFF_ECARRY = OFF_BAR + 64;
    static constexpr int OFF_ACC = OFF_ECARRY + kMaxState * 4;
    static int warp_bytes(int rw, int N) { return ((OFF_ACC + rw * (2 + N) * 128 + 127) / 128) * 128; }
};

template <int CNT>
__device__ __forceinline__ void red_add_seg(float* __restrict__ dst, const float (&v)[CNT]) {
    if constexpr (CNT % 4 == 0) {
#pragma unroll
        for (int i = 0; i < CNT / 4; ++i)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "f"(v[4 * i]),
                         "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                         : "memory");
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) atomicAdd(dst + i, v[i]);
    }
}

template <int TSEG, int NS, typename in_t, typename go_t, bool TMA, int RW>
__global__ void __launch_bounds__(kWarpsPerCta * 32) sscan_bwd_kernel(const SScanArgs p, const int warp_bytes) {
    using SM = BwdSmem<TSEG, in_t, go_t>;
    constexpr int CH = 32 * TSEG;
    constexpr bool PRE = (NS == 1 && RW > 1);  // register pre-reduction of dB/dC over RW rows
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long wglobal = static_cast<long long>(blockIdx.x) * kWarpsPerCta + warp;
    const long long wtotal = static_cast<long long>(gridDim.x) * kWarpsPerCta;
    const int N = NS ? NS : p.N;
    const int NQ = 2 + N;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* wb = smem_raw + static_cast<size_t>(warp) * warp_bytes;
    unsigned char* s_ring = wb;
    in_t* s_out = reinterpret_cast<in_t*>(wb + SM::OFF_OUT);  // [2][2][CH]
    uint64_t* full = reinterpret_cast<uint64_t*>(wb + SM::OFF_BAR);
    float* s_ecarry = reinterpret_cast<float*>(wb + SM::OFF_ECARRY);  // [RW] (NS==1) or [N]
    float* s_acc = reinterpret_cast<float*>(wb + SM::OFF_ACC);        // [RW][NQ][32]

    const in_t* __restrict__ gu = static_cast<const in_t*>(p.u);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.delta);
    const go_t* __restrict__ ggo = static_cast<const go_t*>(p.dout);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bm);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cm);
    in_t* __restrict__ gdu = static_cast<in_t*>(p.du);
    in_t* __restrict__ gdd = static_cast<in_t*>(p.ddelta);

    const long long sunits = (p.units + RW - 1) / RW;
    const long long my_su = wglobal < sunits ? (sunits - wglobal + wtotal - 1) / wtotal : 0;
    const int tiles_per_su = p.cpu * RW;
    const long long T = my_su * tiles_per_su;
    if (T == 0) return;

    if (p.uniform) {
        for (int i = lane; i < RW * NQ * 32; i += 32) s_acc[i] = 0.f;
    }
    if constexpr (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kBwdStages; ++s) mbar_init(&full[s], 1);
            fence_mbar_init();
            fence_proxy_async_smem();
        }
    }
    __syncwarp();

    auto tile_unit = [&](long long j, int cr, int& chunk) -> long long {
        chunk = p.cpu - 1 - cr / RW;
        return (wglobal + j * wtotal) * RW + (cr % RW);
    };

    long long pq = 0, pj = 0;
    int pcr = 0;
    auto issue = [&]() {
        if (pq < T) {
            if (lane == 0) {
                int chunk;
                const long long unit = tile_unit(pj, pcr, chunk);
                const int unit_rows = unit < p.units ? min(p.R, p.rows - static_cast<int>(unit) * p.R) : 0;
                // a partial last unit (rows % R != 0) still walks p.cpu chunks: chunks past its tokens are empty
                if (unit_rows * p.L - chunk * CH > 0) {
                    const int unit_tokens = unit_rows * p.L;
                    const long long tok0 =
                        unit * p.R * static_cast<long long>(p.L) + static_cast<long long>(chunk) * CH;
                    const int nt = min(CH, unit_tokens - chunk * CH);
                    const int st = static_cast<int>(pq % kBwdStages);
                    unsigned char* sb = s_ring + st * SM::STAGE;
                    mbar_expect_tx(&full[st], nt * (2 * sizeof(in_t) + sizeof(go_t)));
                    bulk_g2s(sb, gu + tok0, nt * sizeof(in_t), &full[st]);
                    bulk_g2s(sb + SM::IN_ARR, gd + tok0, nt * sizeof(in_t), &full[st]);
                    bulk_g2s(sb + 2 * SM::IN_ARR, ggo + tok0, nt * sizeof(go_t), &full[st]);
                } else {
                    mbar_arrive(&full[static_cast<int>(pq % kBwdStages)]);  // empty tail tile
                }
            }
            ++pq;
            if (++pcr == tiles_per_su) pcr = 0, ++pj;
        }
    };
    if constexpr (TMA) {
#pragma unroll
        for (int s = 0; s < kBwdStages - 1; ++s) issue();
    }

    float dBacc[TSEG], dCacc[TSEG];
    long long cj = 0;
    int ccr = 0;
    for (long long cq = 0; cq < T; ++cq) {
        if constexpr (TMA) issue();
        int chunk;
        const long long unit = tile_unit(cj, ccr, chunk);
        const int r = ccr % RW;
        const bool valid = unit < p.units;  // tail super-unit may have fewer rows
        const long long unit_c = valid ? unit : p.units - 1;
        const int unit_rows = min(p.R, p.rows - static_cast<int>(unit_c) * p.R);
        const int unit_tokens = valid ? unit_rows * p.L : 0;
        const long long tok0 = unit_c * p.R * static_cast<long long>(p.L) + static_cast<long long>(chunk) * CH;
        const int nt = max(0, min(CH, unit_tokens - chunk * CH));
        const LaneCoord lc = lane_coord(p, unit_c, chunk, lane, TSEG, CH, unit_tokens);

        float u[TSEG], dl[TSEG], go[TSEG], sg[TSEG];
        if constexpr (TMA) {
            const int st = static_cast<int>(cq % kBwdStages);
            mbar_wait(&full[st], static_cast<uint32_t>((cq / kBwdStages) & 1));
            if (lc.active) {
                const unsigned char* sb = s_ring + st * SM::STAGE;
                load_seg<TSEG>(reinterpret_cast<const in_t*>(sb) + lane * TSEG, u);
                load_seg<TSEG>(reinterpret_cast<const in_t*>(sb + SM::IN_ARR) + lane * TSEG, dl);
                load_seg<TSEG>(reinterpret_cast<const go_t*>(sb + 2 * SM::IN_ARR) + lane * TSEG, go);
            }
        } else {
            if (lc.active) {
                load_seg<TSEG>(gu + tok0 + lane * TSEG, u);
                load_seg<TSEG>(gd + tok0 + lane * TSEG, dl);
                load_seg<TSEG>(ggo + tok0 + lane * TSEG, go);
            }
        }
        const float bias = p.bias ? __ldg(p.bias + lc.d) : 0.f;
        const float Dd = p.D ? __ldg(p.D + lc.d) : 0.f;
        float duv[TSEG], dd[TSEG];
        float tD = 0.f;
#pragma unroll
        for (int i = 0; i < TSEG; ++i) {
            if (lc.active) {
                float x = dl[i] + bias;
                if (p.softplus) {
                    float e;
                    dl[i] = softplus20(x, e);
                    sg[i] = (x > 20.f) ? 1.f : e * rcpf(1.f + e);
                } else {
                    dl[i] = x;
                    sg[i] = 1.f;
                }
                duv[i] = Dd * go[i];
                tD = fmaf(go[i], u[i], tD);
            } else {
                u[i] = 0.f, dl[i] = 0.f, go[i] = 0.f, sg[i] = 0.f, duv[i] = 0.f;
            }
            dd[i] = 0.f;
        }
        if (PRE && r == 0) {
#pragma unroll
            for (int i = 0; i < TSEG; ++i) dBacc[i] = 0.f, dCacc[i] = 0.f;
        }
        const bool last_chunk = (chunk == p.cpu - 1);
        for (int n = 0; n < N; ++n) {
            const float An = __ldg(p.A + static_cast<long long>(lc.d) * N + n);
            const float An2 = An * kLog2e;
            float a[TSEG], hp[TSEG], Bv[TSEG], g[TSEG];
            if (lc.active) {
                load_seg<TSEG>(gB + lc.bc + static_cast<long long>(n) * p.L, Bv);
                load_seg<TSEG>(gC + lc.bc + static_cast<long long>(n) * p.L, g);
            } else {
#pragma unroll
                for (int i = 0; i < TSEG; ++i) Bv[i] = 0.f, g[i] = 0.f;
            }
            // ---- forward replay of the chunk: local fold, warp scan, apply (keeps h_{t-1})
            float P = 1.f, S = 0.f;
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                a[i] = ex2f(dl[i] * An2);
                if (i == 0 && lc.rowstart) a[0] = 0.f;
                S = fmaf(a[i], S, dl[i] * u[i] * Bv[i]);
                P *= a[i];
            }
            const float Ploc = P;
            warp_scan_affine(P, S, lane);
            float Pe = __shfl_up_sync(0xffffffffu, P, 1);
            float Se = __shfl_up_sync(0xffffffffu, S, 1);
            if (lane == 0) Pe = 1.f, Se = 0.f;
            const float cin = (chunk > 0 && valid) ? __ldg(p.ckpt_in + (unit_c * p.cpu + chunk - 1) * N + n) : 0.f;
            float h = fmaf(Pe, cin, Se);
            // ---- reverse: local fold of E = a*(g + E_next), warp scan from the right
            float Sr = 0.f;
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                hp[i] = h;
                h = fmaf(a[i], h, dl[i] * u[i] * Bv[i]);
                g[i] *= go[i];
            }
#pragma unroll
            for (int i = TSEG - 1; i >= 0; --i) Sr = a[i] * (Sr + g[i]);
            float Pr = Ploc;
            warp_scan_affine_rev(Pr, Sr, lane);
            float Pre = __shfl_down_sync(0xffffffffu, Pr, 1);
            float Sre = __shfl_down_sync(0xffffffffu, Sr, 1);
            if (lane == 31) Pre = 1.f, Sre = 0.f;
            const float ein = last_chunk ? 0.f : s_ecarry[NS == 1 ? r : n];
            float E = fmaf(Pre, ein, Sre);
            float tA = 0.f;
            float dBv[TSEG], dCv[TSEG];
#pragma unroll
            for (int i = TSEG - 1; i >= 0; --i) {
                const float dh = g[i] + E;
                E = a[i] * dh;
                const float t = dh * dl[i];
                const float ahp = a[i] * hp[i];
                const float bu = Bv[i] * u[i];
                duv[i] = fmaf(t, Bv[i], duv[i]);
                dd[i] = fmaf(dh, fmaf(An, ahp, bu), dd[i]);
                tA = fmaf(t, ahp, tA);
                dBv[i] = t * u[i];
                dCv[i] = go[i] * fmaf(dl[i], bu, ahp);  // dout_t * h_t
            }
            const float eout = __shfl_sync(0xffffffffu, E, 0);
            __syncwarp();
            if (lane == 0) s_ecarry[NS == 1 ? r : n] = eout;
            // ---- dA partial
            if (p.uniform) {
                s_acc[(r * NQ + 2 + n) * 32 + lane] += tA;
            } else {
                float v = tA;
                if (warp_seg_sum(v, lc.row, lane) && valid) atomicAdd(p.dA + static_cast<long long>(lc.d) * N + n, v);
            }
            // ---- dB / dC
            if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < TSEG; ++i) dBacc[i] += dBv[i], dCacc[i] += dCv[i];
            } else {
                if (lc.active) {
                    red_add_seg<TSEG>(p.dB + lc.bc + static_cast<long long>(n) * p.L, dBv);
                    red_add_seg<TSEG>(p.dC + lc.bc + static_cast<long long>(n) * p.L, dCv);
                }
            }
        }
        if (PRE && r == RW - 1 && lc.active) {
            red_add_seg<TSEG>(p.dB + lc.bc, dBacc);
            red_add_seg<TSEG>(p.dC + lc.bc, dCacc);
        }
        float tb = 0.f;
#pragma unroll
        for (int i = 0; i < TSEG; ++i) {
            dd[i] *= sg[i];
            tb += dd[i];
        }
        if (p.uniform) {
            s_acc[(r * NQ + 0) * 32 + lane] += tD;
            s_acc[(r * NQ + 1) * 32 + lane] += tb;
        } else {
            float v = tD;
            bool tail = warp_seg_sum(v, lc.row, lane);
            if (tail && valid && p.dD) atomicAdd(p.dD + lc.d, v);
            v = tb;
            tail = warp_seg_sum(v, lc.row, lane);
            if (tail && valid && p.dbias) atomicAdd(p.dbias + lc.d, v);
        }
        // ---- store du, ddelta
        if constexpr (TMA) {
            const int ob = static_cast<int>(cq & 1);
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            if (lc.active) {
                store_seg<TSEG>(s_out + (ob * 2 + 0) * CH + lane * TSEG, duv);
                store_seg<TSEG>(s_out + (ob * 2 + 1) * CH + lane * TSEG, dd);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (nt > 0) {
                    bulk_s2g(gdu + tok0, s_out + (ob * 2 + 0) * CH, nt * sizeof(in_t));
                    bulk_s2g(gdd + tok0, s_out + (ob * 2 + 1) * CH, nt * sizeof(in_t));
                }
                bulk_commit();
            }
        } else {
            if (lc.active) {
                store_seg<TSEG>(gdu + tok0 + lane * TSEG, duv);
                store_seg<TSEG>(gdd + tok0 + lane * TSEG, dd);
            }
            __syncwarp();
        }
        // ---- end of a super-unit: flush the per-row accumulators (uniform plans)
        if (++ccr == tiles_per_su) {
            if (p.uniform) {
                __syncwarp();
                const long long su = wglobal + cj * wtotal;
                for (int rr = 0; rr < RW; ++rr) {
                    const long long urow = su * RW + rr;  // uniform: unit == row
                    if (urow >= p.units) break;
                    const int b = static_cast<int>(p.divDim.div(static_cast<uint32_t>(urow)));
                    const int d = static_cast<int>(urow) - b * p.dim;
                    for (int q = 0; q < NQ; ++q) {
                        float v = s_acc[(rr * NQ + q) * 32 + lane];
                        s_acc[(rr * NQ + q) * 32 + lane] = 0.f;
                        v = warp_sum(v);
                        if (lane == 0) {
                            if (q == 0) {
                                if (p.dD) atomicAdd(p.dD + d, v);
                            } else if (q == 1) {
                                if (p.dbias) atomicAdd(p.dbias + d, v);
                            } else {
                                atomicAdd(p.dA + static_cast<long long>(d) * N + (q - 2), v);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            ccr = 0, ++cj;
        }
    }
    if constexpr (TMA) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// Regular plans (L % CH == 0, or CH = rpc * L): the lean backward.
//   * a warp owns a super-unit of RW consecutive units of one (batch, group) and walks it
//     chunk-outer (last chunk first) / unit-inner, so the dB/dC contributions of the RW rows
//     add up in registers and reach global memory as one vector reduction per chunk;
//   * 3-stage TMA ring [u | delta | dout]; du / ddelta overwrite the u / delta slots of their
//     own stage and leave with a bulk store, so no separate output buffers are needed;
//   * per-row constants sit in a small shared table, per-row partial sums of dA / dD / dbias in
//     per-lane shared accumulators that are reduced once per super-unit;
//   * checkpoints and B / C rows are fetched one tile ahead.
// ---------------------------------------------------------------------------------------
constexpr int kRegStages = 3;

template <int TSEG, typename in_t, typename go_t>
struct RegSmem {
    static constexpr int CH = 32 * TSEG;
    static constexpr int IN_ARR = CH * sizeof(in_t);
    static constexpr int GO_ARR = CH * sizeof(go_t);
    static constexpr int STAGE = 2 * IN_ARR + GO_ARR;
    static constexpr int OFF_BAR = kRegStages * STAGE;
    static constexpr int OFF_ECARRY = OFF_BAR + 64;                // [max(RW, N)] floats
    static constexpr int OFF_CTX = OFF_ECARRY + kMaxState * 4;     // [RW][RPC] float4 (A, bias, D, -)
    __host__ __device__ static int off_acc(int rw, int rpc) { return OFF_CTX + rw * rpc * 16; }
    static int warp_bytes(int rw, int rpc, int N) {
        return ((off_acc(rw, rpc) + rw * (2 + N) * 128 + 127) / 128) * 128;
    }
};

template <int TSEG, int NS, typename in_t, typename go_t, int LPR, int RW>
__global__ void __launch_bounds__(kWarpsPerCta * 32, TRAMBA_SSCAN_BWD_MINB)
    sscan_bwd_regular_kernel(const SScanArgs p, const int warp_bytes) {
    using SM = RegSmem<TSEG, in_t, go_t>;
    constexpr int CH = 32 * TSEG;
    constexpr int RPC = 32 / LPR;
    constexpr bool PRE = (NS == 1);  // dB / dC accumulate in registers over the RW units
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane / LPR;
    const int sl = lane & (LPR - 1);
    const int wglobal = blockIdx.x * kWarpsPerCta + warp;
    const int wtotal = gridDim.x * kWarpsPerCta;
    const int N = NS ? NS : p.N;
    const int NQ = 2 + N;
    const int cpu = p.cpu;
    const int L = p.L;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* wb = smem_raw + static_cast<size_t>(warp) * warp_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(wb + SM::OFF_BAR);
    float* s_ecarry = reinterpret_cast<float*>(wb + SM::OFF_ECARRY);
    float4* s_ctx = reinterpret_cast<float4*>(wb + SM::OFF_CTX);
    float* s_acc = reinterpret_cast<float*>(wb + SM::off_acc(RW, RPC));

    const in_t* __restrict__ gu = static_cast<const in_t*>(p.u);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.delta);
    const go_t* __restrict__ ggo = static_cast<const go_t*>(p.dout);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bm);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cm);
    in_t* __restrict__ gdu = static_cast<in_t*>(p.du);
    in_t* __restrict__ gdd = static_cast<in_t*>(p.ddelta);

    const int sunits = static_cast<int>(p.units) / RW;  // host guarantees divisibility
    if (wglobal >= sunits) return;
    const int my_su = (sunits - wglobal + wtotal - 1) / wtotal;
    const int tiles_per_su = cpu * RW;
    const int T = my_su * tiles_per_su;
    const long long unit_tok = static_cast<long long>(RPC) * L;
    const long long su_tok = unit_tok * RW;
    const long long stride_tok = su_tok * wtotal;

    for (int i = lane; i < RW * NQ * 32; i += 32) s_acc[i] = 0.f;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kRegStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();

    // ---- producer: tile order inside a super-unit is chunk (descending) outer, unit inner
    int pq = 0, pc = cpu - 1, pr = 0, pst = 0;
    long long ptok = wglobal * su_tok;
    auto issue = [&]() {
        if (pq < T) {
            if (lane == 0) {
                bulk_wait_read<1>();  // the stage being refilled was stored from two tiles ago
                const long long t0 = ptok + pr * unit_tok + pc * CH;
                unsigned char* sb = wb + pst * SM::STAGE;
                mbar_expect_tx(&full[pst], SM::STAGE);
                bulk_g2s(sb, gu + t0, SM::IN_ARR, &full[pst]);
                bulk_g2s(sb + SM::IN_ARR, gd + t0, SM::IN_ARR, &full[pst]);
                bulk_g2s(sb + 2 * SM::IN_ARR, ggo + t0, SM::GO_ARR, &full[pst]);
            }
            ++pq;
            if (++pr == RW) {
                pr = 0;
                if (--pc < 0) pc = cpu - 1, ptok += stride_tok;
            }
            if (++pst == kRegStages) pst = 0;
        }
    };
    issue();

    // ---- per super-unit setup: row constants -> shared table; B / C base offset
    long long bc = 0;  // element offset of (b, g, n = 0, t = sl*TSEG); same for all RW units
    auto setup_su = [&](int su) {
        __syncwarp();
        if (lane < RW * RPC) {
            const int row = (su * RW + lane / RPC) * RPC + (lane % RPC);
            const int b = static_cast<int>(p.divDim.div(static_cast<uint32_t>(row)));
            const int d = row - b * p.dim;
            s_ctx[lane] = make_float4((NS == 1) ? __ldg(p.A + d) : 0.f, p.bias ? __ldg(p.bias + d) : 0.f,
                                      p.D ? __ldg(p.D + d) : 0.f, __int_as_float(d));
        }
        const int row0 = su * RW * RPC;
        const int b0 = static_cast<int>(p.divDim.div(static_cast<uint32_t>(row0)));
        const int d0 = row0 - b0 * p.dim;
        const int g0 = static_cast<int>(p.divDpg.div(static_cast<uint32_t>(d0)));
        bc = (static_cast<long long>(b0) * p.groups + g0) * N * static_cast<long long>(L) + sl * TSEG;
        __syncwarp();
    };
    // flush per-row sums of dD / dbias / dA of a finished super-unit
    auto flush_su = [&](int su) {
        __syncwarp();
        for (int r = 0; r < RW; ++r) {
            for (int q = 0; q < NQ; ++q) {
                float v = s_acc[(r * NQ + q) * 32 + lane];
                s_acc[(r * NQ + q) * 32 + lane] = 0.f;
                v = group_sum<LPR>(v);
                if (sl == 0) {
                    const int d = __float_as_int(s_ctx[r * RPC + sub].w);
                    if (q == 0) {
                        if (p.dD) atomicAdd(p.dD + d, v);
                    } else if (q == 1) {
                        if (p.dbias) atomicAdd(p.dbias + d, v);
                    } else {
                        atomicAdd(p.dA + static_cast<long long>(d) * N + (q - 2), v);
                    }
                }
            }
        }
    };

    int su = wglobal;
    setup_su(su);
    float dBacc[TSEG], dCacc[TSEG];
    int c = cpu - 1, r = 0, cst = 0;
    uint32_t cph = 0;
    long long tok = wglobal * su_tok;
    // checkpoint (carry-in of the chunk) of the next tile, fetched one tile ahead (NS == 1)
    float cin_next = (NS == 1 && cpu > 1) ? __ldg(p.ckpt_in + (static_cast<long long>(su) * RW * cpu + c - 1)) : 0.f;

    for (int q = 0; q < T; ++q) {
        issue();
        const float4 cx = s_ctx[r * RPC + sub];
        const float bias = cx.y, Dd = cx.z;
        const long long unit = static_cast<long long>(su) * RW + r;  // == row when RPC == 1
        const long long boff = bc + c * CH;
        // B / C of this chunk: L1-resident after the first unit of the chunk touched them
        float Bv[TSEG], g[TSEG];
        if (NS == 1) {
            load_seg<TSEG>(gB + boff, Bv);
            load_seg<TSEG>(gC + boff, g);
        }
        const float cin1 = cin_next;
        if (NS == 1 && cpu > 1) {  // prefetch the next tile's checkpoint
            int nr = r + 1, nc = c;
            long long nsu = su;
            if (nr == RW) {
                nr = 0;
                if (--nc < 0) nc = cpu - 1, nsu += wtotal;
            }
            cin_next = (q + 1 < T && nc > 0) ? __ldg(p.ckpt_in + ((nsu * RW + nr) * cpu + nc - 1)) : 0.f;
        }

        unsigned char* sb = wb + cst * SM::STAGE;
        in_t* su_ = reinterpret_cast<in_t*>(sb);
        in_t* sd_ = reinterpret_cast<in_t*>(sb + SM::IN_ARR);
        const go_t* sg_ = reinterpret_cast<const go_t*>(sb + 2 * SM::IN_ARR);
        mbar_wait(&full[cst], cph);
        float dl[TSEG], sg[TSEG], duv[TSEG], dd[TSEG];
        float tD = 0.f;
        {
            float u[TSEG], go[TSEG];
            load_seg<TSEG>(su_ + lane * TSEG, u);
            load_seg<TSEG>(sd_ + lane * TSEG, dl);
            load_seg<TSEG>(sg_ + lane * TSEG, go);
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                const float x = dl[i] + bias;
                if (p.softplus) {
                    float e, w;
                    dl[i] = softplus20(x, e, w);
                    sg[i] = (x > 20.f) ? 1.f : e * rcpf(w);
                } else {
                    dl[i] = x;
                    sg[i] = 1.f;
                }
                duv[i] = Dd * go[i];
                tD = fmaf(go[i], u[i], tD);
                dd[i] = 0.f;
            }
        }
        if (PRE && r == 0) {
#pragma unroll
            for (int i = 0; i < TSEG; ++i) dBacc[i] = 0.f, dCacc[i] = 0.f;
        }
        const bool last_chunk = (c == cpu - 1);
        for (int n = 0; n < N; ++n) {
            float An;
            if (NS == 1) {
                An = cx.x;
            } else {
                An = __ldg(p.A + static_cast<long long>(__float_as_int(cx.w)) * N + n);
                load_seg<TSEG>(gB + boff + static_cast<long long>(n) * L, Bv);
                load_seg<TSEG>(gC + boff + static_cast<long long>(n) * L, g);
            }
            const float An2 = An * kLog2e;
            float a[TSEG], hp[TSEG];
            // ---- forward replay: local fold, group scan, apply (keeps h_{t-1})
            float P = 1.f, S = 0.f;
            {
                float u[TSEG];
                load_seg<TSEG>(su_ + lane * TSEG, u);
#pragma unroll
                for (int i = 0; i < TSEG; ++i) {
                    a[i] = ex2f(dl[i] * An2);
                    hp[i] = dl[i] * u[i] * Bv[i];  // b_t, parked in hp until the apply pass
                    S = fmaf(a[i], S, hp[i]);
                    P *= a[i];
                }
            }
            const float Ploc = P;
            group_scan_affine<LPR>(P, S, sl);
            float Pe = __shfl_up_sync(0xffffffffu, P, 1, LPR);
            float Se = __shfl_up_sync(0xffffffffu, S, 1, LPR);
            if (sl == 0) Pe = 1.f, Se = 0.f;
            float cin = 0.f;
            if (cpu > 1 && c > 0) cin = (NS == 1) ? cin1 : __ldg(p.ckpt_in + (unit * cpu + c - 1) * N + n);
            float h = fmaf(Pe, cin, Se);
            float Sr = 0.f;
            {
                float go[TSEG];
                load_seg<TSEG>(sg_ + lane * TSEG, go);
#pragma unroll
                for (int i = 0; i < TSEG; ++i) {
                    const float b = hp[i];
                    hp[i] = h;
                    h = fmaf(a[i], h, b);
                    g[i] *= go[i];
                }
            }
            // ---- reverse: E = a*(g + E_next); group scan from the right
#pragma unroll
            for (int i = TSEG - 1; i >= 0; --i) Sr = a[i] * (Sr + g[i]);
            float Pr = Ploc;
            group_scan_affine_rev<LPR>(Pr, Sr, sl);
            float Pre = __shfl_down_sync(0xffffffffu, Pr, 1, LPR);
            float Sre = __shfl_down_sync(0xffffffffu, Sr, 1, LPR);
            if (sl == LPR - 1) Pre = 1.f, Sre = 0.f;
            const float ein = (last_chunk || cpu == 1) ? 0.f : s_ecarry[NS == 1 ? r : n];
            float E = fmaf(Pre, ein, Sre);
            float tA = 0.f;
            {
                float u[TSEG], go[TSEG];
                load_seg<TSEG>(su_ + lane * TSEG, u);
                load_seg<TSEG>(sg_ + lane * TSEG, go);
                float dBv[TSEG], dCv[TSEG];
#pragma unroll
                for (int i = TSEG - 1; i >= 0; --i) {
                    const float dh = g[i] + E;
                    E = a[i] * dh;
                    const float t = dh * dl[i];
                    const float ahp = a[i] * hp[i];
                    const float bu = Bv[i] * u[i];
                    duv[i] = fmaf(t, Bv[i], duv[i]);
                    dd[i] = fmaf(dh, fmaf(An, ahp, bu), dd[i]);
                    tA = fmaf(t, ahp, tA);
                    dBv[i] = t * u[i];
                    dCv[i] = go[i] * fmaf(dl[i], bu, ahp);  // dout_t * h_t
                }
                if constexpr (PRE) {
#pragma unroll
                    for (int i = 0; i < TSEG; ++i) dBacc[i] += dBv[i], dCacc[i] += dCv[i];
                } else {
                    red_add_seg<TSEG>(p.dB + boff + static_cast<long long>(n) * L, dBv);
                    red_add_seg<TSEG>(p.dC + boff + static_cast<long long>(n) * L, dCv);
                }
            }
            if (cpu > 1) {
                const float eout = __shfl_sync(0xffffffffu, E, 0);
                __syncwarp();
                if (lane == 0) s_ecarry[NS == 1 ? r : n] = eout;
            }
            s_acc[(r * NQ + 2 + n) * 32 + lane] += tA;
        }
        if (PRE && r == RW - 1) {
            if constexpr (RPC > 1) {  // rows side by side in the warp share the same tokens
#pragma unroll
                for (int i = 0; i < TSEG; ++i) {
#pragma unroll
                    for (int o = LPR; o < 32; o <<= 1) {
                        dBacc[i] += __shfl_xor_sync(0xffffffffu, dBacc[i], o);
                        dCacc[i] += __shfl_xor_sync(0xffffffffu, dCacc[i], o);
                    }
                }
            }
            if (sub == 0) {
                red_add_seg<TSEG>(p.dB + boff, dBacc);
                red_add_seg<TSEG>(p.dC + boff, dCacc);
            }
        }
        float tb = 0.f;
#pragma unroll
        for (int i = 0; i < TSEG; ++i) {
            dd[i] *= sg[i];
            tb += dd[i];
        }
        s_acc[(r * NQ + 0) * 32 + lane] += tD;
        s_acc[(r * NQ + 1) * 32 + lane] += tb;
        // ---- du / ddelta overwrite the u / delta slots of this stage, then leave by bulk store
        __syncwarp();
        store_seg<TSEG>(su_ + lane * TSEG, duv);
        store_seg<TSEG>(sd_ + lane * TSEG, dd);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            const long long t0 = tok + r * unit_tok + c * CH;
            bulk_s2g(gdu + t0, su_, SM::IN_ARR);
            bulk_s2g(gdd + t0, sd_, SM::IN_ARR);
            bulk_commit();
        }
        if (++cst == kRegStages) cst = 0, cph ^= 1;
        if (++r == RW) {
            r = 0;
            if (--c < 0) {
                flush_su(su);
                c = cpu - 1, su += wtotal, tok += stride_tok;
                if (q + 1 < T) setup_su(su);
            }
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
}

template <int TSEG, int NS, typename in_t, typename go_t, int LPR, int RW>
cudaError_t launch_regular(const SScanArgs& a, cudaStream_t st) {
    using SM = RegSmem<TSEG, in_t, go_t>;
    auto kern = sscan_bwd_regular_kernel<TSEG, NS, in_t, go_t, LPR, RW>;
    const int warp_bytes = SM::warp_bytes(RW, 32 / LPR, a.N);
    const int smem = kWarpsPerCta * warp_bytes;
    if (cudaError_t e = ensure_dyn_smem(kern, smem); e != cudaSuccess) return e;
    const int sms = device_sms();
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem);
    if (occ < 1) occ = 1;
    const long long sunits = a.units / RW;
    long long want = (sunits + kWarpsPerCta - 1) / kWarpsPerCta;
    long long cap = static_cast<long long>(sms) * occ;
    int grid = static_cast<int>(want < cap ? want : cap);
    kern<<<grid, kWarpsPerCta * 32, smem, st>>>(a, warp_bytes);
    note_launch();
    return cudaGetLastError();
}

template <int TSEG, typename in_t, typename go_t, int LPR>
cudaError_t launch_regular_rw(const SScanArgs& a, cudaStream_t st) {
    if (a.N != 1) return launch_regular<TSEG, 0, in_t, go_t, LPR, 1>(a, st);
    switch (a.rw) {
        case 8: return launch_regular<TSEG, 1, in_t, go_t, LPR, 8>(a, st);
        case 4: return launch_regular<TSEG, 1, in_t, go_t, LPR, 4>(a, st);
        case 2: return launch_regular<TSEG, 1, in_t, go_t, LPR, 2>(a, st);
        default: return launch_regular<TSEG, 1, in_t, go_t, LPR, 1>(a, st);
    }
}

template <int TSEG, int NS, typename in_t, typename go_t, bool TMA, int RW>
cudaError_t launch_one(const SScanArgs& a, cudaStream_t st) {
    using SM = BwdSmem<TSEG, in_t, go_t>;
    auto kern = sscan_bwd_kernel<TSEG, NS, in_t, go_t, TMA, RW>;
    const int warp_bytes = SM::warp_bytes(RW, a.N);
    const int smem = kWarpsPerCta * warp_bytes;
    if (cudaError_t e = ensure_dyn_smem(kern, smem); e != cudaSuccess) return e;
    const int sms = device_sms();
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem);
    if (occ < 1) occ = 1;
    const long long sunits = (a.units + RW - 1) / RW;
    long long want = (sunits + kWarpsPerCta - 1) / kWarpsPerCta;
    long long cap = static_cast<long long>(sms) * occ * 2;
    int grid = static_cast<int>(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, kWarpsPerCta * 32, smem, st>>>(a, warp_bytes);
    note_launch();
    return cudaGetLastError();
}

template <int TSEG, bool TMA, typename in_t, typename go_t>
cudaError_t launch_ns(const SScanArgs& a, cudaStream_t st) {
    if (a.N == 1) {
        if (a.rw == 8) return launch_one<TSEG, 1, in_t, go_t, TMA, 8>(a, st);
        return launch_one<TSEG, 1, in_t, go_t, TMA, 1>(a, st);
    }
    return launch_one<TSEG, 0, in_t, go_t, TMA, 1>(a, st);
}

template <typename in_t, typename go_t>
cudaError_t launch_t(const SScanArgs& a, const SScanPlan& p, cudaStream_t st) {
    if (p.regular && a.units < (1ll << 30)) {
        if (p.rpc == 1) {
            if (p.tseg == 12) return launch_regular_rw<12, in_t, go_t, 32>(a, st);
            if (p.tseg == 9) return launch_regular_rw<9, in_t, go_t, 32>(a, st);
            return launch_regular_rw<4, in_t, go_t, 32>(a, st);
        }
        if (p.tseg == 9 && p.rpc == 2) return launch_regular_rw<9, in_t, go_t, 16>(a, st);
        if (p.tseg == 9 && p.rpc == 4) return launch_regular_rw<9, in_t, go_t, 8>(a, st);
        if (p.tseg == 4 && p.rpc == 2) return launch_regular_rw<4, in_t, go_t, 16>(a, st);
        if (p.tseg == 4 && p.rpc == 4) return launch_regular_rw<4, in_t, go_t, 8>(a, st);
    }
    if (!p.tma) return launch_ns<1, false, in_t, go_t>(a, st);
    switch (p.tseg) {
        case 12: return launch_ns<12, true, in_t, go_t>(a, st);
        case 9: return launch_ns<9, true, in_t, go_t>(a, st);
        default: return launch_ns<4, true, in_t, go_t>(a, st);
    }
}

}  // namespace

cudaError_t launch_sscan_bwd(const SScanArgs& a, const SScanPlan& p, int in_dtype, int dout_dtype,
                             cudaStream_t st) {
    if (in_dtype == TRAMBA_F32 && dout_dtype == TRAMBA_F32) return launch_t<float, float>(a, p, st);
    if (in_dtype == TRAMBA_BF16 && dout_dtype == TRAMBA_F32) return launch_t<__nv_bfloat16, float>(a, p, st);
    if (in_dtype == TRAMBA_BF16 && dout_dtype == TRAMBA_BF16)
        return launch_t<__nv_bfloat16, __nv_bfloat16>(a, p, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
