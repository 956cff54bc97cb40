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
                if (unit < p.units) {
                    const int unit_rows = min(p.R, p.rows - static_cast<int>(unit) * p.R);
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

template <int TSEG, int NS, typename in_t, typename go_t, bool TMA, int RW>
cudaError_t launch_one(const SScanArgs& a, cudaStream_t st) {
    using SM = BwdSmem<TSEG, in_t, go_t>;
    auto kern = sscan_bwd_kernel<TSEG, NS, in_t, go_t, TMA, RW>;
    const int warp_bytes = SM::warp_bytes(RW, a.N);
    const int smem = kWarpsPerCta * warp_bytes;
    static int configured = 0;  // per instantiation; grows with N
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
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
