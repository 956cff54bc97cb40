// Selective-scan forward kernel (sm_100a).  See sscan.cuh for the data model.
// Replaces selective_scan_cuda_oflex.fwd (reference call site Models/SS2D/csms6s.py:910).
#include "sscan.cuh"

namespace tramba {

namespace {

constexpr int kFwdStages = 3;

template <int TSEG, typename in_t, typename out_t>
struct FwdSmem {
    static constexpr int CH = 32 * TSEG;
    static constexpr int IN_ARR = CH * sizeof(in_t);
    static constexpr int OUT_ARR = CH * sizeof(out_t);
    static constexpr int OFF_OUT = kFwdStages * 2 * IN_ARR;
    static constexpr int OFF_BAR = OFF_OUT + 2 * OUT_ARR;
    static constexpr int OFF_CARRY = OFF_BAR + 64;
    static constexpr int WARP_BYTES = ((OFF_CARRY + kMaxState * 4 + 127) / 128) * 128;
};

template <int TSEG, int NS, typename in_t, typename out_t, bool TMA>
__global__ void __launch_bounds__(kWarpsPerCta * 32) sscan_fwd_kernel(const SScanArgs p) {
    using SM = FwdSmem<TSEG, in_t, out_t>;
    constexpr int CH = 32 * TSEG;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long wglobal = static_cast<long long>(blockIdx.x) * kWarpsPerCta + warp;
    const long long wtotal = static_cast<long long>(gridDim.x) * kWarpsPerCta;
    const int N = NS ? NS : p.N;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* wb = smem_raw + (TMA ? warp * SM::WARP_BYTES : warp * (kMaxState * 4));
    in_t* s_in = reinterpret_cast<in_t*>(wb);
    out_t* s_out = reinterpret_cast<out_t*>(wb + SM::OFF_OUT);
    uint64_t* full = reinterpret_cast<uint64_t*>(wb + SM::OFF_BAR);
    float* s_carry = reinterpret_cast<float*>(TMA ? wb + SM::OFF_CARRY : wb);

    const in_t* __restrict__ gu = static_cast<const in_t*>(p.u);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.delta);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bm);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cm);
    out_t* __restrict__ gout = static_cast<out_t*>(p.out);

    const long long my_units = wglobal < p.units ? (p.units - wglobal + wtotal - 1) / wtotal : 0;
    const long long T = my_units * p.cpu;
    if (T == 0) return;

    if constexpr (TMA) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kFwdStages; ++s) mbar_init(&full[s], 1);
            fence_mbar_init();
            fence_proxy_async_smem();
        }
        __syncwarp();
    }

    // tile -> (unit, chunk) with incremental counters for producer (pj, pc) and consumer (cj, cc)
    long long pq = 0, pj = 0;
    int pc = 0;
    auto issue = [&]() {
        if (pq < T) {
            if (lane == 0) {
                const long long unit = wglobal + pj * wtotal;
                const int unit_rows = min(p.R, p.rows - static_cast<int>(unit) * p.R);
                const int unit_tokens = unit_rows * p.L;
                const long long tok0 = unit * p.R * static_cast<long long>(p.L) + static_cast<long long>(pc) * CH;
                // a partial last unit (rows % R != 0) still walks p.cpu chunks: chunks past its tokens are empty
                const int nt = max(0, min(CH, unit_tokens - pc * CH));
                const uint32_t bytes = nt * sizeof(in_t);
                const int st = static_cast<int>(pq % kFwdStages);
                if (nt > 0) {
                    mbar_expect_tx(&full[st], 2 * bytes);
                    bulk_g2s(s_in + (st * 2 + 0) * CH, gu + tok0, bytes, &full[st]);
                    bulk_g2s(s_in + (st * 2 + 1) * CH, gd + tok0, bytes, &full[st]);
                } else {
                    mbar_arrive(&full[st]);  // empty tail tile: complete the phase without a copy
                }
            }
            ++pq;
            if (++pc == p.cpu) pc = 0, ++pj;
        }
    };
    if constexpr (TMA) {
#pragma unroll
        for (int s = 0; s < kFwdStages - 1; ++s) issue();
    }

    float carry1 = 0.f;  // NS == 1: running chunk carry in a register
    long long cj = 0;
    int cc = 0;
    for (long long cq = 0; cq < T; ++cq) {
        if constexpr (TMA) issue();
        const long long unit = wglobal + cj * wtotal;
        const int unit_rows = min(p.R, p.rows - static_cast<int>(unit) * p.R);
        const int unit_tokens = unit_rows * p.L;
        const long long tok0 = unit * p.R * static_cast<long long>(p.L) + static_cast<long long>(cc) * CH;
        const int nt = max(0, min(CH, unit_tokens - cc * CH));
        const LaneCoord lc = lane_coord(p, unit, cc, lane, TSEG, CH, unit_tokens);

        float u[TSEG], dl[TSEG];
        if constexpr (TMA) {
            const int st = static_cast<int>(cq % kFwdStages);
            mbar_wait(&full[st], static_cast<uint32_t>((cq / kFwdStages) & 1));
            if (lc.active) {
                load_seg<TSEG>(s_in + (st * 2 + 0) * CH + lane * TSEG, u);
                load_seg<TSEG>(s_in + (st * 2 + 1) * CH + lane * TSEG, dl);
            }
        } else {
            if (lc.active) {
                load_seg<TSEG>(gu + tok0 + lane * TSEG, u);
                load_seg<TSEG>(gd + tok0 + lane * TSEG, dl);
            }
        }
        const float bias = p.bias ? __ldg(p.bias + lc.d) : 0.f;
        const float Dd = p.D ? __ldg(p.D + lc.d) : 0.f;
        float y[TSEG], dlu[TSEG];
#pragma unroll
        for (int i = 0; i < TSEG; ++i) {
            if (lc.active) {
                float x = dl[i] + bias;
                float e;
                dl[i] = p.softplus ? softplus20(x, e) : x;
                y[i] = Dd * u[i];
                dlu[i] = dl[i] * u[i];
            } else {
                dl[i] = 0.f, y[i] = 0.f, dlu[i] = 0.f;
            }
        }
        if (cc == 0) {
            carry1 = 0.f;
            if (NS == 0 && lane < N) s_carry[lane] = 0.f, s_carry[lane + 32] = 0.f;
            if (NS == 0) __syncwarp();
        }
        for (int n = 0; n < N; ++n) {
            const float An2 = __ldg(p.A + static_cast<long long>(lc.d) * N + n) * kLog2e;
            float a[TSEG], b[TSEG], Cv[TSEG];
            if (lc.active) {
                load_seg<TSEG>(gB + lc.bc + static_cast<long long>(n) * p.L, b);
                load_seg<TSEG>(gC + lc.bc + static_cast<long long>(n) * p.L, Cv);
            } else {
#pragma unroll
                for (int i = 0; i < TSEG; ++i) b[i] = 0.f, Cv[i] = 0.f;
            }
            float P = 1.f, S = 0.f;
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                a[i] = ex2f(dl[i] * An2);
                if (i == 0 && lc.rowstart) a[0] = 0.f;
                b[i] = dlu[i] * b[i];
                S = fmaf(a[i], S, b[i]);
                P *= a[i];
            }
            warp_scan_affine(P, S, lane);
            float Pe = __shfl_up_sync(0xffffffffu, P, 1);
            float Se = __shfl_up_sync(0xffffffffu, S, 1);
            if (lane == 0) Pe = 1.f, Se = 0.f;
            const float cin = NS == 1 ? carry1 : s_carry[n];
            float h = fmaf(Pe, cin, Se);
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                h = fmaf(a[i], h, b[i]);
                y[i] = fmaf(Cv[i], h, y[i]);
            }
            const float cout = __shfl_sync(0xffffffffu, h, 31);
            if (NS == 1) {
                carry1 = cout;
            } else {
                if (lane == 0) s_carry[n] = cout;
            }
            if (p.ckpt && lane == 0) p.ckpt[(unit * p.cpu + cc) * N + n] = cout;
        }
        // ---- store y
        if constexpr (TMA) {
            const int ob = static_cast<int>(cq & 1);
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            if (lc.active) store_seg<TSEG>(s_out + ob * CH + lane * TSEG, y);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (nt > 0) bulk_s2g(gout + tok0, s_out + ob * CH, nt * sizeof(out_t));
                bulk_commit();
            }
        } else {
            if (lc.active) store_seg<TSEG>(gout + tok0 + lane * TSEG, y);
            __syncwarp();
        }
        if (++cc == p.cpu) cc = 0, ++cj;
    }
    if constexpr (TMA) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// Regular plans (the shapes that matter: L % CH == 0, or CH = rpc * L with rpc in {2, 4}).
// Everything that does not depend on the chunk is hoisted: a lane's row, channel, A / D / bias
// and B/C base live in registers for the whole row; the next row's parameters and the next
// chunk's B/C segment are prefetched while the current chunk computes; stage / phase / pointer
// bookkeeping is incremental 32-bit arithmetic.  LPR = lanes per row (32, 16 or 8): with
// LPR < 32 a warp runs 32/LPR short rows side by side and the scan shuffles stay inside each
// LPR-lane group.
// ---------------------------------------------------------------------------------------
template <int TSEG, int NS, typename in_t, typename out_t, int LPR>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 3) sscan_fwd_regular_kernel(const SScanArgs p) {
    using SM = FwdSmem<TSEG, in_t, out_t>;
    constexpr int CH = 32 * TSEG;
    constexpr int RPC = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane / LPR;      // row inside the chunk
    const int sl = lane & (LPR - 1); // lane inside the row
    const int wglobal = blockIdx.x * kWarpsPerCta + warp;
    const int wtotal = gridDim.x * kWarpsPerCta;
    const int N = NS ? NS : p.N;
    const int cpu = p.cpu;
    const int L = p.L;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* wb = smem_raw + warp * SM::WARP_BYTES;
    in_t* s_in = reinterpret_cast<in_t*>(wb);
    out_t* s_out = reinterpret_cast<out_t*>(wb + SM::OFF_OUT);
    uint64_t* full = reinterpret_cast<uint64_t*>(wb + SM::OFF_BAR);
    float* s_carry = reinterpret_cast<float*>(wb + SM::OFF_CARRY);

    const in_t* __restrict__ gu = static_cast<const in_t*>(p.u);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.delta);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bm);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cm);
    out_t* __restrict__ gout = static_cast<out_t*>(p.out);

    const int units = static_cast<int>(p.units);
    if (wglobal >= units) return;
    const int my_units = (units - wglobal + wtotal - 1) / wtotal;
    const int T = my_units * cpu;
    const long long unit_tok = static_cast<long long>(RPC) * L;          // tokens per unit
    const long long stride_tok = unit_tok * wtotal;                      // to this warp's next unit

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncwarp();

    // ---- producer (lane 0): TMA bulk loads, kFwdStages-1 tiles ahead
    int pq = 0, pc = 0, pst = 0;
    long long ptok = wglobal * unit_tok;
    auto issue = [&]() {
        if (pq < T) {
            if (lane == 0) {
                const long long t0 = ptok + pc * CH;
                mbar_expect_tx(&full[pst], 2 * CH * sizeof(in_t));
                bulk_g2s(s_in + (pst * 2 + 0) * CH, gu + t0, CH * sizeof(in_t), &full[pst]);
                bulk_g2s(s_in + (pst * 2 + 1) * CH, gd + t0, CH * sizeof(in_t), &full[pst]);
            }
            ++pq;
            if (++pc == cpu) pc = 0, ptok += stride_tok;
            if (++pst == kFwdStages) pst = 0;
        }
    };
#pragma unroll
    for (int s = 0; s < kFwdStages - 1; ++s) issue();

    // ---- per-row context (registers)
    struct Ctx {
        float A2, bias, Dd;
        int d;
        long long bc;  // element offset of (b, g, n = 0, t = sl*TSEG) in B / C
    };
    auto make_ctx = [&](int unit) {
        Ctx c;
        const int row = unit * RPC + sub;
        const int b = static_cast<int>(p.divDim.div(static_cast<uint32_t>(row)));
        c.d = row - b * p.dim;
        const int g = static_cast<int>(p.divDpg.div(static_cast<uint32_t>(c.d)));
        c.bc = (static_cast<long long>(b) * p.groups + g) * N * static_cast<long long>(L) + sl * TSEG;
        c.bias = p.bias ? __ldg(p.bias + c.d) : 0.f;
        c.Dd = p.D ? __ldg(p.D + c.d) : 0.f;
        c.A2 = (NS == 1) ? __ldg(p.A + c.d) * kLog2e : 0.f;
        return c;
    };

    int unit = wglobal;
    Ctx cur = make_ctx(unit), nxt = cur;
    float Bn[TSEG], Cn[TSEG];  // prefetched B / C segment of the next (tile, state)
    load_seg<TSEG>(gB + cur.bc, Bn);
    load_seg<TSEG>(gC + cur.bc, Cn);

    float carry1 = 0.f;
    int c = 0, cst = 0;
    uint32_t cph = 0;
    long long tok = wglobal * unit_tok;  // first token of the current unit
    for (int q = 0; q < T; ++q) {
        issue();
        const bool last_c = (c + 1 == cpu);
        const bool more = (q + 1 < T);
        if (last_c && more) nxt = make_ctx(unit + wtotal);

        mbar_wait(&full[cst], cph);
        float u[TSEG], dl[TSEG];
        load_seg<TSEG>(s_in + (cst * 2 + 0) * CH + lane * TSEG, u);
        load_seg<TSEG>(s_in + (cst * 2 + 1) * CH + lane * TSEG, dl);
        float y[TSEG], dlu[TSEG];
#pragma unroll
        for (int i = 0; i < TSEG; ++i) {
            const float x = dl[i] + cur.bias;
            float e;
            dl[i] = p.softplus ? softplus20(x, e) : x;
            y[i] = cur.Dd * u[i];
            dlu[i] = dl[i] * u[i];
        }
        if (c == 0) {
            carry1 = 0.f;
            if (NS == 0) {
                if (lane < N) s_carry[lane] = 0.f;
                if (lane + 32 < N) s_carry[lane + 32] = 0.f;
                __syncwarp();
            }
        }
        for (int n = 0; n < N; ++n) {
            float a[TSEG], b[TSEG], Cv[TSEG];
#pragma unroll
            for (int i = 0; i < TSEG; ++i) b[i] = Bn[i], Cv[i] = Cn[i];
            // prefetch the B / C segment of the next state, or of the next tile's state 0
            {
                long long off = -1;
                if (n + 1 < N)
                    off = cur.bc + static_cast<long long>(n + 1) * L + c * CH;
                else if (more)
                    off = last_c ? nxt.bc : cur.bc + (c + 1) * CH;
                if (off >= 0) {
                    load_seg<TSEG>(gB + off, Bn);
                    load_seg<TSEG>(gC + off, Cn);
                }
            }
            const float An2 = (NS == 1) ? cur.A2 : __ldg(p.A + static_cast<long long>(cur.d) * N + n) * kLog2e;
            float P = 1.f, S = 0.f;
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                a[i] = ex2f(dl[i] * An2);
                b[i] = dlu[i] * b[i];
                S = fmaf(a[i], S, b[i]);
                P *= a[i];
            }
            group_scan_affine<LPR>(P, S, sl);
            float Pe = __shfl_up_sync(0xffffffffu, P, 1, LPR);
            float Se = __shfl_up_sync(0xffffffffu, S, 1, LPR);
            if (sl == 0) Pe = 1.f, Se = 0.f;
            const float cin = (NS == 1) ? carry1 : s_carry[n];
            float h = fmaf(Pe, cin, Se);
#pragma unroll
            for (int i = 0; i < TSEG; ++i) {
                h = fmaf(a[i], h, b[i]);
                y[i] = fmaf(Cv[i], h, y[i]);
            }
            if (cpu > 1) {  // RPC == 1 here: one row per warp, warp-uniform carry
                const float cout = __shfl_sync(0xffffffffu, h, 31);
                if (NS == 1) {
                    carry1 = cout;
                } else if (lane == 0) {
                    s_carry[n] = cout;
                }
                if (p.ckpt && lane == 0) p.ckpt[(static_cast<long long>(unit) * cpu + c) * N + n] = cout;
            }
        }
        // ---- store y through shared memory with a bulk copy
        {
            const int ob = q & 1;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            store_seg<TSEG>(s_out + ob * CH + lane * TSEG, y);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(gout + tok + c * CH, s_out + ob * CH, CH * sizeof(out_t));
                bulk_commit();
            }
        }
        if (++cst == kFwdStages) cst = 0, cph ^= 1;
        if (last_c) {
            c = 0, unit += wtotal, tok += stride_tok, cur = nxt;
        } else {
            ++c;
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
}

template <int TSEG, int NS, typename in_t, typename out_t, int LPR>
cudaError_t launch_regular(const SScanArgs& a, cudaStream_t st) {
    using SM = FwdSmem<TSEG, in_t, out_t>;
    auto kern = sscan_fwd_regular_kernel<TSEG, NS, in_t, out_t, LPR>;
    const int smem = kWarpsPerCta * SM::WARP_BYTES;
    if (cudaError_t e = ensure_dyn_smem(kern, smem); e != cudaSuccess) return e;
    const int sms = device_sms();
    static int occ = 0;  // a property of the kernel image, identical on every (sm_100) device of the process
    if (!occ) {
        int o = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kWarpsPerCta * 32, smem);
        occ = o < 1 ? 1 : o;
    }
    // one resident wave of CTAs; each warp strides over its share of the units
    long long want = (a.units + kWarpsPerCta - 1) / kWarpsPerCta;
    long long cap = static_cast<long long>(sms) * occ;
    int grid = static_cast<int>(want < cap ? want : cap);
    kern<<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    note_launch();
    return cudaGetLastError();
}

template <int TSEG, typename in_t, typename out_t, int LPR>
cudaError_t launch_regular_ns(const SScanArgs& a, cudaStream_t st) {
    if (a.N == 1) return launch_regular<TSEG, 1, in_t, out_t, LPR>(a, st);
    return launch_regular<TSEG, 0, in_t, out_t, LPR>(a, st);
}

template <int TSEG, int NS, typename in_t, typename out_t, bool TMA>
cudaError_t launch_one(const SScanArgs& a, cudaStream_t st) {
    using SM = FwdSmem<TSEG, in_t, out_t>;
    auto kern = sscan_fwd_kernel<TSEG, NS, in_t, out_t, TMA>;
    const int smem = TMA ? kWarpsPerCta * SM::WARP_BYTES : kWarpsPerCta * kMaxState * 4;
    if (cudaError_t e = ensure_dyn_smem(kern, smem); e != cudaSuccess) return e;
    const int sms = device_sms();
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem);
    if (occ < 1) occ = 1;
    long long want = (a.units + kWarpsPerCta - 1) / kWarpsPerCta;
    long long cap = static_cast<long long>(sms) * occ * 2;
    int grid = static_cast<int>(want < cap ? want : cap);
    if (grid < 1) grid = 1;
    kern<<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    note_launch();
    return cudaGetLastError();
}

template <int TSEG, bool TMA, typename in_t, typename out_t>
cudaError_t launch_ns(const SScanArgs& a, cudaStream_t st) {
    if (a.N == 1) return launch_one<TSEG, 1, in_t, out_t, TMA>(a, st);
    return launch_one<TSEG, 0, in_t, out_t, TMA>(a, st);
}

template <typename in_t, typename out_t>
cudaError_t launch_t(const SScanArgs& a, const SScanPlan& p, cudaStream_t st) {
    if (p.regular && a.units < (1ll << 30)) {
        if (p.rpc == 1) {
            if (p.tseg == 12) return launch_regular_ns<12, in_t, out_t, 32>(a, st);
            if (p.tseg == 9) return launch_regular_ns<9, in_t, out_t, 32>(a, st);
            return launch_regular_ns<4, in_t, out_t, 32>(a, st);
        }
        if (p.tseg == 9 && p.rpc == 2) return launch_regular_ns<9, in_t, out_t, 16>(a, st);
        if (p.tseg == 9 && p.rpc == 4) return launch_regular_ns<9, in_t, out_t, 8>(a, st);
        if (p.tseg == 4 && p.rpc == 2) return launch_regular_ns<4, in_t, out_t, 16>(a, st);
        if (p.tseg == 4 && p.rpc == 4) return launch_regular_ns<4, in_t, out_t, 8>(a, st);
    }
    if (!p.tma) return launch_ns<1, false, in_t, out_t>(a, st);
    switch (p.tseg) {
        case 12: return launch_ns<12, true, in_t, out_t>(a, st);
        case 9: return launch_ns<9, true, in_t, out_t>(a, st);
        default: return launch_ns<4, true, in_t, out_t>(a, st);
    }
}

}  // namespace

cudaError_t launch_sscan_fwd(const SScanArgs& a, const SScanPlan& p, int in_dtype, int out_dtype,
                             cudaStream_t st) {
    if (in_dtype == TRAMBA_F32 && out_dtype == TRAMBA_F32) return launch_t<float, float>(a, p, st);
    if (in_dtype == TRAMBA_BF16 && out_dtype == TRAMBA_F32) return launch_t<__nv_bfloat16, float>(a, p, st);
    if (in_dtype == TRAMBA_BF16 && out_dtype == TRAMBA_BF16)
        return launch_t<__nv_bfloat16, __nv_bfloat16>(a, p, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
