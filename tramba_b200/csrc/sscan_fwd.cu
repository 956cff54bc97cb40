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
                const int nt = min(CH, unit_tokens - pc * CH);
                const uint32_t bytes = nt * sizeof(in_t);
                const int st = static_cast<int>(pq % kFwdStages);
                mbar_expect_tx(&full[st], 2 * bytes);
                bulk_g2s(s_in + (st * 2 + 0) * CH, gu + tok0, bytes, &full[st]);
                bulk_g2s(s_in + (st * 2 + 1) * CH, gd + tok0, bytes, &full[st]);
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
        const int nt = min(CH, unit_tokens - cc * CH);
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
                bulk_s2g(gout + tok0, s_out + ob * CH, nt * sizeof(out_t));
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

template <int TSEG, int NS, typename in_t, typename out_t, bool TMA>
cudaError_t launch_one(const SScanArgs& a, cudaStream_t st) {
    using SM = FwdSmem<TSEG, in_t, out_t>;
    auto kern = sscan_fwd_kernel<TSEG, NS, in_t, out_t, TMA>;
    const int smem = TMA ? kWarpsPerCta * SM::WARP_BYTES : kWarpsPerCta * kMaxState * 4;
    static bool configured = false;  // per instantiation
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
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
