// Selective-scan (S6) forward / backward kernels for sm_100a -- shared definitions.
//
// Data model.  u/delta/out are (rows = batch*dim, L) row-major, i.e. ONE flat token stream of
// rows*L tokens.  A warp owns a contiguous piece of that stream (a "unit" of R whole rows) and
// walks it in chunks of CH = 32*TSEG tokens; lane l holds the TSEG consecutive tokens
// [l*TSEG, (l+1)*TSEG) of the chunk (a lane's tokens never straddle a row: L % TSEG == 0).
// The recurrence h_t = a_t h_{t-1} + b_t is evaluated as a three-level scan:
//   lane   : sequential fold over its TSEG tokens            -> (P, S) = (prod a, local h)
//   warp   : 5-step shuffle scan over the 32 (P, S) pairs     -> carry-in of every lane
//   chunk  : the warp's running carry h_end (one value / state), checkpointed per chunk so
//            that the backward can restart any chunk without storing states.
// A row start resets the recurrence by forcing a_0 := 0 on the row's first token, which makes
// chunks that cover several short rows (L < CH) correct without any special casing.
//
// Memory path (TMA mode).  Each warp runs its own producer/consumer ring: lane 0 issues 1-D
// TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) of the next chunks of u / delta
// (/ dout) into the warp's private shared-memory stages; all lanes wait on the stage's
// mbarrier, read their TSEG tokens with conflict-free vector LDS (lane stride TSEG words,
// TSEG/4 odd, or TSEG odd for scalar reads), and results leave through shared memory with a
// bulk store (cp.async.bulk.global.shared::cta).  No __syncthreads anywhere: warps are
// independent pipelines.  DIRECT mode (TSEG = 1) is the fully general path for shapes whose
// rows are not 16-byte multiples: plain coalesced scalar loads/stores, same scan structure.
#pragma once
#include "common.cuh"

namespace tramba {

struct SScanPlan {
    int tseg;     // tokens per lane per chunk
    int ch;       // 32 * tseg
    bool tma;     // TMA-staged (rows are 16-byte multiples) vs DIRECT scalar path
    bool uniform; // every chunk lies inside one row (L % ch == 0)
    bool regular; // uniform, or a chunk holds exactly rpc whole rows (rpc a power of two <= 4)
    int rpc;      // rows per chunk (1 when uniform)
    int R;        // rows per unit
    int cpu;      // chunks per unit
    long long units;
};

inline long long gcd_ll(long long a, long long b) {
    while (b) {
        long long t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// The plan depends only on shapes and element sizes, so forward, backward and the checkpoint
// size always agree.
inline SScanPlan make_sscan_plan(long long rows, int L, int es_in) {
    SScanPlan p;
    const bool al = (static_cast<long long>(L) * es_in) % 16 == 0;
    if (!al)
        p.tseg = 1;
#ifdef TRAMBA_SSCAN_TSEG_FORCE   // A/B builds only: a shorter per-lane segment (fewer registers, more resident warps)
    else if (TRAMBA_SSCAN_TSEG_FORCE == 4 && L % 128 == 0)
        p.tseg = 4;
    else if (TRAMBA_SSCAN_TSEG_FORCE == 9 && L % 288 == 0)
        p.tseg = 9;
#endif
    else if (L % 384 == 0)
        p.tseg = 12;
    else if (L % 288 == 0 || (288 % L == 0 && L % 9 == 0))
        p.tseg = 9;
    else
        p.tseg = 4;
    p.tma = al;
    p.ch = 32 * p.tseg;
    if (L % p.ch == 0) {
        p.uniform = true;
        p.R = 1;
        p.cpu = L / p.ch;
    } else {
        p.uniform = false;
        long long l = static_cast<long long>(L) / gcd_ll(L, p.ch) * p.ch;  // lcm
        long long r = l / L;
        if (r > 64) r = (L >= p.ch) ? 1 : (p.ch / L);
        if (r > rows) r = rows;
        p.R = static_cast<int>(r);
        p.cpu = static_cast<int>((r * L + p.ch - 1) / p.ch);
    }
    p.units = (rows + p.R - 1) / p.R;
    p.rpc = 1;
    p.regular = p.uniform;
    if (!p.uniform && p.tma && p.ch % L == 0) {
        const int rpc = p.ch / L;
        if ((rpc == 2 || rpc == 4) && rows % rpc == 0 && p.R == rpc && p.cpu == 1) {
            p.rpc = rpc;
            p.regular = true;
        }
    }
    return p;
}

struct SScanArgs {
    // forward
    const void* u;
    const void* delta;
    const void* Bm;
    const void* Cm;
    const float* A;
    const float* D;
    const float* bias;
    void* out;
    float* ckpt;
    // backward
    const void* dout;
    const float* ckpt_in;
    void* du;
    void* ddelta;
    float* dA;
    float* dB;
    float* dC;
    float* dD;
    float* dbias;
    // shapes
    int batch, dim, L, N, groups, dpg;
    int rows;  // batch * dim
    int R, cpu;
    long long units;
    int softplus;
    int uniform;
    int rpc;  // regular plans: whole rows per chunk (1, 2 or 4)
    int rw;   // backward: rows pre-reduced per warp for dB/dC (uniform plans only)
    FastDiv divL, divDim, divDpg;
};

cudaError_t launch_sscan_fwd(const SScanArgs& a, const SScanPlan& p, int in_dtype, int out_dtype,
                             cudaStream_t st);
cudaError_t launch_sscan_bwd(const SScanArgs& a, const SScanPlan& p, int in_dtype, int dout_dtype,
                             cudaStream_t st);

// ---------------------------------------------------------------------------------------
// Device pieces shared by forward and backward.
// ---------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 4;
constexpr int kMaxState = 64;

// Inclusive warp scan of the affine maps h -> P*h + S, composed left to right.
__device__ __forceinline__ void warp_scan_affine(float& P, float& S, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float Pp = __shfl_up_sync(0xffffffffu, P, off);
        float Sp = __shfl_up_sync(0xffffffffu, S, off);
        if (lane >= off) {
            S = fmaf(P, Sp, S);
            P = P * Pp;
        }
    }
}
// Same, composed right to left (lane 31 first): used by the reverse recurrence.
__device__ __forceinline__ void warp_scan_affine_rev(float& P, float& S, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float Pp = __shfl_down_sync(0xffffffffu, P, off);
        float Sp = __shfl_down_sync(0xffffffffu, S, off);
        if (lane + off < 32) {
            S = fmaf(P, Sp, S);
            P = P * Pp;
        }
    }
}
// Width-LPR variants (LPR lanes per row, several rows side by side in one warp).
template <int LPR>
__device__ __forceinline__ void group_scan_affine(float& P, float& S, int sl) {
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) {
        float Pp = __shfl_up_sync(0xffffffffu, P, off, LPR);
        float Sp = __shfl_up_sync(0xffffffffu, S, off, LPR);
        if (sl >= off) {
            S = fmaf(P, Sp, S);
            P = P * Pp;
        }
    }
}
template <int LPR>
__device__ __forceinline__ void group_scan_affine_rev(float& P, float& S, int sl) {
#pragma unroll
    for (int off = 1; off < LPR; off <<= 1) {
        float Pp = __shfl_down_sync(0xffffffffu, P, off, LPR);
        float Sp = __shfl_down_sync(0xffffffffu, S, off, LPR);
        if (sl + off < LPR) {
            S = fmaf(P, Sp, S);
            P = P * Pp;
        }
    }
}
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off, LPR);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// Segmented (by ascending key) inclusive sum over lanes; returns true on the last lane of
// each segment, whose `v` then holds the segment total.
__device__ __forceinline__ bool warp_seg_sum(float& v, int key, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float v2 = __shfl_up_sync(0xffffffffu, v, off);
        int k2 = __shfl_up_sync(0xffffffffu, key, off);
        if (lane >= off && k2 == key) v += v2;
    }
    int kn = __shfl_down_sync(0xffffffffu, key, 1);
    return lane == 31 || kn != key;
}

// Read CNT consecutive elements of a lane's segment.  Vector width follows what the lane
// stride guarantees: TSEG % 4 == 0 -> 16 B (fp32) / 8 B (16-bit); otherwise scalar.
template <int CNT, typename T>
__device__ __forceinline__ void load_seg(const T* __restrict__ p, float (&v)[CNT]) {
    if constexpr (CNT % 4 == 0) {
        load_vec<CNT>(p, v);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) v[i] = to_f32(p[i]);
    }
}
template <int CNT, typename T>
__device__ __forceinline__ void store_seg(T* __restrict__ p, const float (&v)[CNT]) {
    if constexpr (CNT % 4 == 0) {
        store_vec<CNT>(p, v);
    } else {
#pragma unroll
        for (int i = 0; i < CNT; ++i) p[i] = from_f32<T>(v[i]);
    }
}

// Per-lane coordinates of its segment inside the current chunk.
struct LaneCoord {
    bool active;   // segment lies inside the unit
    bool rowstart; // segment starts a row
    int row;       // global row (b*dim + d)
    int d;         // channel
    int t0;        // token offset inside the row
    long long bc;  // element offset of (b, g, n=0, t0) in B / C
};

__device__ __forceinline__ LaneCoord lane_coord(const SScanArgs& p, long long unit, int chunk, int lane,
                                                int tseg, int ch, int unit_tokens) {
    LaneCoord c;
    int tiu = chunk * ch + lane * tseg;
    c.active = tiu < unit_tokens;
    int riu = static_cast<int>(p.divL.div(static_cast<uint32_t>(tiu)));
    c.t0 = tiu - riu * p.L;
    c.rowstart = (c.t0 == 0);
    int row = static_cast<int>(unit) * p.R + riu;
    if (!c.active) row = p.rows - 1, c.t0 = 0;
    c.row = row;
    int b = static_cast<int>(p.divDim.div(static_cast<uint32_t>(row)));
    c.d = row - b * p.dim;
    int g = static_cast<int>(p.divDpg.div(static_cast<uint32_t>(c.d)));
    c.bc = (static_cast<long long>(b) * p.groups + g) * p.N * static_cast<long long>(p.L) + c.t0;
    return c;
}

}  // namespace tramba
