// Fused SS2D core, backward WITH the cross-merge of du folded in (sm_100a).  See ss2d_fused.cuh for the data model.
//
// Replaces CrossMerge*.backward -> SelectiveScanOflex.backward -> CrossScan*.backward of
// Models/SS2D/csms6s.py:45-55, 914-923, 25-31 (and the _Dilation / _Window / _Line variants) for N = 1 in ONE kernel:
//     in : x, dy (B, D, L) IMAGE order; delta (B, K, D, L), B, C (B, K, L) scan order
//     out: dx (B, D, L) fp32 IMAGE order (accumulated over the K directions with TMA bulk reductions at L2),
//          ddelta (B, K, D, L) scan order, dB / dC (B, K, L), dA / dD / ddelta_bias (K*D) (fp32 reds over b).
// Neither xs, nor the K-fold copy of dy, nor du (the K-fold gradient w.r.t. xs) is ever written to HBM.
//
// How it is built (the first-generation kernel -- cp.async-staged skewed planes, five table-driven gathers per
// token, du written to HBM in scan order and folded by a second kernel -- was measured against it and removed,
// profiles/r02_fused_ab_bf16.jsonl):
//   * every operand plane moves with 1-D TMA bulk copies (no per-thread cp.async staging, no skewed planes);
//   * geometry is a template parameter (TP threads per plane, L = 18 * TP): all index arithmetic folds;
//   * the scan order is applied ONCE per round by a shared -> shared permutation pass that follows a host-made,
//     bank-conflict-free schedule (tramba_order_sched: an edge colouring of pixel banks x token banks), after which
//     the three recurrence passes read token-ordered chunks with conflict-free vector loads -- instead of five
//     table-driven gathers per token; directions whose table is the identity skip the pass altogether;
//   * du never leaves the SM in scan order: it is scattered back into an image-ordered fp32 plane by the inverse of
//     the same schedule and added to dx in global memory with cp.reduce.async.bulk (.add.f32).
//
// Work item = (b, k, block of R*PP channels); round = PP whole planes of direction k (NT * S = 9216 tokens).
// Per round:  wait x/dy -> [permute image -> token order] -> pass 1 (forward fold + reverse fold of the segment)
//   -> scans over the plane's threads -> pass 2a (dh, right to left) -> pass 2b (h, du, ddelta, dB/dC/dA/dD/dbias)
//   -> [scatter du token -> image order] -> bulk store ddelta, bulk reduce-add dx.
// Per item:   dB / dC accumulated in registers over the item's rounds -> folded over the PP planes of a round in shared
//   memory -> written token-ordered into an idle plane -> two bulk reduce-adds into the (b, k) rows (coalesced).
// PIPE (16-bit operands): two (x, dy) plane sets used alternately and two delta buffers, so the TMA loads of round
// s + 1 overlap round s.  !PIPE (fp32 operands, which need twice the staging): one set (+ the token-ordered set),
// loads issued when the round's stores have been read.
#include <type_traits>

#include "ss2d_fused.cuh"

namespace tramba {

namespace {

// TRAMBA_KO: knock-out timing experiments (WRONG results; never set in the product build): bit 0 skips the dB / dC
// vector reds, bit 1 stores dx instead of reduce-adding it, bit 2 skips both permutation passes
#ifndef TRAMBA_KO
#define TRAMBA_KO 0
#endif
#ifndef TRAMBA_ACC_MIN_PP
#define TRAMBA_ACC_MIN_PP 1
#endif
#ifndef TRAMBA_NKEEP
#define TRAMBA_NKEEP 18
#endif
constexpr int S = kFusedS;
constexpr int NT = kFusedNT;
constexpr int TOK = kFusedTok;
constexpr int NW = NT / 32;

constexpr size_t al16(size_t v) { return (v + 15) / 16 * 16; }

template <int ES, int TP, bool PIPE>
struct DxLayout {
    static constexpr int L = TP * S;
    static constexpr size_t XB = static_cast<size_t>(TOK) * ES, YB = static_cast<size_t>(TOK) * 4;
    static constexpr size_t x0 = 0, y0 = XB, x1 = XB + YB, y1 = 2 * XB + YB;  // plane sets 0 and 1
    static constexpr size_t d0 = 2 * (XB + YB);                               // delta in / ddelta out (in place)
    static constexpr size_t d1 = PIPE ? d0 + XB : d0;
    static constexpr size_t dx = d1 + XB;                                     // PIPE: image-ordered du staging
    static constexpr size_t b = PIPE ? dx + YB : dx;
    // fp32 operands at L = 9216: the (b, k) rows of B and C do not fit next to the planes; they are read from global
    // memory instead (36.9 KB rows shared by every channel of (b, k): L2 hits)
    static constexpr bool BCG = !PIPE && TP == NT;
    static constexpr size_t bcb = BCG ? 0 : al16(static_cast<size_t>(L) * ES);
    static constexpr size_t c = b + bcb;
    static constexpr bool SCH = TP < NT;                                      // room for the item's schedule (L <= 2304)
    static constexpr size_t sch = c + bcb;                                    // [2 * L] entries (orders with repeats: < 1.6 L)
    static constexpr size_t tot = sch + (SCH ? 2 * static_cast<size_t>(L) * 4 : 0);
    static constexpr size_t red = tot + 2 * 2 * NW * sizeof(float2);
    static constexpr size_t bar = al16(red + 3 * NW * sizeof(float));
    static constexpr size_t total = bar + 8 * sizeof(uint64_t);
};

template <typename in_t, bool SOFTPLUS, int TP, bool PIPE>
__global__ void __launch_bounds__(NT, 1) ss2d_bwd_kernel(const FusedBwdArgs p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int ES = sizeof(in_t);
    constexpr int L = TP * S, PP = NT / TP;
    using Lay = DxLayout<ES, TP, PIPE>;
    float* const sDXs = reinterpret_cast<float*>(smem + Lay::dx);  // PIPE only
    in_t* const sB = reinterpret_cast<in_t*>(smem + Lay::b);
    in_t* const sC = reinterpret_cast<in_t*>(smem + Lay::c);
    float2* const stot = reinterpret_cast<float2*>(smem + Lay::tot);  // [2 parity][2 fwd/rev][NW]
    float* const sred = reinterpret_cast<float*>(smem + Lay::red);    // [3][NW]
    uint64_t* const barXY = reinterpret_cast<uint64_t*>(smem + Lay::bar);  // [2] per plane set
    uint64_t* const barD = barXY + 2;                                      // [2] per delta buffer
    uint64_t* const barBC = barXY + 4;                                     // [1]
    auto sX = [&](int set) { return reinterpret_cast<in_t*>(smem + (set ? Lay::x1 : Lay::x0)); };
    auto sY = [&](int set) { return reinterpret_cast<float*>(smem + (set ? Lay::y1 : Lay::y0)); };
    auto sD = [&](int buf) { return reinterpret_cast<in_t*>(smem + (buf ? Lay::d1 : Lay::d0)); };

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = tid / TP;
    const int j = tid - q * TP;
    constexpr int width = TP < 32 ? TP : 32;
    const int sl = lane & (width - 1);
    constexpr int nwg = TP > 32 ? (TP >> 5) : 1;  // warps per plane
    const int wq = TP > 32 ? (j >> 5) : 0;
    const int gw0 = TP > 32 ? q * nwg : 0;

    const in_t* __restrict__ gx = static_cast<const in_t*>(p.x);
    const in_t* __restrict__ gd = static_cast<const in_t*>(p.dts);
    const in_t* __restrict__ gB = static_cast<const in_t*>(p.Bs);
    const in_t* __restrict__ gC = static_cast<const in_t*>(p.Cs);
    in_t* __restrict__ gdd = static_cast<in_t*>(p.ddts);

    const int K = p.K, D = p.D, R = p.R;
    const int RB = R * PP;  // channels per item
    const int dblocks = D / RB;
    const int items = p.B * K * dblocks;
    const int my_items = (items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int nrounds = my_items * R;
    if (nrounds <= 0) return;

    if (tid == 0) {
        mbar_init(&barXY[0], 1), mbar_init(&barXY[1], 1), mbar_init(&barD[0], 1), mbar_init(&barD[1], 1);
        mbar_init(&barBC[0], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncthreads();

    auto item_coord = [&](int it, int& b, int& k, int& d0) {
        const uint32_t item = blockIdx.x + static_cast<uint32_t>(it) * gridDim.x;
        const uint32_t bk = p.div_db.div(item);
        const uint32_t db = item - bk * static_cast<uint32_t>(dblocks);
        b = static_cast<int>(p.div_k.div(bk));
        k = static_cast<int>(bk) - b * K;
        d0 = static_cast<int>(db) * RB;
    };
    auto round_coord = [&](int s, int& b, int& k, int& d0) {  // global round s of this CTA -> its PP planes
        const int it = static_cast<int>(p.div_r.div(static_cast<uint32_t>(s)));
        item_coord(it, b, k, d0);
        d0 += (s - it * R) * PP;
    };
    // ---- producers (thread 0 only): coordinates of global round s of this CTA
    auto issue_xy = [&](int set, int s) {
        int b, k, d0;
        round_coord(s, b, k, d0);
        const long long pl = (static_cast<long long>(b) * D + d0) * L;
        mbar_expect_tx(&barXY[set], TOK * (ES + 4));
        bulk_g2s(sX(set), gx + pl, TOK * ES, &barXY[set]);
        bulk_g2s(sY(set), p.dy + pl, TOK * 4, &barXY[set]);
    };
    auto issue_d = [&](int buf, int s) {
        int b, k, d0;
        round_coord(s, b, k, d0);
        mbar_expect_tx(&barD[buf], TOK * ES);
        bulk_g2s(sD(buf), gd + ((static_cast<long long>(b) * K + k) * D + d0) * L, TOK * ES, &barD[buf]);
    };
    auto issue_bc = [&](int it) {
        int b, k, d0;
        item_coord(it, b, k, d0);
        constexpr uint32_t lb = L * ES;
        mbar_expect_tx(&barBC[0], 2 * lb);
        bulk_g2s(sB, gB + (static_cast<long long>(b) * K + k) * L, lb, &barBC[0]);
        bulk_g2s(sC, gC + (static_cast<long long>(b) * K + k) * L, lb, &barBC[0]);
    };
    constexpr bool BCG = Lay::BCG;
    if (tid == 0) {
        if (!BCG) issue_bc(0);
        issue_xy(0, 0);
        issue_d(0, 0);
    }

    uint32_t phXY0 = 0, phXY1 = 0, phD0 = 0, phD1 = 0;  // mbarrier phase parities (uniform)
    int ld = 0;                                         // plane set that receives / holds this round's x, dy
    int s = 0;
    for (int it = 0; it < my_items; ++it) {
        int b, k, d00;
        item_coord(it, b, k, d00);
        const int kind = (p.kinds >> (4 * k)) & 15;  // 0 identity, 1 reversed identity, 2 permutation, 3 with repeats
        const bool perm = kind >= 2 && !(TRAMBA_KO & 4);
        const bool dup = kind == 3;
        const bool dense = kind == 2 && p.sched_used[k] == L;  // a permutation: every schedule entry is a move
        // the direction's permutation schedule: staged in shared memory for the item where there is room (short
        // planes: no L2 latency inside the permutation passes, no registers to hold entries), else read from L2
        constexpr bool SCH = Lay::SCH;
        const uint2* gsched = reinterpret_cast<const uint2*>(p.sched + static_cast<size_t>(k) * p.sched_len);
        // (PIPE, accumulated dB / dC) the previous item's flush is still reading its scratch plane: it must have been
        // read before this item's first round loads or permutes into that plane set
        constexpr bool FLUSH_WAIT = PIPE && (PP > 1 || ES == 2);
        if (FLUSH_WAIT && it > 0 && tid == 0) bulk_wait_read<0>();
        if (FLUSH_WAIT && !Lay::SCH && it > 0) __syncthreads();
        if constexpr (SCH) {
            uint2* ss = reinterpret_cast<uint2*>(smem + Lay::sch);
            const int n2 = min(p.sched_used[k], 2 * L) / 2;
            if (kind >= 2)
                for (int e = tid; e < n2; e += NT) ss[e] = __ldg(gsched + e);
            __syncthreads();
        }
        const bool sched_in_smem = SCH && p.sched_used[k] <= 2 * L;
        const uint2* __restrict__ sched = (sched_in_smem ? reinterpret_cast<const uint2*>(smem + Lay::sch) : gsched) + j;
        const int sched_iters = p.sched_used[k] / (2 * TP);
        // (the item's B / C rows are awaited inside its first round, after the permutation pass that does not need them)
        // dB / dC: this thread's tokens of direction k, (b, k) row of the (B, K, L) outputs
        float* const gdB = p.dBs + (static_cast<size_t>(b) * K + k) * L + j * S;
        float* const gdC = p.dCs + (static_cast<size_t>(b) * K + k) * L + j * S;
        const bool odd = j & 1;
        // ACC (several planes per round, L <= 2304): the PP planes of a round -- and every CTA working on the same
        // (b, k) -- would all add into the same few KB of dB / dC, and same-address reds serialise at L2.  There the
        // contributions stay in registers over the item's rounds and are folded over the planes in shared memory
        // first (one red per token per ITEM).  PP == 1 (L = 9216): with 16-bit operands the accumulators are kept as
        // well (reds once per item instead of once per round: the per-round 16-byte reds are 73 KB per round and SM of
        // L2 atomic traffic -- more than the kernel's HBM traffic -- and cost 26 % of the kernel in a knock-out run;
        // the accumulators spill ~30 words but are 4-9 % faster on the std / helix sites and equal on the others,
        // profiles/r02_fused_bwd_knockout.txt).  fp32 operands keep the per-round vector reds (392 bytes of spills
        // otherwise).  -DTRAMBA_ACC_MIN_PP=4 (A/B builds): per-round reds for up to 4 planes per round.
        constexpr bool ACC = PP > TRAMBA_ACC_MIN_PP || (ES == 2 && TRAMBA_ACC_MIN_PP == 1);
        float accB[ACC ? S : 1], accC[ACC ? S : 1];
        if constexpr (ACC) {
#pragma unroll
            for (int m = 0; m < S; ++m) accB[m] = 0.f, accC[m] = 0.f;
        }
        const in_t* Bq = BCG ? gB + (static_cast<size_t>(b) * K + k) * L + j * S : sB + j * S;
        const in_t* Cq = BCG ? gC + (static_cast<size_t>(b) * K + k) * L + j * S : sC + j * S;

        // One round of the item.  REV (compile time): the direction walks the image backwards (token t = pixel
        // L-1-t): the thread owning tokens [jS, (j+1)S) reads / writes image chunk TP-1-j back to front, no
        // permutation pass.
        auto round = [&](auto rev_tag, const int r) __attribute__((always_inline)) {
            constexpr bool REV = decltype(rev_tag)::value;
            const bool has_next = s + 1 < nrounds;
            const int dbuf = PIPE ? (s & 1) : 0;
            // the row's constants first: their L2 latency hides behind the plane wait and the permutation pass
            const int d0 = d00 + r * PP;
            const int d = d0 + q;
            const int row = k * D + d;
            const float Ar = __ldg(p.A + row);
            const float Dd = p.Dp ? __ldg(p.Dp + row) : 0.f;
            const float bias = p.bias ? __ldg(p.bias + row) : 0.f;
            if (PIPE && !perm && has_next && tid == 0) issue_xy(ld ^ 1, s + 1);
            if (ld == 0) {
                mbar_wait(&barXY[0], phXY0);
                phXY0 ^= 1;
            } else {
                mbar_wait(&barXY[1], phXY1);
                phXY1 ^= 1;
            }
            int ds = ld;  // set the recurrence passes read
            if (perm) {
                // ---- image order -> token order, by the conflict-free schedule (entry: 2*pixel | 2*token << 16).
                // Batches of 6 moves: all loads, then all stores; the next batch's entries are fetched meanwhile.
                const in_t* xi = sX(ld) + static_cast<size_t>(q) * L;
                const float* yi = sY(ld) + static_cast<size_t>(q) * L;
                in_t* xt = sX(ld ^ 1) + static_cast<size_t>(q) * L;
                float* yt = sY(ld ^ 1) + static_cast<size_t>(q) * L;
                auto move = [&](auto dense_tag, const uint2 ea, const uint2 eb, const uint2 ec, int cnt)
                                __attribute__((always_inline)) {
                    constexpr bool DENSE = decltype(dense_tag)::value;
                    const uint32_t ee[6] = {ea.x, ea.y, eb.x, eb.y, ec.x, ec.y};
                    uint32_t xv[6];
                    float yv[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const uint32_t e = ee[i];
                        const uint32_t src2 = (DENSE ? e : (e == 0xFFFFFFFFu ? 0u : e)) & 0xFFFFu;
                        if (DENSE || (i >> 1) < cnt) {
                            if constexpr (ES == 2) xv[i] = lds_u16(xi, src2);
                            else xv[i] = __float_as_uint(lds_f32(xi, 2 * src2));
                            yv[i] = lds_f32(yi, 2 * src2);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const uint32_t e = ee[i];
                        const uint32_t dst2 = e >> 16;
                        if (DENSE || ((i >> 1) < cnt && e != 0xFFFFFFFFu)) {
                            if constexpr (ES == 2) sts_u16(xt, dst2, xv[i]);
                            else sts_f32(xt, 2 * dst2, __uint_as_float(xv[i]));
                            sts_f32(yt, 2 * dst2, yv[i]);
                        }
                    }
                };
                if (sched_in_smem) {
                    if (dense) walk_schedule<TP, true, true>(sched, sched_iters, move);
                    else walk_schedule<TP, false, true>(sched, sched_iters, move);
                } else {
                    if (dense) walk_schedule_dense<TP>(sched, move);
                    else walk_schedule<TP, false>(sched, sched_iters, move);
                }
                __syncthreads();  // token-ordered set complete; the image-ordered set is dead
                if (PIPE && has_next && tid == 0) issue_xy(ld, s + 1);
                ds = ld ^ 1;
            }
            const float A2 = Ar * kLog2e;
            if (dbuf == 0) {
                mbar_wait(&barD[0], phD0);
                phD0 ^= 1;
            } else {
                mbar_wait(&barD[1], phD1);
                phD1 ^= 1;
            }
            if (!BCG && r == 0) mbar_wait(&barBC[0], static_cast<uint32_t>(it & 1));
            const int jc = REV ? TP - 1 - j : j;  // chunk of the plane that holds this thread's tokens
            const in_t* xq = sX(ds) + static_cast<size_t>(q) * L + jc * S;
            float* yq = sY(ds) + static_cast<size_t>(q) * L + jc * S;
            in_t* dq = sD(dbuf) + static_cast<size_t>(q) * L + j * S;  // delta / ddelta: always token order

            // ---- pass 1: forward fold (P, Sf) and reverse fold (P, Sr) of the segment; g_t = C_t * dy_t stays in dh
            constexpr bool KEEPG = true;  // (recomputing g in pass 2a measured MORE spills, not fewer)
            // a_t = exp(delta_t A) kept for passes 2a / 2b for the first NKEEP tokens of the segment, recomputed
            // (one MUFU) for the rest: as many as fit WITHOUT spilling -- with ~221 KB of shared memory the L1 is a few
            // KB, so every spilled word costs an L2 round trip on the critical path (long-scoreboard stalls)
            constexpr int NKEEP = PP == 1 ? TRAMBA_NKEEP : 0;
            float sp[S], dh[S], av[NKEEP ? NKEEP : 1];
            float P = 1.f, Sf = 0.f, Sr = 0.f;
            {
                load_tokens<S>(dq, sp);
#pragma unroll
                for (int mp = 0; mp < S / 2; ++mp) {
                    float Bv[2], Cv[2], uv[2], yv[2];
                    load_pair(Bq, mp, Bv[0], Bv[1]);
                    load_pair(Cq, mp, Cv[0], Cv[1]);
                    chunk_pair<REV>(xq, mp, uv[0], uv[1]);
                    chunk_pair<REV>(yq, mp, yv[0], yv[1]);
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int m = 2 * mp + h2;
                        const float raw = sp[m] + bias;
                        float e;
                        sp[m] = !SOFTPLUS ? raw : (ES == 2 ? softplus20_fast(raw) : softplus20(raw, e));
                        const float a = ex2f(sp[m] * A2);
                        if (m < NKEEP) av[m] = a;
                        const float bb = sp[m] * Bv[h2] * uv[h2];
                        const float g = Cv[h2] * yv[h2];
                        if constexpr (KEEPG) dh[m] = g;
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
            if (PIPE && tid == 0 && s > 0) bulk_wait_read<0>();  // previous round's ddelta / du staging have left
            __syncthreads();
            if (PIPE && has_next && tid == 0) issue_d((s + 1) & 1, s + 1);
            float* const dxp = PIPE ? sDXs : sY(ld);  // image-ordered du of this round (!PIPE: the dead dy planes)
            if (dup) {  // directions that skip pixels / hit pixels repeatedly: start from zero, add atomically
                float4* z = reinterpret_cast<float4*>(dxp);
                for (int e = tid; e < TOK / 4; e += NT) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
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
            if constexpr (KEEPG) {
                float E = Ein;
#pragma unroll
                for (int m = S - 1; m >= 0; --m) {
                    const float a = m < NKEEP ? av[m] : ex2f(sp[m] * A2);
                    dh[m] += E;
                    E = a * dh[m];
                }
            } else {
                float E = Ein;
#pragma unroll
                for (int mp = S / 2 - 1; mp >= 0; --mp) {
                    float Cv[2], yv[2];
                    load_pair(Cq, mp, Cv[0], Cv[1]);
                    chunk_pair<REV>(yq, mp, yv[0], yv[1]);
#pragma unroll
                    for (int h2 = 1; h2 >= 0; --h2) {
                        const int m = 2 * mp + h2;
                        const float a = ex2f(sp[m] * A2);
                        dh[m] = fmaf(Cv[h2], yv[h2], E);
                        E = a * dh[m];
                    }
                }
            }
            // ---- pass 2b: left-to-right, h_t; du, ddelta; accumulate dB, dC, dA, dD, dbias
            // dB_t += dh_t*delta_t*u_t and dC_t += dy_t*h_t: accumulated in registers (ACC, see above) or -- fp32 operands
            // at L = 9216 only -- sent per round as 16-byte vector reds (fp32 adds at L2 over the channels of (b, k)):
            // no accumulator registers live across the rounds of an item.  A chunk of
            // 18 floats starts 16-byte aligned for even j only, so a lane pair shares its 36 floats as 9 quads: the
            // even lane sends quads 0-3 from its own values and quad 4 = its last two + the odd lane's first two
            // (one shuffle pair after the loop), the odd lane quads 5-8 = its values 2..17.
            float h = hin, sA = 0.f, sDs = 0.f, sb = 0.f;
            float cB[S], cC[S];
#pragma unroll
            for (int mp = 0; mp < S / 2; ++mp) {
                float Bv[2], uv[2], yv[2];
                load_pair(Bq, mp, Bv[0], Bv[1]);
                chunk_pair<REV>(xq, mp, uv[0], uv[1]);
                chunk_pair<REV>(yq, mp, yv[0], yv[1]);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int m = 2 * mp + h2;
                    const float u = uv[h2], dyv = yv[h2], Bm = Bv[h2];
                    const float a = m < NKEEP ? av[m] : ex2f(sp[m] * A2);
                    const float w = a * h;  // a_t * h_{t-1}
                    const float bu = Bm * u;
                    h = fmaf(sp[m], bu, w);  // h_t
                    const float dhs = dh[m] * sp[m];
                    const float du = fmaf(Dd, dyv, dhs * Bm);
                    float dd = dh[m] * fmaf(Ar, w, bu);
                    if (SOFTPLUS) {
                        // sigmoid(raw) = 1 - exp(-softplus(raw)) = -expm1(-sp); series below 1/8 keeps fp32 relative
                        // accuracy where 1 - exp(.) would cancel (moot for 16-bit operands: <= 6e-8 absolute)
                        const float t = sp[m];
                        if (ES == 2) {
                            dd *= 1.0f - ex2f(-t * kLog2e);
                        } else {
                            const float ser = t * (1.0f - t * (0.5f - t * (0.16666667f - t * 0.041666668f)));
                            dd *= (t < 0.125f) ? ser : (1.0f - ex2f(-t * kLog2e));
                        }
                    }
                    if constexpr (ACC) {
                        accB[m] = fmaf(dhs, u, accB[m]);
                        accC[m] = fmaf(dyv, h, accC[m]);
                    } else {
                        cB[m] = dhs * u;
                        cC[m] = dyv * h;
                    }
                    sA = fmaf(dhs, w, sA);
                    sDs = fmaf(dyv, u, sDs);
                    sb += dd;
                    sp[m] = dd;  // reuse: sp <- ddelta, dh <- du
                    dh[m] = du;
                }
                // quads that are complete after this pair of tokens (m = 2*mp + 1)
                if (!ACC && mp % 2 == 1 && mp < S / 2 - 1) {  // m = 3, 7, 11, 15: even lanes, values m-3 .. m
                    if (!odd && !((TRAMBA_KO & 1) && p.softplus != 77)) {
                        red_add_v4(gdB + 2 * mp - 2, cB[2 * mp - 2], cB[2 * mp - 1], cB[2 * mp], cB[2 * mp + 1]);
                        red_add_v4(gdC + 2 * mp - 2, cC[2 * mp - 2], cC[2 * mp - 1], cC[2 * mp], cC[2 * mp + 1]);
                    }
                }
                if (!ACC && mp % 2 == 0 && mp >= 2) {  // m = 5, 9, 13, 17: odd lanes, values m-3 .. m
                    if (odd && !((TRAMBA_KO & 1) && p.softplus != 77)) {
                        red_add_v4(gdB + 2 * mp - 2, cB[2 * mp - 2], cB[2 * mp - 1], cB[2 * mp], cB[2 * mp + 1]);
                        red_add_v4(gdC + 2 * mp - 2, cC[2 * mp - 2], cC[2 * mp - 1], cC[2 * mp], cC[2 * mp + 1]);
                    }
                }
            }
            if constexpr (!ACC) {
                const float b0 = __shfl_down_sync(0xffffffffu, cB[0], 1, width);
                const float b1 = __shfl_down_sync(0xffffffffu, cB[1], 1, width);
                const float c0 = __shfl_down_sync(0xffffffffu, cC[0], 1, width);
                const float c1 = __shfl_down_sync(0xffffffffu, cC[1], 1, width);
                if (!odd && !((TRAMBA_KO & 1) && p.softplus != 77)) {
                    red_add_v4(gdB + S - 2, cB[S - 2], cB[S - 1], b0, b1);
                    red_add_v4(gdC + S - 2, cC[S - 2], cC[S - 1], c0, c1);
                }
            }
            store_tokens<S>(dq, sp);  // ddelta over delta, in place (this thread's own chunk)
            // du: contiguous orders -> straight into the image-ordered plane (!PIPE: over this thread's own dy chunk,
            // all of which has been read); otherwise token order over dy_tok, scattered below
            if constexpr (REV) {
                store_tokens_rev<S>(dxp + static_cast<size_t>(q) * L + jc * S, dh);
            } else {
                store_tokens<S>((perm ? yq : dxp + static_cast<size_t>(q) * L + j * S), dh);
            }
            // ---- per-row sums over L: warp reduce, then over the plane's warps
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                if (off < width) {
                    sA += __shfl_xor_sync(0xffffffffu, sA, off, width);
                    sDs += __shfl_xor_sync(0xffffffffu, sDs, off, width);
                    sb += __shfl_xor_sync(0xffffffffu, sb, off, width);
                }
            }
            if (TP > 32 && lane == 0) sred[warp] = sA, sred[NW + warp] = sDs, sred[2 * NW + warp] = sb;
            if (perm) {
                __syncthreads();  // du in token order complete (and the zero fill of a dup direction)
                // ---- token order -> image order: inverse walk of the same schedule
                const float* ut = sY(ds) + static_cast<size_t>(q) * L;
                float* di = dxp + static_cast<size_t>(q) * L;
                auto move = [&](auto dense_tag, const uint2 ea, const uint2 eb, const uint2 ec, int cnt)
                                __attribute__((always_inline)) {
                    constexpr bool DENSE = decltype(dense_tag)::value;
                    const uint32_t ee[6] = {ea.x, ea.y, eb.x, eb.y, ec.x, ec.y};
                    float v[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const uint32_t e = ee[i];
                        const uint32_t dst2 = (DENSE ? e : (e == 0xFFFFFFFFu ? 0u : e)) >> 16;
                        if (DENSE || (i >> 1) < cnt) v[i] = lds_f32(ut, 2 * dst2);
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const uint32_t e = ee[i];
                        const uint32_t src2 = e & 0xFFFFu;
                        if (DENSE || ((i >> 1) < cnt && e != 0xFFFFFFFFu)) sts_f32(di, 2 * src2, v[i]);
                    }
                };
                const uint2* sc = sched;
                if (dup) {
                    // repeated pixels: one move at a time, added with shared-memory atomics (a CAS loop each; the
                    // schedule still keeps the 32 lanes of a step on distinct banks)
                    for (int n = 0; n < sched_iters; ++n, sc += TP) {
                        const uint2 e2 = sched_in_smem ? *sc : __ldg(sc);
                        if (e2.x != 0xFFFFFFFFu) atomicAdd(di + ((e2.x & 0xFFFFu) >> 1), ut[e2.x >> 17]);
                        if (e2.y != 0xFFFFFFFFu) atomicAdd(di + ((e2.y & 0xFFFFu) >> 1), ut[e2.y >> 17]);
                    }
                } else {
                    if (sched_in_smem) {
                        if (dense) walk_schedule<TP, true, true>(sc, sched_iters, move);
                        else walk_schedule<TP, false, true>(sc, sched_iters, move);
                    } else {
                        if (dense) walk_schedule_dense<TP>(sc, move);
                        else walk_schedule<TP, false>(sc, sched_iters, move);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(gdd + ((static_cast<long long>(b) * K + k) * D + d0) * L, sD(dbuf), TOK * ES);
                if (TRAMBA_KO & 2) bulk_s2g(p.dx + (static_cast<long long>(b) * D + d0) * L, dxp, TOK * 4);
                else bulk_reduce_add_f32_s2g(p.dx + (static_cast<long long>(b) * D + d0) * L, dxp, TOK * 4);
                bulk_commit();
                if (!PIPE) {
                    bulk_wait_read<0>();
                    if (has_next) {
                        issue_xy(ld, s + 1);
                        issue_d(0, s + 1);
                    }
                }
            }
            if (TP > 32 && j == 0) {
                float tA = 0.f, tD = 0.f, tb = 0.f;
                for (int w = 0; w < nwg; ++w) tA += sred[gw0 + w], tD += sred[NW + gw0 + w], tb += sred[2 * NW + gw0 + w];
                sA = tA, sDs = tD, sb = tb;
            }
            if (j == 0) {  // per-row sums over (b, t): B contributions per row, added at L2 (buffers zeroed by the host)
                atomicAdd(p.dA + row, sA);
                if (p.dD) atomicAdd(p.dD + row, sDs);
                if (p.dbias) atomicAdd(p.dbias + row, sb);
            }
            if (PIPE && !perm) ld ^= 1;
            ++s;
        };
        for (int r = 0; r < R; ++r) {
            if (kind == 1) round(std::true_type{}, r);
            else round(std::false_type{}, r);
        }
        if constexpr (ACC) if (!((TRAMBA_KO & 8) && p.softplus != 77)) {
            // ---- item flush.  Several planes per round: the PP planes' partial sums are folded in shared memory in two
            // steps -- the upper half of the planes parks its 2*S values per thread (token-ordered rows of 2L floats
            // per plane), the lower half adds them and parks the sums, then ALL threads add up the PP/2 rows column by
            // column (consecutive lanes = consecutive tokens) -- 5 barriers instead of the 2*log2(PP)+2 of a binary
            // tree, and every thread busy in the last step.  The sums (for one plane per round: every thread's
            // accumulators) then leave COALESCED, token-ordered from the scratch, by TMA bulk reductions into the
            // (b, k) rows: reds straight from the registers -- lane stride 18 floats, one 32-byte sector per lane --
            // cost ~20 % of the kernel in a knock-out run (profiles/r02_fused_bwd_knockout.txt).  No shared-memory
            // atomics (a CAS loop each).  Scratch = the dy plane of the set that holds no live data between items (the
            // next round's loads went / go to the other set, or -- PIPE -- are issued only after the item-start wait
            // below); (PP/2) rows of 2L floats = TOK floats.  It is free at once: no wait for the round's own stores.
            float* f = sY(ld ^ 1);
            float* const rowB = p.dBs + (static_cast<size_t>(b) * K + k) * L;
            float* const rowC = p.dCs + (static_cast<size_t>(b) * K + k) * L;
            if constexpr (PP > 1) {
                constexpr int HP = PP / 2;
                if (q >= HP) {
                    store_tokens<S>(f + (q - HP) * 2 * L + j * S, accB);
                    store_tokens<S>(f + (q - HP) * 2 * L + L + j * S, accC);
                }
                __syncthreads();
                if (q < HP) {
                    float tb[S], tc[S];
                    load_tokens<S>(f + q * 2 * L + j * S, tb);
                    load_tokens<S>(f + q * 2 * L + L + j * S, tc);
#pragma unroll
                    for (int m = 0; m < S; ++m) accB[m] += tb[m], accC[m] += tc[m];
                }
                __syncthreads();  // every parked row has been read
                if (q < HP) {
                    store_tokens<S>(f + q * 2 * L + j * S, accB);
                    store_tokens<S>(f + q * 2 * L + L + j * S, accC);
                }
                if constexpr (HP > 1) {
                    __syncthreads();
                    for (int o = tid; o < 2 * L; o += NT) {
                        float a = f[o];
#pragma unroll 4
                        for (int pl = 1; pl < HP; ++pl) a += f[pl * 2 * L + o];
                        f[o] = a;
                    }
                }
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_reduce_add_f32_s2g(rowB, f, L * 4);
                    bulk_reduce_add_f32_s2g(rowC, f + L, L * 4);
                    bulk_commit();
                    if (!PIPE) bulk_wait_read<0>();  // (PIPE: the next writer of the scratch waits for it mid-round)
                }
            } else {
                // one plane per round: dB fills the scratch plane (L floats); dC goes over the item's B / C rows, which
                // every thread has finished reading (2 * L 16-bit values = L floats) -- the next item's rows are loaded
                // only after this reduction has read them (below)
                static_assert(!ACC || PP > 1 || (ES == 2 && !BCG), "one-plane flush assumes staged 16-bit B / C rows");
                float* const fc = reinterpret_cast<float*>(sB);
                store_tokens<S>(f + j * S, accB);
                store_tokens<S>(fc + j * S, accC);
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_reduce_add_f32_s2g(rowB, f, L * 4);
                    bulk_reduce_add_f32_s2g(rowC, fc, L * 4);
                    bulk_commit();
                    bulk_wait_read<0>();  // tid 0 only: it issues the next item's B / C load right after
                }
            }
            __syncthreads();  // everybody is done with the scratch (and with sB / sC)
        }
        // the last round's closing barrier has every thread past its reads of sB / sC: the next item's B / C may land
        if (!BCG && tid == 0 && it + 1 < my_items) issue_bc(it + 1);
    }
    if (tid == 0) bulk_wait_all<0>();
}

// Rounds (channel groups) per item.  More rounds amortise the per-item dB / dC flush (>= 8 channels is the target);
// fewer rounds make more, smaller items, which matters when the item count is a small non-multiple of the SM count
// (512 items on 148 persistent CTAs leave a 14 % tail).  Largest R <= target whose last wave is >= 92 % full, else
// the best-filled one.
int pick_rounds(int B, int K, int D, int PP, int sms) {
    int best = 1;
    double best_eff = -1.0;
    for (int R = 8; R >= 1; R >>= 1) {
        if (R * PP > 8 && R > 1) continue;
        if (D % (R * PP) != 0) continue;
        const long long items = static_cast<long long>(B) * K * (D / (R * PP));
        const long long waves = (items + sms - 1) / sms;
        const double eff = static_cast<double>(items) / static_cast<double>(waves * sms);
        if (eff >= 0.92) return R;
        if (eff > best_eff) best_eff = eff, best = R;
    }
    return best;
}

template <typename in_t, bool SOFTPLUS, int TP>
cudaError_t launch_tp(const FusedBwdArgs& a, cudaStream_t st) {
    constexpr bool PIPE = sizeof(in_t) == 2;
    using Lay = DxLayout<sizeof(in_t), TP, PIPE>;
    if (Lay::total > 227 * 1024) return cudaErrorInvalidConfiguration;
    auto kern = ss2d_bwd_kernel<in_t, SOFTPLUS, TP, PIPE>;
    if (cudaError_t e = ensure_dyn_smem(kern, static_cast<int>(Lay::total)); e != cudaSuccess) return e;
    const int sms = device_sms();
    FusedBwdArgs b = a;
    constexpr int PP = NT / TP;
    b.R = pick_rounds(a.B, a.K, a.D, PP, sms);
    b.div_db.init(static_cast<uint32_t>(a.D / (b.R * PP)));
    b.div_k.init(static_cast<uint32_t>(a.K));
    b.div_r.init(static_cast<uint32_t>(b.R));
    const int items = a.B * a.K * (a.D / (b.R * PP));
    const int grid = items < sms ? items : sms;
    kern<<<grid, NT, Lay::total, st>>>(b);
    note_launch();
    return cudaGetLastError();
}

template <typename in_t, bool SOFTPLUS>
cudaError_t launch_t(const FusedBwdArgs& a, int TP, cudaStream_t st) {
    switch (TP) {
        case 512: return launch_tp<in_t, SOFTPLUS, 512>(a, st);
        case 128: return launch_tp<in_t, SOFTPLUS, 128>(a, st);
        case 32: return launch_tp<in_t, SOFTPLUS, 32>(a, st);
        case 8: return launch_tp<in_t, SOFTPLUS, 8>(a, st);
        default: return cudaErrorInvalidConfiguration;
    }
}

}  // namespace

bool ss2d_fused_bwd_fits(int L, int es) {
    const int TP = L / S;
    if (TP * S != L) return false;
    auto tot = [&](auto lay) { return decltype(lay)::total <= 227 * 1024; };
    switch (TP) {
        case 512: return es == 2 ? tot(DxLayout<2, 512, true>{}) : tot(DxLayout<4, 512, false>{});
        case 128: return es == 2 ? tot(DxLayout<2, 128, true>{}) : tot(DxLayout<4, 128, false>{});
        case 32: return es == 2 ? tot(DxLayout<2, 32, true>{}) : tot(DxLayout<4, 32, false>{});
        case 8: return es == 2 ? tot(DxLayout<2, 8, true>{}) : tot(DxLayout<4, 8, false>{});
        default: return false;
    }
}

cudaError_t launch_ss2d_fused_bwd(const FusedBwdArgs& a, int L, int in_dtype, cudaStream_t st) {
    const int TP = L / S;
    if (in_dtype == TRAMBA_F32) return a.softplus ? launch_t<float, true>(a, TP, st) : launch_t<float, false>(a, TP, st);
    if (in_dtype == TRAMBA_BF16)
        return a.softplus ? launch_t<__nv_bfloat16, true>(a, TP, st) : launch_t<__nv_bfloat16, false>(a, TP, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
