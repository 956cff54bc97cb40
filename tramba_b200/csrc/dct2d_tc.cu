// DCT2D high/low split on the tensor pipe (sm_100a): Y = Wy . X . Wx^T per plane, forward and
// adjoint, as two chained in-shared-memory GEMMs per plane evaluated with warp-level
// mma.sync.m16n8k8 TF32 tensor-core instructions and 3-term error compensation
//     a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,   x_hi = tf32(x), x_lo = x - x_hi
// which keeps fp32-grade accuracy (the reference multiplies in fp32: Models/DCT_2D.py:47-53,
// 71-77; single-pass TF32 would miss the 1e-3 tolerance over two chained contractions).
// Only the two quadrants the caller keeps are formed: low = Y[:H/2, :W/2], high = Y[H/2:, W/2:].
//
// A persistent CTA (12 warps) keeps both bases in shared memory for its whole life, streams
// planes through a double-buffered cp.async stage, and splits every stage's output tiles over
// its warps.  Row strides are padded so that every fragment load is bank-conflict-free
// (stride = 4 mod 32 words for row-major A / "n-major" B operands, 8 mod 32 for "k-major" B).
#include "common.cuh"

namespace tramba {

namespace {

constexpr int kWarps = 12;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One warp: C (16 x 8*NT, fp32 registers) += A (16 x K) . B (K x 8*NT), 3xTF32.
//   A element (m, k) at A[m*lda + k]                    (row-major, lda = 4 mod 32)
//   B element (k, n) at B[k*ldb_k + n*ldb_n]            (either orientation; see strides above)
template <int NT>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const float* __restrict__ A, int lda,
                                          const float* __restrict__ B, int ldb_k, int ldb_n, int K, int lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t ah[4], al[4];
        split_tf32(A[g * lda + k0 + t], ah[0], al[0]);
        split_tf32(A[(g + 8) * lda + k0 + t], ah[1], al[1]);
        split_tf32(A[g * lda + k0 + t + 4], ah[2], al[2]);
        split_tf32(A[(g + 8) * lda + k0 + t + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh[2], bl[2];
            const float* bp = B + (k0 + t) * ldb_k + (nt * 8 + g) * ldb_n;
            split_tf32(bp[0], bh[0], bl[0]);
            split_tf32(bp[4 * ldb_k], bh[1], bl[1]);
            mma_tf32(acc[nt], al, bh);
            mma_tf32(acc[nt], ah, bl);
            mma_tf32(acc[nt], ah, bh);
        }
    }
}

struct DctGeom {
    int H, W, Hh, Wh;
    int ldw4_y, ldw4_x;  // bases, stride = 4 mod 32
    int ldx;             // staged planes, stride = 8 mod 32
    int ldt;             // intermediate, stride = 4 mod 32 (fwd) / 8 mod 32 (bwd)
    int rows_t;          // rows of the intermediate incl. tile overhang
    int ppl;             // planes per stage
};
__host__ __device__ inline int pad_to(int n, int r) {  // smallest n' >= n with n' = r (mod 32)
    int p = n + ((r - n) % 32 + 32) % 32;
    return p;
}

// ------------------------------------------------------------------------------------ forward
// stage A: T = Wy . X        (A = Wy row-major [u][h];  B[k=h][n=w] = X[h][w]: ldb_k = ldx, ldb_n = 1)
// stage B: low  = T[:Hh]  . Wx[:Wh]^T ; high = T[Hh:] . Wx[Wh:]^T
//                            (A = T row-major [u][w];  B[k=w][n=v] = Wx[v][w]: ldb_k = 1, ldb_n = ldw4_x)
template <int NTA, int NTB>
__global__ void __launch_bounds__(kThreads, 1) dct2d_fwd_tc_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ wy,
                                                                   const float* __restrict__ wx,
                                                                   float* __restrict__ high, float* __restrict__ low,
                                                                   int planes, DctGeom gm) {
    extern __shared__ __align__(16) float sm[];
    const int H = gm.H, W = gm.W, Hh = gm.Hh, Wh = gm.Wh, PPL = gm.ppl;
    const int rows_w = ((H + 15) / 16) * 16 + 16;  // bases padded so that overhanging tiles stay in bounds
    float* Wys = sm;                                // [rows_w][ldw4_y]
    float* Wxs = Wys + rows_w * gm.ldw4_y;          // [W + 16][ldw4_x]
    float* Xs = Wxs + (W + 16) * gm.ldw4_x;         // [2][PPL][H][ldx]
    float* Ts = Xs + 2 * PPL * H * gm.ldx;          // [PPL][rows_t][ldt]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    for (int e = tid; e < rows_w * gm.ldw4_y; e += kThreads) {
        const int r = e / gm.ldw4_y, c = e - r * gm.ldw4_y;
        Wys[e] = (r < H && c < H) ? wy[r * H + c] : 0.f;
    }
    for (int e = tid; e < (W + 16) * gm.ldw4_x; e += kThreads) {
        const int r = e / gm.ldw4_x, c = e - r * gm.ldw4_x;
        Wxs[e] = (r < W && c < W) ? wx[r * W + c] : 0.f;
    }
    for (int e = tid; e < PPL * gm.rows_t * gm.ldt; e += kThreads) Ts[e] = 0.f;

    const int groups = (planes + PPL - 1) / PPL;
    const int chunks_row = W / 4;  // 16-byte chunks per image row
    auto stage = [&](int grp, int buf) {
        const int p0 = grp * PPL;
        const int np = min(PPL, planes - p0);
        const int total = np * H * chunks_row;
        float* dst = Xs + static_cast<size_t>(buf) * PPL * H * gm.ldx;
        const float* src = x + static_cast<size_t>(p0) * H * W;
        for (int e = tid; e < total; e += kThreads) {
            const int row = e / chunks_row, c = e - row * chunks_row;  // row = plane * H + h
            cp_async16(dst + row * gm.ldx + c * 4, src + static_cast<size_t>(row) * W + c * 4);
        }
        cp_async_commit();
    };
    int it = 0;
    if (static_cast<int>(blockIdx.x) < groups) stage(blockIdx.x, 0);
    __syncthreads();
    const int mtA = (H + 15) / 16, nsA = W / (8 * NTA);
    const int mtB = (Hh + 15) / 16, nsB = (Wh + 8 * NTB - 1) / (8 * NTB);
    for (int grp = blockIdx.x; grp < groups; grp += gridDim.x, ++it) {
        const int buf = it & 1;
        const int nxt = grp + gridDim.x;
        if (nxt < groups) stage(nxt, buf ^ 1);
        if (nxt < groups) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const int p0 = grp * PPL;
        const int np = min(PPL, planes - p0);
        const float* Xb = Xs + static_cast<size_t>(buf) * PPL * H * gm.ldx;
        // ---- stage A
        const int tasksA = np * mtA * nsA;
        for (int task = warp; task < tasksA; task += kWarps) {
            const int pl = task / (mtA * nsA), r = task - pl * (mtA * nsA);
            const int mt = r / nsA, ns = r - mt * nsA;
            float acc[NTA][4];
#pragma unroll
            for (int i = 0; i < NTA; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            warp_gemm<NTA>(acc, Wys + mt * 16 * gm.ldw4_y, gm.ldw4_y, Xb + pl * H * gm.ldx + ns * NTA * 8, gm.ldx, 1,
                           H, lane);
            float* Tp = Ts + (pl * gm.rows_t + mt * 16) * gm.ldt + ns * NTA * 8;
#pragma unroll
            for (int nt = 0; nt < NTA; ++nt) {
                *reinterpret_cast<float2*>(Tp + g * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
                *reinterpret_cast<float2*>(Tp + (g + 8) * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
        __syncthreads();
        // ---- stage B: the two quadrants
        const int tasksB = np * 2 * mtB * nsB;
        for (int task = warp; task < tasksB; task += kWarps) {
            int r = task;
            const int pl = r / (2 * mtB * nsB);
            r -= pl * (2 * mtB * nsB);
            const int qd = r / (mtB * nsB);  // 0 = low, 1 = high
            r -= qd * (mtB * nsB);
            const int mt = r / nsB, ns = r - mt * nsB;
            const int u0 = qd * Hh + mt * 16, v0 = qd * Wh + ns * NTB * 8;
            float acc[NTB][4];
#pragma unroll
            for (int i = 0; i < NTB; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            warp_gemm<NTB>(acc, Ts + (pl * gm.rows_t + u0) * gm.ldt, gm.ldt, Wxs + v0 * gm.ldw4_x, 1, gm.ldw4_x, W, lane);
            float* out = (qd ? high : low) + static_cast<size_t>(p0 + pl) * Hh * Wh;
#pragma unroll
            for (int nt = 0; nt < NTB; ++nt) {
                const int v = ns * NTB * 8 + nt * 8 + 2 * t;  // column inside the quadrant
                const int ua = mt * 16 + g, ub = ua + 8;
                if (v < Wh) {
                    if (ua < Hh) *reinterpret_cast<float2*>(out + ua * Wh + v) = make_float2(acc[nt][0], acc[nt][1]);
                    if (ub < Hh) *reinterpret_cast<float2*>(out + ub * Wh + v) = make_float2(acc[nt][2], acc[nt][3]);
                }
            }
        }
        __syncthreads();  // Ts and the stage buffer are free again
    }
}

// ------------------------------------------------------------------------------------ adjoint
// stage 1: S[u][w] = sum_v dY[u][v] Wx[v][w]   rows u < Hh use Gl / Wx[:Wh], rows u >= Hh use Gh / Wx[Wh:]
//          (A = G row-major [u][v], lda = ldg;  B[k=v][n=w] = Wx[v][w]: ldb_k = ldw8_x, ldb_n = 1)
// stage 2: gx[h][w] = sum_u Wy[u][h] S[u][w]
//          (A[m=h][k=u] = Wy^T staged transposed row-major [h][u];  B[k=u][n=w] = S[u][w]: ldb_k = lds, ldb_n = 1)
template <int NT1, int NT2>
__global__ void __launch_bounds__(kThreads, 1) dct2d_bwd_tc_kernel(const float* __restrict__ ghigh,
                                                                   const float* __restrict__ glow,
                                                                   const float* __restrict__ wy,
                                                                   const float* __restrict__ wx, float* __restrict__ gx,
                                                                   int planes, DctGeom gm) {
    extern __shared__ __align__(16) float sm[];
    const int H = gm.H, W = gm.W, Hh = gm.Hh, Wh = gm.Wh, PPL = gm.ppl;
    const int ldg = gm.ldw4_x;                 // G planes: [Hh rows (+overhang)][Wh] per quadrant, stride = 4 mod 32
    const int rows_g = ((Hh + 15) / 16) * 16;  // per quadrant
    const int rows_wt = ((H + 15) / 16) * 16;
    float* WyT = sm;                                  // [rows_wt][ldw4_y]   WyT[h][u] = Wy[u][h]
    float* Wxs = WyT + rows_wt * gm.ldw4_y;           // [W][ldx]            row-major, stride = 8 mod 32
    float* Gs = Wxs + (W + 8) * gm.ldx;               // [2 buf][PPL][2 quadrants][rows_g][ldg]
    float* Ss = Gs + 2 * PPL * 2 * rows_g * ldg;      // [PPL][H][ldt]       stride = 8 mod 32
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    for (int e = tid; e < rows_wt * gm.ldw4_y; e += kThreads) {
        const int h = e / gm.ldw4_y, u = e - h * gm.ldw4_y;
        WyT[e] = (h < H && u < H) ? wy[u * H + h] : 0.f;
    }
    for (int e = tid; e < (W + 8) * gm.ldx; e += kThreads) {
        const int r = e / gm.ldx, c = e - r * gm.ldx;
        Wxs[e] = (r < W && c < W) ? wx[r * W + c] : 0.f;
    }
    for (int e = tid; e < 2 * PPL * 2 * rows_g * ldg; e += kThreads) Gs[e] = 0.f;
    __syncthreads();

    const int groups = (planes + PPL - 1) / PPL;
    const int chunks_row = Wh / 4;
    auto stage = [&](int grp, int buf) {
        const int p0 = grp * PPL;
        const int np = min(PPL, planes - p0);
        const int per_q = Hh * chunks_row;
        const int total = np * 2 * per_q;
        for (int e = tid; e < total; e += kThreads) {
            const int pq = e / per_q, r2 = e - pq * per_q;  // pq = plane * 2 + quadrant
            const int row = r2 / chunks_row, c = r2 - row * chunks_row;
            const int pl = pq >> 1, qd = pq & 1;
            const float* src = (qd ? ghigh : glow) + (static_cast<size_t>(p0 + pl) * Hh + row) * Wh + c * 4;
            cp_async16(Gs + ((static_cast<size_t>(buf) * PPL + pl) * 2 + qd) * rows_g * ldg + row * ldg + c * 4, src);
        }
        cp_async_commit();
    };
    int it = 0;
    if (static_cast<int>(blockIdx.x) < groups) stage(blockIdx.x, 0);
    const int mt1 = rows_g / 16, ns1 = W / (8 * NT1);
    const int mt2 = (H + 15) / 16, ns2 = W / (8 * NT2);
    for (int grp = blockIdx.x; grp < groups; grp += gridDim.x, ++it) {
        const int buf = it & 1;
        const int nxt = grp + gridDim.x;
        if (nxt < groups) stage(nxt, buf ^ 1);
        if (nxt < groups) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const int p0 = grp * PPL;
        const int np = min(PPL, planes - p0);
        // ---- stage 1
        const int tasks1 = np * 2 * mt1 * ns1;
        for (int task = warp; task < tasks1; task += kWarps) {
            int r = task;
            const int pl = r / (2 * mt1 * ns1);
            r -= pl * (2 * mt1 * ns1);
            const int qd = r / (mt1 * ns1);
            r -= qd * (mt1 * ns1);
            const int mt = r / ns1, ns = r - mt * ns1;
            float acc[NT1][4];
#pragma unroll
            for (int i = 0; i < NT1; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            const float* A = Gs + ((static_cast<size_t>(buf) * PPL + pl) * 2 + qd) * rows_g * ldg + mt * 16 * ldg;
            // K = Wh rounded up to the k-step: the pad columns of G are zero, the pad rows of Wx finite
            warp_gemm<NT1>(acc, A, ldg, Wxs + qd * Wh * gm.ldx + ns * NT1 * 8, gm.ldx, 1, (Wh + 7) & ~7, lane);
            float* Sp = Ss + (pl * H + qd * Hh + mt * 16) * gm.ldt + ns * NT1 * 8;
            const int ua = mt * 16 + g, ub = ua + 8;
#pragma unroll
            for (int nt = 0; nt < NT1; ++nt) {
                if (ua < Hh) *reinterpret_cast<float2*>(Sp + g * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
                if (ub < Hh) *reinterpret_cast<float2*>(Sp + (g + 8) * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
        __syncthreads();
        // ---- stage 2
        const int tasks2 = np * mt2 * ns2;
        for (int task = warp; task < tasks2; task += kWarps) {
            const int pl = task / (mt2 * ns2), r = task - pl * (mt2 * ns2);
            const int mt = r / ns2, ns = r - mt * ns2;
            float acc[NT2][4];
#pragma unroll
            for (int i = 0; i < NT2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            warp_gemm<NT2>(acc, WyT + mt * 16 * gm.ldw4_y, gm.ldw4_y, Ss + pl * H * gm.ldt + ns * NT2 * 8, gm.ldt, 1, H,
                           lane);
            float* out = gx + static_cast<size_t>(p0 + pl) * H * W + ns * NT2 * 8;
            const int ha = mt * 16 + g, hb = ha + 8;
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) {
                if (ha < H) *reinterpret_cast<float2*>(out + ha * W + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
                if (hb < H) *reinterpret_cast<float2*>(out + hb * W + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ large planes
// Planes whose bases do not fit next to them in one SM's shared memory (H = W = 192 of the 768^2 stress
// configuration; supported: H = W in {64, 128, 192}).  Same contraction,
// same 3xTF32 tensor-core arithmetic, but every operand is STREAMED through double-buffered cp.async k-chunks
// (KC = 32) and the intermediate of one work item stays in shared memory:
//   forward  item = (plane, half q):  T_q = Wy[q*Hh:(q+1)*Hh, :] . X      (Hh x W, shared memory)
//                                     out_q = T_q . Wx[q*Wh:(q+1)*Wh, :]^T (Hh x Wh: low for q = 0, high for q = 1)
//   adjoint  item = (plane, column half c):  S = [G_0 . Wx[:Wh, cols] ; G_1 . Wx[Wh:, cols]]   (H x W/2, shared)
//                                            gx[:, cols] = Wy^T . S
// 12 warps; stage-1 warps own a 16-row x (N/2)-column block, stage-2 warps a 16-row block of the output.
constexpr int kKC = 32;

// A operand k-major: element (m, k) at A[k * ldk + m], ldk = 8 mod 32 (conflict-free fragment loads)
template <int NT>
__device__ __forceinline__ void warp_gemm_akm(float (&acc)[NT][4], const float* __restrict__ A, int ldk,
                                              const float* __restrict__ B, int ldb_k, int ldb_n, int K, int lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t ah[4], al[4];
        split_tf32(A[(k0 + t) * ldk + g], ah[0], al[0]);
        split_tf32(A[(k0 + t) * ldk + g + 8], ah[1], al[1]);
        split_tf32(A[(k0 + t + 4) * ldk + g], ah[2], al[2]);
        split_tf32(A[(k0 + t + 4) * ldk + g + 8], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            uint32_t bh[2], bl[2];
            const float* bp = B + (k0 + t) * ldb_k + (nt * 8 + g) * ldb_n;
            split_tf32(bp[0], bh[0], bl[0]);
            split_tf32(bp[4 * ldb_k], bh[1], bl[1]);
            mma_tf32(acc[nt], al, bh);
            mma_tf32(acc[nt], ah, bl);
            mma_tf32(acc[nt], ah, bh);
        }
    }
}

struct BigGeom {
    int H, W, Hh, Wh;
    int lda;   // [rows][kKC + 4]: row-major A chunks / n-major B chunks (4 mod 32)
    int ldb;   // forward stage A: X chunk rows [kKC][W + 8]; adjoint stage 1: Wx chunk rows [kKC][Wh + 8] (8 mod 32)
    int ldt;   // forward: T [Hh][W + 4] (4 mod 32);  adjoint: S [H][Wh + 8] (8 mod 32)
    int ldw;   // adjoint stage 2: Wy chunk rows [kKC][H + 8] (k-major A, 8 mod 32)
};

// rows x cols floats (cols % 4 == 0) global (row stride gs) -> shared (row stride ss), 16-byte cp.async
__device__ __forceinline__ void stage_tile(float* dst, int ss, const float* src, int gs, int rows, int cols, int tid) {
    const int cpr = cols >> 2;
    for (int e = tid; e < rows * cpr; e += kThreads) {
        const int r = e / cpr, c = e - r * cpr;
        cp_async16(dst + r * ss + c * 4, src + static_cast<size_t>(r) * gs + c * 4);
    }
}

template <int NTA, int NTB>  // NTA = W / 16, NTB = Wh / 16 (n-tiles of 8 per warp in stage A / B)
__global__ void __launch_bounds__(kThreads, 1) dct2d_fwd_big_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ wy,
                                                                    const float* __restrict__ wx,
                                                                    float* __restrict__ high, float* __restrict__ low,
                                                                    int planes, BigGeom gm) {
    extern __shared__ __align__(16) float sm[];
    const int H = gm.H, W = gm.W, Hh = gm.Hh, Wh = gm.Wh;
    float* As = sm;                          // [2][Hh][lda]   Wy / Wx row chunks
    float* Bs = As + 2 * Hh * gm.lda;        // [2][kKC][ldb]  X row chunks
    float* Ts = Bs + 2 * kKC * gm.ldb;       // [Hh][ldt]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mtiles = Hh / 16;              // <= 6
    const int mt = warp >> 1, nh = warp & 1; // stage tiles: 16 rows x half of the columns
    const bool active = mt < mtiles;
    const int items = planes * 2;
    const int nk = H / kKC;                  // k-chunks of stage A (over h) and of stage B (over w; W == H)
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int pl = item >> 1, q = item & 1;
        const float* X = x + static_cast<size_t>(pl) * H * W;
        const float* Wyq = wy + static_cast<size_t>(q) * Hh * H;   // rows q*Hh .. of Wy (H columns)
        const float* Wxq = wx + static_cast<size_t>(q) * Wh * W;   // rows q*Wh .. of Wx (W columns)
        auto issueA = [&](int kc, int buf) {
            stage_tile(As + buf * Hh * gm.lda, gm.lda, Wyq + kc * kKC, H, Hh, kKC, tid);
            stage_tile(Bs + buf * kKC * gm.ldb, gm.ldb, X + static_cast<size_t>(kc) * kKC * W, W, kKC, W, tid);
            cp_async_commit();
        };
        auto issueB = [&](int kc, int buf) {
            stage_tile(As + buf * Hh * gm.lda, gm.lda, Wxq + kc * kKC, W, Wh, kKC, tid);
            cp_async_commit();
        };
        // ---- stage A: T = Wy_q . X
        float acc[NTA][4];
#pragma unroll
        for (int i = 0; i < NTA; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        issueA(0, 0);
        for (int kc = 0; kc < nk; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < nk) issueA(kc + 1, buf ^ 1);
            if (kc + 1 < nk) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();
            if (active)
                warp_gemm<NTA>(acc, As + buf * Hh * gm.lda + mt * 16 * gm.lda, gm.lda,
                               Bs + buf * kKC * gm.ldb + nh * NTA * 8, gm.ldb, 1, kKC, lane);
            __syncthreads();
        }
        issueB(0, 0);  // the first Wx chunk travels while T is written
        if (active) {
            float* Tp = Ts + mt * 16 * gm.ldt + nh * NTA * 8;
#pragma unroll
            for (int nt = 0; nt < NTA; ++nt) {
                *reinterpret_cast<float2*>(Tp + g * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
                *reinterpret_cast<float2*>(Tp + (g + 8) * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
            }
        }
        // ---- stage B: out_q = T . Wx_q^T   (B[k = w][n = v] = Wx_q[v][w]: n-major chunk rows)
        float acc2[NTB][4];
#pragma unroll
        for (int i = 0; i < NTB; ++i) acc2[i][0] = acc2[i][1] = acc2[i][2] = acc2[i][3] = 0.f;
        for (int kc = 0; kc < nk; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < nk) issueB(kc + 1, buf ^ 1);
            if (kc + 1 < nk) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();  // (first pass: also publishes T)
            if (active)
                warp_gemm<NTB>(acc2, Ts + mt * 16 * gm.ldt + kc * kKC, gm.ldt,
                               As + buf * Hh * gm.lda + (nh * NTB * 8) * gm.lda, 1, gm.lda, kKC, lane);
            __syncthreads();
        }
        if (active) {
            float* out = (q ? high : low) + static_cast<size_t>(pl) * Hh * Wh + (mt * 16) * Wh + nh * NTB * 8;
#pragma unroll
            for (int nt = 0; nt < NTB; ++nt) {
                *reinterpret_cast<float2*>(out + g * Wh + nt * 8 + 2 * t) = make_float2(acc2[nt][0], acc2[nt][1]);
                *reinterpret_cast<float2*>(out + (g + 8) * Wh + nt * 8 + 2 * t) = make_float2(acc2[nt][2], acc2[nt][3]);
            }
        }
    }
}

template <int NT1, int NT2>  // NT1 = Wh / 16 (stage 1: half of the item's Wh columns per warp), NT2 = Wh / 8 (stage 2)
__global__ void __launch_bounds__(kThreads, 1) dct2d_bwd_big_kernel(const float* __restrict__ ghigh,
                                                                    const float* __restrict__ glow,
                                                                    const float* __restrict__ wy,
                                                                    const float* __restrict__ wx, float* __restrict__ gx,
                                                                    int planes, BigGeom gm) {
    extern __shared__ __align__(16) float sm[];
    const int H = gm.H, W = gm.W, Hh = gm.Hh, Wh = gm.Wh;
    float* As = sm;                          // [2][Hh][lda]   G_q column chunks (row-major A)
    float* Bs = As + 2 * Hh * gm.lda;        // [2][kKC][ldb]  Wx row chunks restricted to the item's columns
    float* Ws = Bs + 2 * kKC * gm.ldb;       // [2][kKC][ldw]  Wy row chunks (k-major A of stage 2)
    float* Ss = Ws + 2 * kKC * gm.ldw;       // [H][ldt]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mtiles1 = Hh / 16;             // <= 6
    const int mt1 = warp >> 1, nh1 = warp & 1;
    const bool active1 = mt1 < mtiles1;
    const int mtiles2 = H / 16;              // <= 12
    const bool active2 = warp < mtiles2;
    const int items = planes * 2;
    const int nk1 = Wh / kKC;                // per quadrant (over v)
    const int nk2 = H / kKC;                 // over u
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int pl = item >> 1, c = item & 1;  // c: column half [c*Wh, (c+1)*Wh) of gx / S
        auto issue1 = [&](int step, int buf) {   // step = q * nk1 + kc
            const int q = step / nk1, kc = step - q * nk1;
            const float* G = (q ? ghigh : glow) + static_cast<size_t>(pl) * Hh * Wh;
            stage_tile(As + buf * Hh * gm.lda, gm.lda, G + kc * kKC, Wh, Hh, kKC, tid);
            stage_tile(Bs + buf * kKC * gm.ldb, gm.ldb,
                       wx + (static_cast<size_t>(q) * Wh + kc * kKC) * W + c * Wh, W, kKC, Wh, tid);
            cp_async_commit();
        };
        auto issue2 = [&](int kc, int buf) {
            stage_tile(Ws + buf * kKC * gm.ldw, gm.ldw, wy + static_cast<size_t>(kc) * kKC * H, H, kKC, H, tid);
            cp_async_commit();
        };
        // ---- stage 1: S[q*Hh + u][w] = sum_v G_q[u][v] Wx[q*Wh + v][c*Wh + w]
        issue1(0, 0);
        for (int q = 0; q < 2; ++q) {
            float acc[NT1][4];
#pragma unroll
            for (int i = 0; i < NT1; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            for (int kc = 0; kc < nk1; ++kc) {
                const int step = q * nk1 + kc, buf = step & 1;
                if (step + 1 < 2 * nk1) issue1(step + 1, buf ^ 1);
                if (step + 1 < 2 * nk1) cp_async_wait<1>(); else cp_async_wait<0>();
                __syncthreads();
                if (active1)
                    warp_gemm<NT1>(acc, As + buf * Hh * gm.lda + mt1 * 16 * gm.lda, gm.lda,
                                   Bs + buf * kKC * gm.ldb + nh1 * NT1 * 8, gm.ldb, 1, kKC, lane);
                __syncthreads();
            }
            if (q == 1) issue2(0, 0);  // the first Wy chunk travels while S is completed
            if (active1) {
                float* Sp = Ss + (q * Hh + mt1 * 16) * gm.ldt + nh1 * NT1 * 8;
#pragma unroll
                for (int nt = 0; nt < NT1; ++nt) {
                    *reinterpret_cast<float2*>(Sp + g * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
                    *reinterpret_cast<float2*>(Sp + (g + 8) * gm.ldt + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
                }
            }
        }
        // ---- stage 2: gx[h][c*Wh + w] = sum_u Wy[u][h] S[u][w]   (A k-major: chunk rows u, columns h)
        float acc2[NT2][4];
#pragma unroll
        for (int i = 0; i < NT2; ++i) acc2[i][0] = acc2[i][1] = acc2[i][2] = acc2[i][3] = 0.f;
        for (int kc = 0; kc < nk2; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < nk2) issue2(kc + 1, buf ^ 1);
            if (kc + 1 < nk2) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncthreads();  // (first pass: also publishes S)
            if (active2)
                warp_gemm_akm<NT2>(acc2, Ws + buf * kKC * gm.ldw + warp * 16, gm.ldw, Ss + kc * kKC * gm.ldt, gm.ldt, 1,
                                   kKC, lane);
            __syncthreads();
        }
        if (active2) {
            float* out = gx + static_cast<size_t>(pl) * H * W + (warp * 16) * W + c * Wh;
#pragma unroll
            for (int nt = 0; nt < NT2; ++nt) {
                *reinterpret_cast<float2*>(out + g * W + nt * 8 + 2 * t) = make_float2(acc2[nt][0], acc2[nt][1]);
                *reinterpret_cast<float2*>(out + (g + 8) * W + nt * 8 + 2 * t) = make_float2(acc2[nt][2], acc2[nt][3]);
            }
        }
    }
}

bool big_supported(int H, int W) {
    return H == W && H % 64 == 0 && H <= 192;  // H/2 a multiple of the k-chunk (32) and of the m-tile (16)
}
BigGeom big_geom(int H, int W) {
    BigGeom g;
    g.H = H, g.W = W, g.Hh = H / 2, g.Wh = W / 2;
    g.lda = kKC + 4;
    return g;
}

struct Plan {
    DctGeom gm;
    size_t smem_fwd, smem_bwd;
    bool ok;
};

Plan make_plan(int H, int W) {
    Plan best{};
    best.ok = false;
    if (H % 8 != 0 || W % 8 != 0 || H < 16 || W < 16) return best;
    const size_t cap = 220 * 1024;
    const int nta = (W % 48 == 0) ? 6 : (W % 24 == 0) ? 3 : (W % 16 == 0) ? 2 : 1;
    double best_eff = -1.0;
    for (int ppl : {8, 6, 4, 3, 2, 1}) {
        Plan p{};
        DctGeom& g = p.gm;
        g.H = H, g.W = W, g.Hh = H / 2, g.Wh = W / 2;
        g.ldw4_y = pad_to(H, 4);
        g.ldw4_x = pad_to(W, 4);
        g.ldx = pad_to(W, 8);
        g.ppl = ppl;
        g.ldt = pad_to(W, 4);  // forward; the backward launcher overrides it with 8 mod 32
        g.rows_t = ((H + 15) / 16) * 16 + 16;
        const int rows_w = ((H + 15) / 16) * 16 + 16;
        p.smem_fwd = (static_cast<size_t>(rows_w) * g.ldw4_y + static_cast<size_t>(W + 16) * g.ldw4_x +
                      2ull * ppl * H * g.ldx + static_cast<size_t>(ppl) * g.rows_t * g.ldt) * 4;
        const int ldt_b = pad_to(W, 8);
        const int rows_g = ((g.Hh + 15) / 16) * 16;
        p.smem_bwd = (static_cast<size_t>(((H + 15) / 16) * 16) * g.ldw4_y + static_cast<size_t>(W + 8) * g.ldx +
                      2ull * ppl * 2 * rows_g * g.ldw4_x + static_cast<size_t>(ppl) * H * ldt_b) * 4;
        if (p.smem_fwd > cap || p.smem_bwd > cap) continue;
        const int tasks = ppl * ((H + 15) / 16) * (W / (8 * nta));  // stage-A tasks per stage
        const double eff = static_cast<double>(tasks) / (((tasks + kWarps - 1) / kWarps) * kWarps);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = p;
            best.ok = true;
        }
    }
    return best;
}

}  // namespace

bool dct2d_tc_supported(int H, int W) { return make_plan(H, W).ok || big_supported(H, W); }

template <int NTA, int NTB>
static cudaError_t launch_fwd_big(const float* x, const float* wy, const float* wx, float* high, float* low, int planes,
                                  int H, int W, cudaStream_t st) {
    BigGeom g = big_geom(H, W);
    g.ldb = pad_to(W, 8), g.ldt = pad_to(W, 4), g.ldw = 0;
    const size_t smem = (2ull * g.Hh * g.lda + 2ull * kKC * g.ldb + static_cast<size_t>(g.Hh) * g.ldt) * 4;
    auto kern = dct2d_fwd_big_kernel<NTA, NTB>;
    if (cudaError_t e = ensure_dyn_smem(kern, 220 * 1024); e != cudaSuccess) return e;
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;
    const int items = planes * 2, sms = device_sms();
    kern<<<items < sms ? items : sms, kThreads, smem, st>>>(x, wy, wx, high, low, planes, g);
    return cudaGetLastError();
}
template <int NT1, int NT2>
static cudaError_t launch_bwd_big(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                                  int planes, int H, int W, cudaStream_t st) {
    BigGeom g = big_geom(H, W);
    g.ldb = pad_to(g.Wh, 8), g.ldt = pad_to(g.Wh, 8), g.ldw = pad_to(H, 8);
    const size_t smem = (2ull * g.Hh * g.lda + 2ull * kKC * g.ldb + 2ull * kKC * g.ldw + static_cast<size_t>(H) * g.ldt) * 4;
    auto kern = dct2d_bwd_big_kernel<NT1, NT2>;
    if (cudaError_t e = ensure_dyn_smem(kern, 220 * 1024); e != cudaSuccess) return e;
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;
    const int items = planes * 2, sms = device_sms();
    kern<<<items < sms ? items : sms, kThreads, smem, st>>>(ghigh, glow, wy, wx, gx, planes, g);
    return cudaGetLastError();
}


cudaError_t launch_dct2d_fwd_tc(const float* x, const float* wy, const float* wx, float* high, float* low, int planes,
                                int H, int W, cudaStream_t st) {
    Plan p = make_plan(H, W);
    if (!p.ok) {  // bases + planes exceed one SM's shared memory: the streamed kernel
        if (!big_supported(H, W)) return cudaErrorInvalidConfiguration;
        cudaError_t e = cudaErrorInvalidConfiguration;
        switch (W) {
            case 192: e = launch_fwd_big<12, 6>(x, wy, wx, high, low, planes, H, W, st); break;
            case 128: e = launch_fwd_big<8, 4>(x, wy, wx, high, low, planes, H, W, st); break;
            case 64: e = launch_fwd_big<4, 2>(x, wy, wx, high, low, planes, H, W, st); break;
            default: break;
        }
        note_launch();
        return e;
    }
    const int sms = device_sms();
    const int groups = (planes + p.gm.ppl - 1) / p.gm.ppl;
    const int grid = groups < sms ? groups : sms;
    // NTA * 8 must divide W; NTB * 8 tiles the quadrant width
    auto launch = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<grid, kThreads, p.smem_fwd, st>>>(x, wy, wx, high, low, planes, p.gm);
        return cudaGetLastError();
    };
    cudaError_t e;
    if (W % 48 == 0) e = launch(dct2d_fwd_tc_kernel<6, 3>);       // W = 48, 96, 144, 192
    else if (W % 24 == 0) e = launch(dct2d_fwd_tc_kernel<3, 2>);  // W = 24, 72, 120
    else if (W % 16 == 0) e = launch(dct2d_fwd_tc_kernel<2, 2>);
    else e = launch(dct2d_fwd_tc_kernel<1, 1>);
    note_launch();
    return e;
}

cudaError_t launch_dct2d_bwd_tc(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                                int planes, int H, int W, cudaStream_t st) {
    Plan p = make_plan(H, W);
    if (!p.ok) {
        if (!big_supported(H, W)) return cudaErrorInvalidConfiguration;
        cudaError_t e = cudaErrorInvalidConfiguration;
        switch (W) {
            case 192: e = launch_bwd_big<6, 12>(ghigh, glow, wy, wx, gx, planes, H, W, st); break;
            case 128: e = launch_bwd_big<4, 8>(ghigh, glow, wy, wx, gx, planes, H, W, st); break;
            case 64: e = launch_bwd_big<2, 4>(ghigh, glow, wy, wx, gx, planes, H, W, st); break;
            default: break;
        }
        note_launch();
        return e;
    }
    const int sms = device_sms();
    DctGeom gm = p.gm;
    gm.ldt = pad_to(W, 8);
    const int groups = (planes + gm.ppl - 1) / gm.ppl;
    const int grid = groups < sms ? groups : sms;
    auto launch = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<grid, kThreads, p.smem_bwd, st>>>(ghigh, glow, wy, wx, gx, planes, gm);
        return cudaGetLastError();
    };
    cudaError_t e;
    if (W % 48 == 0) e = launch(dct2d_bwd_tc_kernel<6, 6>);
    else if (W % 24 == 0) e = launch(dct2d_bwd_tc_kernel<3, 3>);
    else if (W % 16 == 0) e = launch(dct2d_bwd_tc_kernel<2, 2>);
    else e = launch(dct2d_bwd_tc_kernel<1, 1>);
    note_launch();
    return e;
}

}  // namespace tramba
