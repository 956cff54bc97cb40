/* tramba_b200.h -- C ABI of the B200-native (sm_100a) 2D selective-scan hot path of Tramba.
 *
 * This is the drop-in boundary: plain pointers, sizes and enums; no torch types.  Device
 * pointers are raw CUDA device addresses owned by the caller (torch's caching allocator in
 * the Python host); nothing is allocated or freed here; inputs are never mutated.  `stream`
 * is a cudaStream_t passed as void* (NULL = legacy default stream).  Every entry point
 * returns 0 on success, a negative TRAMBA_E_* code for a rejected argument, or a positive
 * cudaError_t value for a CUDA failure; tramba_error_string() renders all three.
 * There is no CPU fallback: the compute entry points need an sm_100 device.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef TRAMBA_B200_H
#define TRAMBA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRAMBA_ABI_VERSION 1

#if defined(__GNUC__)
#define TRAMBA_API __attribute__((visibility("default")))
#else
#define TRAMBA_API
#endif

typedef enum { TRAMBA_F32 = 0, TRAMBA_BF16 = 1, TRAMBA_F16 = 2 } tramba_dtype_t;

/* Scan-order families.  LINE is the Helix-SS2D order: 4 standard directions followed by the
 * 4 Bresenham radial-line lists (K = 8).  CUSTOM takes a caller-supplied table with arbitrary
 * (possibly repeating) indices.  Values 5..11 are the reference's ablation orders (not on the
 * Tramba-V path): they run through the table-driven scan / merge kernels,
 * never through the fused core. */
typedef enum {
    TRAMBA_ORDER_STD = 0,      /* Models/SS2D/csms6s.py:13-55   CrossScan / CrossMerge            */
    TRAMBA_ORDER_DILATION = 1, /* Models/SS2D/csms6s.py:64-103  + Models/SS2D/Dilation.py:3-96    */
    TRAMBA_ORDER_WINDOW = 2,   /* Models/SS2D/csms6s.py:113-152 + Models/SS2D/Window.py:3-86      */
    TRAMBA_ORDER_LINE = 3,     /* Models/SS2D/csms6s.py:161-216 + Models/SS2D/SpiralLine.py:3-133 */
    TRAMBA_ORDER_CUSTOM = 4,
    TRAMBA_ORDER_LINE4 = 5,    /* csms6s.py:219-265  CrossScan_Line_4direction: the 4 helix lists, K = 4  */
    TRAMBA_ORDER_SPIRAL = 6,   /* csms6s.py:264-368 + Spiral.py:3-82: std + cw/ccw spirals + flips, K = 8 */
    TRAMBA_ORDER_HILBERT = 7,  /* csms6s.py:371-473 + Hilbert.py:9-331 (sides 2^m and 3*2^m), K = 4        */
    TRAMBA_ORDER_DIAGONAL = 8, /* csms6s.py:478-615: std + (anti-)diagonal wraps + flips, K = 8           */
    TRAMBA_ORDER_DS = 9,       /* csms6s.py:619-674: helix lists + diagonal wraps + flips, K = 8          */
    TRAMBA_ORDER_AB2 = 10,     /* csms6s.py:678-710  CrossScan_Ab_2direction, K = 4                       */
    TRAMBA_ORDER_AB1 = 11      /* csms6s.py:713-737  CrossScan_Ab_1direction, K = 4                       */
} tramba_order_t;

enum {
    TRAMBA_OK = 0,
    TRAMBA_E_NULL = -1,      /* a required pointer is NULL                         */
    TRAMBA_E_SHAPE = -2,     /* a size is non-positive or inconsistent             */
    TRAMBA_E_DTYPE = -3,     /* unsupported dtype combination                      */
    TRAMBA_E_ORDER = -4,     /* unknown order / parameter not valid for this H     */
    TRAMBA_E_ALIGN = -5,     /* a pointer is not aligned to its element size       */
    TRAMBA_E_NODEVICE = -6,  /* no CUDA device / not an sm_100 device              */
    TRAMBA_E_LIMIT = -7      /* size exceeds what the kernels support (see docs)   */
};

TRAMBA_API int tramba_abi_version(void);
TRAMBA_API const char* tramba_error_string(int code);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
TRAMBA_API uint64_t tramba_launch_count(void);

/* ----------------------------------------------------------------------------------------
 * Scan-order tables (host code, no GPU needed).
 * Replaces generate_dilation_indices (Models/SS2D/Dilation.py:3-45), generate_window_indices
 * (Models/SS2D/Window.py:3-35), generate_indices (Models/SS2D/SpiralLine.py:27-82) and the
 * module-level dicts at Models/SS2D/csms6s.py:58-61,107-110,157-158.  The table is the flat
 * pixel index idx[k*L + t] (row*W + col), K = tramba_order_num_directions(order), L = H*W;
 * square maps only (H == W), like the reference.  `param` = dilation rate / window size
 * (ignored by every other family).  Bit-exact with the reference generators.
 * -------------------------------------------------------------------------------------- */
TRAMBA_API int tramba_order_num_directions(int order);
TRAMBA_API int tramba_order_table(int order, int H, int W, int param, int32_t* idx_out /* K*L, host */);

/* Order plan = what the device kernels consume: one int32 blob of
 * tramba_order_plan_elems(K, L) = K*(3L+1) entries laid out as
 *   [0, K*L)              forward table idx[k][t]                    (cross-scan gathers with it)
 *   [K*L, 2*K*L)          inverse entries: tokens t sorted by (pixel, t) per direction
 *   [2*K*L, +K*(L+1))     CSR offsets per direction: tokens of pixel p are
 *                         entries[k][off[k][p] .. off[k][p+1])        (cross-merge gathers with it)
 * The inverse replaces the reference's scatter_add_ (Dilation.py:88-95, Window.py:78-85,
 * SpiralLine.py:125-132) by a deterministic gather that adds duplicate hits in token order,
 * i.e. exactly the order of a sequential scatter_add_.  _from_table builds a plan for a
 * caller-supplied forward table (TRAMBA_ORDER_CUSTOM). */
TRAMBA_API size_t tramba_order_plan_elems(int K, int L);
TRAMBA_API int tramba_order_plan(int order, int H, int W, int param, int32_t* plan_out /* host */);
TRAMBA_API int tramba_order_plan_from_table(int K, int L, const int32_t* table, int32_t* plan_out /* host */);

/* ----------------------------------------------------------------------------------------
 * Cross-scan: xs[b,k,c,t] = x[b,c].flat[idx[k,t]].
 * Replaces CrossScan*.forward and CrossMerge*.backward (Models/SS2D/csms6s.py:15-22,45-55,
 * 66-72,94-103,115-121,143-152,163-172,204-216).  x: (B,C,H,W) contiguous; xs: (B,K,C,L).
 * `plan_dev`: device copy of the order plan.  dtype: element type of x and xs.
 * -------------------------------------------------------------------------------------- */
TRAMBA_API int tramba_cross_scan(const void* x, void* xs, const int32_t* plan_dev, int order, int K,
                      int B, int C, int H, int W, int dtype, void* stream);

/* Cross-merge: y[b,d,p] = sum_k sum_{t: idx[k,t]==p} ys[b,k,d,t]  (the exact transpose).
 * Replaces CrossMerge*.forward and CrossScan*.backward (Models/SS2D/csms6s.py:25-31,36-42,
 * 75-80,85-91,124-129,134-140,175-185,190-201) incl. get_re_*_scan's scatter_add_.
 * ys: (B,K,D,L) contiguous; y: (B,D,L), same dtype.  Accumulation is fp32, deterministic, and
 * follows the reference's association order: STD (ys0+flip ys2) + T(ys1+flip ys3);
 * DILATION/WINDOW/CUSTOM sequential over k then t from zero; LINE = (sequential k=4..7) +
 * (STD fold of k=0..3) -- bit-exact with the reference evaluated on CPU.
 * -------------------------------------------------------------------------------------- */
TRAMBA_API int tramba_cross_merge(const void* ys, void* y, const int32_t* plan_dev, int order, int K,
                       int B, int D, int H, int W, int dtype, void* stream);

/* ----------------------------------------------------------------------------------------
 * Selective scan (S6 recurrence).
 * Replaces selective_scan_cuda_oflex.fwd / .bwd as called at Models/SS2D/csms6s.py:910 and
 * :920-922 (operator SelectiveScanOflex, csms6s.py:904-923; arguments prepared at
 * Models/vmamba.py:244-257).
 *   u, delta            (batch, dim, L)        in_dtype, contiguous, L innermost
 *   A                   (dim, N)               fp32
 *   Bm, Cm              (batch, groups, N, L)  in_dtype
 *   D, delta_bias       (dim)                  fp32, either may be NULL (treated as 0)
 *   out                 (batch, dim, L)        out_dtype (F32 when oflex=True)
 *   ckpt                tramba_sscan_ckpt_elems() floats, or NULL for inference; opaque chunk
 *                       boundary states handed back to _bwd (the reference's `x` tensor)
 * dim % groups == 0; row d uses group d / (dim/groups); 1 <= N <= 64.
 * -------------------------------------------------------------------------------------- */
TRAMBA_API size_t tramba_sscan_ckpt_elems(int batch, int dim, int L, int N, int in_dtype);

TRAMBA_API int tramba_sscan_fwd(const void* u, const void* delta, const float* A, const void* Bm,
                     const void* Cm, const float* D, const float* delta_bias, void* out,
                     float* ckpt, int batch, int dim, int L, int N, int groups, int in_dtype,
                     int out_dtype, int delta_softplus, void* stream);

/* Backward.  dout: (batch, dim, L) dout_dtype.  Outputs: du, ddelta (in_dtype); dA (dim,N),
 * dD (dim), ddelta_bias (dim) fp32 summed over batch and L; dB, dC (batch,groups,N,L) fp32
 * summed over the channels of a group.  dD / ddelta_bias may be NULL when D / delta_bias are.
 * All outputs are overwritten (zero-initialised internally where accumulated). */
TRAMBA_API int tramba_sscan_bwd(const void* u, const void* delta, const float* A, const void* Bm,
                     const void* Cm, const float* D, const float* delta_bias, const void* dout,
                     const float* ckpt, void* du, void* ddelta, float* dA, float* dB, float* dC,
                     float* dD, float* ddelta_bias, int batch, int dim, int L, int N,
                     int groups, int in_dtype, int dout_dtype, int delta_softplus, void* stream);

/* ----------------------------------------------------------------------------------------
 * Fused SS2D core: cross-scan -> selective scan -> cross-merge in one kernel (forward) and its
 * adjoint (backward), for d_state N = 1 (every SS2D of Tramba-V).  Replaces the composition in
 * SS2Dv2.forward_corev2, Models/vmamba.py:231-259, i.e. CrossScan*.apply (:231) ->
 * SelectiveScanOflex.apply (:255-257) -> CrossMerge*.apply (:259), without materialising the
 * permuted copies xs / ys of the feature map.
 *   x      (B, D, L)        in_dtype   IMAGE order (L = H*W, square maps); gathered in shared memory
 *   dts    (B, K, D, L)     in_dtype   scan order: the dt_proj output, vmamba.py:236
 *   Bs, Cs (B, K, L)        in_dtype   scan order
 *   A, Dp, delta_bias (K*D) fp32       Dp / delta_bias may be NULL
 *   idx16  (K, L)           uint16, device: SKEWED pixel p + 2*(p >> 6) of token t of direction k
 *                           (tramba_order_table16_skewed; tramba_order_table16 is the plain table,
 *                           bit-exact with the reference generators)
 *   y      (B, D, L)        fp32       IMAGE order, merged over the K directions (oflex=True contract)
 * Supported: order STD / DILATION / WINDOW / LINE, N == 1, L in {144, 576, 2304, 9216} (H = 12, 24, 48, 96 -- every
 * Tramba-V map at 384^2) and D a multiple of 512*18/L (tramba_ss2d_fused_supported says which; everything else runs
 * through the un-fused entry points above).  Forward and backward, F32 and BF16, at every one of these sizes (the
 * backward in F32 at L = 9216 reads the B / C rows from global memory instead of staging them).  Helix-line directions add duplicate hits of a pixel with shared-memory fp32 atomics, so their merged
 * sums are order-nondeterministic at rounding level, like the reference's CUDA scatter_add_.
 *
 * Backward = CrossMerge*.backward -> SelectiveScanOflex.backward -> CrossScan*.backward (csms6s.py:45-55, 914-923,
 * 25-31) in ONE kernel: dy (B, D, L) fp32 image order in; dx (B, D, L) fp32 IMAGE order out, already summed over the
 * K directions (TMA bulk reductions at L2) -- the K-fold gradient du w.r.t. xs never exists in HBM;
 * ddts (B, K, D, L) in_dtype scan order; dBs, dCs (B, K, L) fp32 (summed over the channels: per-item TMA bulk reductions);
 * dA, dD, dbias (K*D) fp32 TOTALS over batch and tokens (reds).  Every fp32 output is an accumulation target and is
 * zeroed here -- with one memset when dx, dBs, dCs, dA, dD, dbias are carved contiguously, in that order, out of one
 * allocation.
 *   sched (K, sched_len) uint32, device: tramba_order_sched -- per direction a bank-conflict-free schedule of the
 *          shared-memory permutation image order <-> token order (entry = 2*pixel | 2*token << 16, 0xFFFFFFFF = none);
 *          tramba_order_sched_len = entries per direction (row stride; 0 = unsupported size)
 *   kinds (2K) int32, HOST, as tramba_order_sched writes them: [k] = 0 identity, 1 reversed identity, 2 permutation,
 *          3 table with repeated pixels; [K + k] = schedule entries direction k uses (<= sched_len)
 * -------------------------------------------------------------------------------------- */
TRAMBA_API int tramba_order_table16(int order, int H, int W, int param, uint16_t* idx_out /* K*L, host */);
TRAMBA_API int tramba_order_table16_skewed(int order, int H, int W, int param, uint16_t* idx_out /* K*L, host */);
TRAMBA_API size_t tramba_order_sched_len(int order, int H, int W, int param);
TRAMBA_API int tramba_order_sched(int order, int H, int W, int param, uint32_t* sched_out /* K*len, host */,
                                  int32_t* kind_out /* 2K, host */);
TRAMBA_API int tramba_ss2d_fused_supported(int order, int K, int D, int H, int W, int N, int in_dtype,
                                           int need_bwd);
TRAMBA_API int tramba_ss2d_fused_fwd(const void* x, const void* dts, const void* Bs, const void* Cs,
                                     const float* A, const float* Dp, const float* delta_bias,
                                     const uint16_t* idx16, float* y, int order, int K, int B, int D, int H,
                                     int W, int N, int in_dtype, int delta_softplus, void* stream);
TRAMBA_API int tramba_ss2d_fused_bwd(const void* x, const void* dts, const void* Bs, const void* Cs,
                                     const float* A, const float* Dp, const float* delta_bias,
                                     const uint32_t* sched, int sched_len, const int32_t* kinds,
                                     const float* dy, float* dx, void* ddts, float* dBs, float* dCs,
                                     float* dA, float* dD, float* dbias, int order, int K, int B,
                                     int D, int H, int W, int N, int in_dtype, int delta_softplus,
                                     void* stream);

/* ----------------------------------------------------------------------------------------
 * DCT2D high/low split.  Replaces DCT2D.forward (Models/DCT_2D.py:12-29) and the two basis
 * layers (DCT_2D.py:32-77): Y = Wy . X . Wx^T per plane, high = Y[H/2:, W/2:],
 * low = Y[:H/2, :W/2].  x: (planes, H, W) fp32; wy: (H,H), wx: (W,W) fp32 bases (the
 * module's registered buffers dct_y.weight / dct_x.weight); high, low: (planes, H/2, W/2).
 * H and W even.  _bwd is the adjoint: gx = Wy^T . dY . Wx with dY zero outside the quadrants.
 * -------------------------------------------------------------------------------------- */
TRAMBA_API void tramba_dct_basis(int n, float* w_out /* n*n, host */); /* DCT_2D.py:37-45 */
/* 1 when a plane of this size fits the kernels (bases + plane resident in one SM's shared memory: up to
 * ~128 x 128); larger planes (H = 192 of the 768^2 stress config) return TRAMBA_E_LIMIT from _fwd/_bwd. */
TRAMBA_API int tramba_dct2d_supported(int H, int W);

TRAMBA_API int tramba_dct2d_fwd(const float* x, const float* wy, const float* wx, float* high, float* low,
                     int planes, int H, int W, void* stream);
TRAMBA_API int tramba_dct2d_bwd(const float* ghigh, const float* glow, const float* wy, const float* wx,
                     float* gx, int planes, int H, int W, void* stream);

/* ----------------------------------------------------------------------------------------
 * Projected-channel gather (host side of the fused core).  x_proj is a per-token linear map, so the
 * fused path evaluates it once on the image-ordered map; these two entry points carry its R + 2N
 * output channels into scan order and split them -- what CrossScan* -> x_proj -> torch.split produce
 * in the reference (Models/vmamba.py:231-235), at 1/25 of the traffic -- and back:
 *   gather : src (B, K, R+2N, L) image order -> dts (B, K, R, L), Bs (B, K, N, L), Cs (B, K, N, L),
 *            out[b,k,c,t] = src[b,k,c, idx[k][t]]
 *   scatter: the adjoint, gsrc[b,k,c,p] = sum_{t: idx[k][t]==p} g[b,k,c,t], as a deterministic gather
 *            through the plan's inverse (CSR) table (no atomics; duplicate helix hits in token order).
 * plan_dev = device copy of tramba_order_plan(order, H, W, param).  dtype F32 / BF16; the scatter's
 * gBs / gCs may be F32 while gdts / gsrc are BF16 (bc_dtype).
 * -------------------------------------------------------------------------------------- */
TRAMBA_API int tramba_proj_gather(const void* src, void* dts, void* Bs, void* Cs, const int32_t* plan_dev, int K,
                                  int B, int R, int N, int L, int dtype, void* stream);
/* out (G, n) = sum_b in (G, B, n), fp32 accumulation, same dtype in and out: the reduction over the image batch of the
 * x_proj / dt_proj weight gradients, which the host evaluates as ONE batched GEMM dY_b . X_b^T (the reference's
 * grouped conv1d weight gradient, Models/vmamba.py:233-236 backward), and the fold of the per-tile weight / bias
 * gradient partials of LayerNorm2d (G = 2: dweight | dbias in one launch) and the depthwise conv.  in / out 16-byte
 * aligned; G > 1 needs B * n * elem_size % 16 == 0. */
TRAMBA_API int tramba_batch_sum(const void* in, void* out, int G, int B, long long n, int dtype, void* stream);
TRAMBA_API int tramba_proj_scatter(const void* gdts, const void* gBs, const void* gCs, void* gsrc,
                                   const int32_t* plan_dev, int K, int B, int R, int N, int L, int dtype, int bc_dtype,
                                   void* stream);

/* ----------------------------------------------------------------------------------------
 * LayerNorm2d (+ GELU): layer norm over the CHANNEL axis of a channel-first (B, C, L) map.
 * Replaces LayerNorm2d.forward (Models/modules.py:22-27: permute -> F.layer_norm -> permute)
 * and, with act = TRAMBA_ACT_GELU, the out_norm -> out_act chain applied to the merged map
 * after every cross-merge (Models/vmamba.py:271 and :287) -- the op next to the path (SURVEY
 * 8(f) item 2).  One HBM read and one write of the map instead of 6-8 map-sized passes.
 *   x (B, C, L) in_dtype; weight (C) fp32; bias (C) fp32 or NULL; y (B, C, L) out_dtype =
 *   act(((x - mean_c) * rstd_c) * weight + bias), statistics over C per (b, l), biased variance,
 *   fp32 internals; mean, rstd (B, L) fp32 saved for the backward (NULL = do not save).
 * Backward: gy (B, C, L) g_dtype; gx (B, C, L) in_dtype; gw_part / gb_part
 * (B * tramba_layernorm2d_tiles(B, C, L), C) fp32 per-tile partial sums of dweight / dbias (the
 * caller adds the rows; no atomics, deterministic).  gb_part may be NULL.
 * in/out/g dtypes: F32 or BF16.  x, y, gy, gx must be 8-byte aligned.
 * -------------------------------------------------------------------------------------- */
enum { TRAMBA_ACT_NONE = 0, TRAMBA_ACT_GELU = 1, TRAMBA_ACT_SILU = 2 };
TRAMBA_API int tramba_layernorm2d_tiles(int B, int C, int L);
TRAMBA_API int tramba_layernorm2d_fwd(const void* x, const float* weight, const float* bias, void* y, float* mean,
                                      float* rstd, int B, int C, int L, float eps, int act, int in_dtype,
                                      int out_dtype, void* stream);
TRAMBA_API int tramba_layernorm2d_bwd(const void* gy, const void* x, const float* weight, const float* bias,
                                      const float* mean, const float* rstd, void* gx, float* gw_part, float* gb_part,
                                      int B, int C, int L, int act, int in_dtype, int g_dtype, void* stream);

/* ----------------------------------------------------------------------------------------
 * Depthwise k x k convolution (+ bias, + SiLU) on channel-first planes, stride 1, "same" zero
 * padding.  Replaces `x = self.act(self.conv2d(x))` of SS2D.forward (Models/vmamba.py:281-285;
 * nn.Conv2d(d_inner, d_inner, groups=d_inner, kernel_size=d_conv, padding=(d_conv-1)//2),
 * vmamba.py:72-79; act = SiLU) -- the producer of the scan's input.
 *   x, y (B, C, H, W) dtype (F32 / BF16); weight (C, 1, K, K) fp32; bias (C) fp32 or NULL;
 *   K in {3, 5, 7}; act TRAMBA_ACT_NONE or TRAMBA_ACT_SILU; fp32 accumulation.
 * Backward: g = dy * act'(conv(x, w) + b) (pre-activation recomputed); dx = conv(g, flip(w));
 * dw_part (B*C, K*K + 1): per-plane partial sums of dweight (K*K taps) and dbias (last column); the
 * caller adds over B.
 * Planes up to ~100 x 100 (tramba_dwconv2d_supported); pointers 16-byte aligned.
 * -------------------------------------------------------------------------------------- */
TRAMBA_API int tramba_dwconv2d_supported(int B, int C, int H, int W, int K);
TRAMBA_API int tramba_dwconv2d_fwd(const void* x, const float* weight, const float* bias, void* y, int B, int C,
                                   int H, int W, int K, int act, int dtype, void* stream);
TRAMBA_API int tramba_dwconv2d_bwd(const void* x, const float* weight, const float* bias, const void* dy, void* dx,
                                   float* dw_part, int B, int C, int H, int W, int K, int act, int dtype,
                                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRAMBA_B200_H */
