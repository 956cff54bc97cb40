// extern "C" entry points of libtramba_b200.so (see include/tramba_b200.h).
// Argument validation, plan selection and kernel launch; no torch, no allocation.
#include <algorithm>
#include <array>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "sscan.cuh"
#include "ss2d_fused.cuh"

namespace tramba {

static std::atomic<uint64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int* dyn_smem_slot(const void* kern, int dev) {
    static std::mutex mu;
    static std::unordered_map<const void*, std::array<int, kMaxDevices>> table;  // node-based: slots never move
    std::lock_guard<std::mutex> lock(mu);
    auto it = table.find(kern);
    if (it == table.end()) it = table.emplace(kern, std::array<int, kMaxDevices>{}).first;
    return &it->second[dev];
}

cudaError_t launch_cross_scan(const void* x, void* xs, const int32_t* plan, int K, int B, int C, int L,
                              int dtype, cudaStream_t st);
cudaError_t launch_cross_merge(const void* ys, void* y, const int32_t* plan, int order, int K, int B, int D,
                               int L, int dtype, cudaStream_t st);
cudaError_t launch_dct2d_fwd(const float* x, const float* wy, const float* wx, float* high, float* low,
                             int planes, int H, int W, cudaStream_t st);
cudaError_t launch_dct2d_bwd(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                             int planes, int H, int W, cudaStream_t st);

cudaError_t launch_batch_sum(const void* in, void* out, int G, int B, long long n, int dtype, cudaStream_t st);

bool dct2d_tc_supported(int H, int W);
cudaError_t launch_dct2d_fwd_tc(const float* x, const float* wy, const float* wx, float* high, float* low, int planes,
                                int H, int W, cudaStream_t st);
cudaError_t launch_dct2d_bwd_tc(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                                int planes, int H, int W, cudaStream_t st);

int layernorm2d_tiles(int B, int C, int L);
cudaError_t launch_layernorm2d_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd,
                                   int B, int C, int L, float eps, int act, int in_dtype, int out_dtype,
                                   cudaStream_t st);
cudaError_t launch_layernorm2d_bwd(const void* gy, const void* x, const float* w, const float* b, const float* mean,
                                   const float* rstd, void* gx, float* gw_part, float* gb_part, int B, int C, int L,
                                   int act, int in_dtype, int g_dtype, cudaStream_t st);

bool dwconv2d_supported(int B, int C, int H, int W, int K);
cudaError_t launch_dwconv2d_fwd(const void* x, const float* w, const float* b, void* y, int B, int C, int H, int W,
                                int K, int act, int dtype, cudaStream_t st);
cudaError_t launch_dwconv2d_bwd(const void* x, const float* w, const float* b, const void* dy, void* dx, float* dwp,
                                int B, int C, int H, int W, int K, int act, int dtype, cudaStream_t st);

cudaError_t launch_proj_gather(const void* src, void* dts, void* Bs, void* Cs, const int32_t* plan, int K, int B, int R,
                               int N, int L, int dtype, cudaStream_t st);
cudaError_t launch_proj_scatter(const void* gdts, const void* gBs, const void* gCs, void* gsrc, const int32_t* plan,
                                int K, int B, int R, int N, int L, int dtype, int bc_dtype, cudaStream_t st);

static int elem_size(int dtype) {
    switch (dtype) {
        case TRAMBA_F32: return 4;
        case TRAMBA_BF16:
        case TRAMBA_F16: return 2;
        default: return 0;
    }
}
static bool aligned(const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// One-time device check: the kernels are compiled for sm_100a only.
static int check_device() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_rc = 0;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return TRAMBA_E_NODEVICE;
    }
    if (dev == cached_dev) return cached_rc;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return TRAMBA_E_NODEVICE;
    }
    cached_dev = dev;
    cached_rc = (major == 10) ? TRAMBA_OK : TRAMBA_E_NODEVICE;
    return cached_rc;
}

static int cuda_rc(cudaError_t e) { return e == cudaSuccess ? TRAMBA_OK : static_cast<int>(e); }

}  // namespace tramba

using namespace tramba;

extern "C" int tramba_abi_version(void) { return TRAMBA_ABI_VERSION; }

extern "C" uint64_t tramba_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* tramba_error_string(int code) {
    switch (code) {
        case TRAMBA_OK: return "ok";
        case TRAMBA_E_NULL: return "tramba: a required pointer is NULL";
        case TRAMBA_E_SHAPE: return "tramba: a size is non-positive or inconsistent";
        case TRAMBA_E_DTYPE: return "tramba: unsupported dtype combination";
        case TRAMBA_E_ORDER: return "tramba: unknown scan order or invalid order parameter";
        case TRAMBA_E_ALIGN: return "tramba: pointer not sufficiently aligned (16 bytes required)";
        case TRAMBA_E_NODEVICE: return "tramba: no CUDA device with compute capability 10.x (sm_100a) is current";
        case TRAMBA_E_LIMIT: return "tramba: size exceeds what the kernels support";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "tramba: unknown error";
}

// ------------------------------------------------------------------------------ order plans
extern "C" size_t tramba_order_plan_elems(int K, int L) {
    if (K <= 0 || L <= 0) return 0;
    return static_cast<size_t>(K) * (3 * static_cast<size_t>(L) + 1);
}

extern "C" int tramba_order_plan_from_table(int K, int L, const int32_t* table, int32_t* plan) {
    if (!table || !plan) return TRAMBA_E_NULL;
    if (K <= 0 || L <= 0) return TRAMBA_E_SHAPE;
    int32_t* fwd = plan;
    int32_t* ent = plan + static_cast<size_t>(K) * L;
    int32_t* off = plan + 2 * static_cast<size_t>(K) * L;
    if (fwd != table) std::memmove(fwd, table, sizeof(int32_t) * static_cast<size_t>(K) * L);
    std::vector<int32_t> cur(L + 1);
    for (int k = 0; k < K; ++k) {
        const int32_t* tk = fwd + static_cast<size_t>(k) * L;
        int32_t* ok = off + static_cast<size_t>(k) * (L + 1);
        int32_t* ek = ent + static_cast<size_t>(k) * L;
        std::fill(ok, ok + L + 1, 0);
        for (int t = 0; t < L; ++t) {
            if (tk[t] < 0 || tk[t] >= L) return TRAMBA_E_ORDER;
            ++ok[tk[t] + 1];
        }
        for (int p = 0; p < L; ++p) ok[p + 1] += ok[p];
        std::copy(ok, ok + L + 1, cur.begin());
        for (int t = 0; t < L; ++t) ek[cur[tk[t]]++] = t;  // stable: ascending t per pixel
    }
    return TRAMBA_OK;
}

extern "C" int tramba_order_plan(int order, int H, int W, int param, int32_t* plan) {
    const int K = tramba_order_num_directions(order);
    if (K < 0) return K;
    int rc = tramba_order_table(order, H, W, param, plan);
    if (rc != TRAMBA_OK) return rc;
    return tramba_order_plan_from_table(K, H * W, plan, plan);
}

// ------------------------------------------------------------------------------ scan / merge
static int check_scan_args(const void* a, const void* b, const int32_t* plan, int order, int K, int B, int C,
                           int H, int W, int dtype) {
    if (!a || !b || !plan) return TRAMBA_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || K <= 0) return TRAMBA_E_SHAPE;
    if (order < TRAMBA_ORDER_STD || order > TRAMBA_ORDER_AB1) return TRAMBA_E_ORDER;
    if (order != TRAMBA_ORDER_CUSTOM && K != tramba_order_num_directions(order)) return TRAMBA_E_ORDER;
    if (dtype != TRAMBA_F32 && dtype != TRAMBA_BF16) return TRAMBA_E_DTYPE;
    if (!aligned(a, elem_size(dtype)) || !aligned(b, elem_size(dtype)) || !aligned(plan, 4)) return TRAMBA_E_ALIGN;
    if (static_cast<long long>(H) * W > (1ll << 24)) return TRAMBA_E_LIMIT;
    if (static_cast<long long>(B) * C >= (1ll << 31)) return TRAMBA_E_LIMIT;
    return check_device();
}

extern "C" int tramba_cross_scan(const void* x, void* xs, const int32_t* plan_dev, int order, int K, int B,
                                 int C, int H, int W, int dtype, void* stream) {
    int rc = check_scan_args(x, xs, plan_dev, order, K, B, C, H, W, dtype);
    if (rc != TRAMBA_OK) return rc;
    cudaError_t e = launch_cross_scan(x, xs, plan_dev, K, B, C, H * W, dtype, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidConfiguration) return TRAMBA_E_LIMIT;
    return cuda_rc(e);
}

extern "C" int tramba_cross_merge(const void* ys, void* y, const int32_t* plan_dev, int order, int K, int B,
                                  int D, int H, int W, int dtype, void* stream) {
    int rc = check_scan_args(ys, y, plan_dev, order, K, B, D, H, W, dtype);
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(
        launch_cross_merge(ys, y, plan_dev, order, K, B, D, H * W, dtype, static_cast<cudaStream_t>(stream)));
}

// ------------------------------------------------------------------------------ selective scan
static int sscan_common(int batch, int dim, int L, int N, int groups, int in_dtype, int io_dtype) {
    if (batch <= 0 || dim <= 0 || L <= 0 || N <= 0 || groups <= 0) return TRAMBA_E_SHAPE;
    if (dim % groups != 0) return TRAMBA_E_SHAPE;
    if (N > kMaxState) return TRAMBA_E_LIMIT;
    if (static_cast<long long>(batch) * dim >= (1ll << 31)) return TRAMBA_E_LIMIT;
    if (in_dtype != TRAMBA_F32 && in_dtype != TRAMBA_BF16) return TRAMBA_E_DTYPE;
    if (io_dtype != TRAMBA_F32 && io_dtype != in_dtype) return TRAMBA_E_DTYPE;
    return TRAMBA_OK;
}

static void fill_args(SScanArgs& a, const SScanPlan& p, int batch, int dim, int L, int N, int groups,
                      int softplus) {
    a.batch = batch, a.dim = dim, a.L = L, a.N = N, a.groups = groups, a.dpg = dim / groups;
    a.rows = batch * dim;
    a.R = p.R, a.cpu = p.cpu, a.units = p.units;
    a.softplus = softplus ? 1 : 0;
    a.uniform = p.uniform ? 1 : 0;
    a.rpc = p.rpc;
    a.rw = 1;
    a.divL.init(static_cast<uint32_t>(L));
    a.divDim.init(static_cast<uint32_t>(dim));
    a.divDpg.init(static_cast<uint32_t>(dim / groups));
}

extern "C" size_t tramba_sscan_ckpt_elems(int batch, int dim, int L, int N, int in_dtype) {
    if (batch <= 0 || dim <= 0 || L <= 0 || N <= 0 || elem_size(in_dtype) == 0) return 0;
    const SScanPlan p = make_sscan_plan(static_cast<long long>(batch) * dim, L, elem_size(in_dtype));
    return static_cast<size_t>(p.units) * p.cpu * N;
}

extern "C" int tramba_sscan_fwd(const void* u, const void* delta, const float* A, const void* Bm,
                                const void* Cm, const float* D, const float* delta_bias, void* out,
                                float* ckpt, int batch, int dim, int L, int N, int groups, int in_dtype,
                                int out_dtype, int delta_softplus, void* stream) {
    if (!u || !delta || !A || !Bm || !Cm || !out) return TRAMBA_E_NULL;
    int rc = sscan_common(batch, dim, L, N, groups, in_dtype, out_dtype);
    if (rc != TRAMBA_OK) return rc;
    const SScanPlan p = make_sscan_plan(static_cast<long long>(batch) * dim, L, elem_size(in_dtype));
    const int need = p.tma ? 16 : elem_size(in_dtype);
    if (!aligned(u, need) || !aligned(delta, need) || !aligned(Bm, need) || !aligned(Cm, need) ||
        !aligned(out, p.tma ? 16 : elem_size(out_dtype)) || !aligned(A, 4) || !aligned(D, 4) ||
        !aligned(delta_bias, 4) || !aligned(ckpt, 4))
        return TRAMBA_E_ALIGN;
    rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    SScanArgs a;
    std::memset(&a, 0, sizeof(a));
    a.u = u, a.delta = delta, a.A = A, a.Bm = Bm, a.Cm = Cm, a.D = D, a.bias = delta_bias;
    a.out = out, a.ckpt = ckpt;
    fill_args(a, p, batch, dim, L, N, groups, delta_softplus);
    return cuda_rc(launch_sscan_fwd(a, p, in_dtype, out_dtype, static_cast<cudaStream_t>(stream)));
}

extern "C" int tramba_sscan_bwd(const void* u, const void* delta, const float* A, const void* Bm,
                                const void* Cm, const float* D, const float* delta_bias, const void* dout,
                                const float* ckpt, void* du, void* ddelta, float* dA, float* dB, float* dC,
                                float* dD, float* ddelta_bias, int batch, int dim, int L, int N, int groups,
                                int in_dtype, int dout_dtype, int delta_softplus, void* stream) {
    if (!u || !delta || !A || !Bm || !Cm || !dout || !ckpt || !du || !ddelta || !dA || !dB || !dC)
        return TRAMBA_E_NULL;
    if ((D && !dD) || (delta_bias && !ddelta_bias)) return TRAMBA_E_NULL;
    int rc = sscan_common(batch, dim, L, N, groups, in_dtype, dout_dtype);
    if (rc != TRAMBA_OK) return rc;
    SScanPlan p = make_sscan_plan(static_cast<long long>(batch) * dim, L, elem_size(in_dtype));
    // The side-by-side kernel (rpc rows per chunk) takes one B/C base per chunk and reduces dB/dC over the
    // sub-rows: all of them must belong to one (batch, group).  Otherwise: the generic kernel (same chunking,
    // per-lane coordinates).  The forward's checkpoints are not read when a chunk holds whole rows (cpu == 1).
    if (p.regular && p.rpc > 1 && (dim / groups) % p.rpc != 0) p.regular = false, p.rpc = 1;
    const int need = p.tma ? 16 : elem_size(in_dtype);
    if (!aligned(u, need) || !aligned(delta, need) || !aligned(Bm, need) || !aligned(Cm, need) ||
        !aligned(du, need) || !aligned(ddelta, need) || !aligned(dout, p.tma ? 16 : elem_size(dout_dtype)) ||
        !aligned(dB, 16) || !aligned(dC, 16) || !aligned(dA, 4) || !aligned(ckpt, 4))
        return TRAMBA_E_ALIGN;
    rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t nbc = static_cast<size_t>(batch) * groups * N * L * sizeof(float);
    cudaError_t e = cudaMemsetAsync(dB, 0, nbc, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dC, 0, nbc, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dA, 0, static_cast<size_t>(dim) * N * sizeof(float), st);
    if (e == cudaSuccess && dD) e = cudaMemsetAsync(dD, 0, static_cast<size_t>(dim) * sizeof(float), st);
    if (e == cudaSuccess && ddelta_bias) e = cudaMemsetAsync(ddelta_bias, 0, static_cast<size_t>(dim) * sizeof(float), st);
    if (e != cudaSuccess) return cuda_rc(e);
    SScanArgs a;
    std::memset(&a, 0, sizeof(a));
    a.u = u, a.delta = delta, a.A = A, a.Bm = Bm, a.Cm = Cm, a.D = D, a.bias = delta_bias;
    a.dout = dout, a.ckpt_in = ckpt, a.du = du, a.ddelta = ddelta;
    a.dA = dA, a.dB = dB, a.dC = dC, a.dD = dD, a.dbias = ddelta_bias;
    fill_args(a, p, batch, dim, L, N, groups, delta_softplus);
    // dB/dC pre-reduction over RW units of one (batch, group): largest RW in {8,4,2,1} that
    // still leaves at least ~12 warps of work per SM (one warp per super-unit).
    if (p.regular) {
        if (N == 1) {
            int sms = 148, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int rows_per_group = dim / groups;
            for (int rw = 8; rw >= 1; rw >>= 1) {
                if (rows_per_group % (rw * p.rpc) != 0) continue;
                a.rw = rw;
                if (p.units / rw >= 12ll * sms) break;
            }
        }
    } else if (p.uniform && N == 1 && (dim / groups) % 8 == 0) {
        a.rw = 8;
    }
    return cuda_rc(launch_sscan_bwd(a, p, in_dtype, dout_dtype, st));
}

// ------------------------------------------------------------------------------ DCT2D
extern "C" int tramba_dct2d_supported(int H, int W) {
    if (H <= 0 || W <= 0 || (H & 1) || (W & 1)) return 0;
    if (dct2d_tc_supported(H, W)) return 1;
    // SIMT kernel: plane + intermediate + both bases in one CTA's shared memory (dct2d.cu)
    const size_t smem = (2ull * H * W + static_cast<size_t>(W) * W + static_cast<size_t>(H) * H) * 4;
    return smem <= 220 * 1024 ? 1 : 0;
}

extern "C" int tramba_dct2d_fwd(const float* x, const float* wy, const float* wx, float* high, float* low,
                                int planes, int H, int W, void* stream) {
    if (!x || !wy || !wx || !high || !low) return TRAMBA_E_NULL;
    if (planes <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return TRAMBA_E_SHAPE;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    if (!aligned(x, 16) || !aligned(high, 8) || !aligned(low, 8)) return TRAMBA_E_ALIGN;
    // tensor-core path (TF32 x3) for every shape it covers; exact-fp32 SIMT kernel otherwise
    cudaError_t e = dct2d_tc_supported(H, W)
                        ? launch_dct2d_fwd_tc(x, wy, wx, high, low, planes, H, W, static_cast<cudaStream_t>(stream))
                        : launch_dct2d_fwd(x, wy, wx, high, low, planes, H, W, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidConfiguration) return TRAMBA_E_LIMIT;
    return cuda_rc(e);
}

extern "C" int tramba_dct2d_bwd(const float* ghigh, const float* glow, const float* wy, const float* wx,
                                float* gx, int planes, int H, int W, void* stream) {
    if (!ghigh || !glow || !wy || !wx || !gx) return TRAMBA_E_NULL;
    if (planes <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return TRAMBA_E_SHAPE;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    if (!aligned(ghigh, 16) || !aligned(glow, 16) || !aligned(gx, 8)) return TRAMBA_E_ALIGN;
    cudaError_t e = dct2d_tc_supported(H, W)
                        ? launch_dct2d_bwd_tc(ghigh, glow, wy, wx, gx, planes, H, W, static_cast<cudaStream_t>(stream))
                        : launch_dct2d_bwd(ghigh, glow, wy, wx, gx, planes, H, W, static_cast<cudaStream_t>(stream));
    if (e == cudaErrorInvalidConfiguration) return TRAMBA_E_LIMIT;
    return cuda_rc(e);
}


// ------------------------------------------------------------------------------ fused SS2D core
extern "C" int tramba_order_table16(int order, int H, int W, int param, uint16_t* idx_out) {
    if (!idx_out) return TRAMBA_E_NULL;
    const int K = tramba_order_num_directions(order);
    if (K < 0) return K;
    if (static_cast<long long>(H) * W > 65536) return TRAMBA_E_LIMIT;
    std::vector<int32_t> t(static_cast<size_t>(K) * H * W);
    int rc = tramba_order_table(order, H, W, param, t.data());
    if (rc != TRAMBA_OK) return rc;
    for (size_t i = 0; i < t.size(); ++i) idx_out[i] = static_cast<uint16_t>(t[i]);
    return TRAMBA_OK;
}

// The table the fused kernels consume: skewed shared-memory position p + 2*(p >> 6) of each token's pixel.
extern "C" int tramba_order_table16_skewed(int order, int H, int W, int param, uint16_t* idx_out) {
    int rc = tramba_order_table16(order, H, W, param, idx_out);
    if (rc != TRAMBA_OK) return rc;
    const int K = tramba_order_num_directions(order);
    const size_t n = static_cast<size_t>(K) * H * W;
    if (fused_skew(H * W - 1) > 65535) return TRAMBA_E_LIMIT;
    for (size_t i = 0; i < n; ++i) idx_out[i] = static_cast<uint16_t>(fused_skew(idx_out[i]));
    return TRAMBA_OK;
}

static uint32_t fused_dup_mask(int order) { return order == TRAMBA_ORDER_LINE ? 0xF0u : 0u; }

static int fused_geometry(FusedArgs& a, int order, int K, int B, int D, int H, int W, int N) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || K <= 0) return TRAMBA_E_SHAPE;
    if (order < TRAMBA_ORDER_STD || order > TRAMBA_ORDER_LINE) return TRAMBA_E_ORDER;
    if (K != tramba_order_num_directions(order)) return TRAMBA_E_ORDER;
    if (N != 1) return TRAMBA_E_LIMIT;
    const long long L = static_cast<long long>(H) * W;
    if (L > 65536 || L % kFusedS != 0) return TRAMBA_E_LIMIT;
    const int TP = static_cast<int>(L / kFusedS);
    if (TP != 8 && TP != 32 && TP != 128 && TP != 512) return TRAMBA_E_LIMIT;  // L = 144, 576, 2304, 9216 (H = 12 .. 96)
    const int PP = kFusedNT / TP;
    if (D % PP != 0) return TRAMBA_E_LIMIT;
    if (static_cast<long long>(B) * D / PP >= (1ll << 30)) return TRAMBA_E_LIMIT;
    a.B = B, a.K = K, a.D = D, a.L = static_cast<int>(L), a.TP = TP, a.PP = PP;
    a.rounds = B * D / PP;
    a.dup_mask = fused_dup_mask(order);
    return TRAMBA_OK;
}

extern "C" int tramba_ss2d_fused_supported(int order, int K, int D, int H, int W, int N, int in_dtype,
                                           int need_bwd) {
    FusedArgs a;
    std::memset(&a, 0, sizeof(a));
    const int es = elem_size(in_dtype);
    if (in_dtype != TRAMBA_F32 && in_dtype != TRAMBA_BF16) return 0;
    if (fused_geometry(a, order, K, 1, D, H, W, N) != TRAMBA_OK) return 0;
    if (ss2d_fused_fwd_smem(a.L, es, 1) > 227 * 1024) return 0;
    if (need_bwd && !ss2d_fused_bwd_fits(a.L, es)) return 0;
    return 1;
}

extern "C" int tramba_ss2d_fused_fwd(const void* x, const void* dts, const void* Bs, const void* Cs,
                                     const float* A, const float* Dp, const float* delta_bias,
                                     const uint16_t* idx16, float* y, int order, int K, int B, int D, int H,
                                     int W, int N, int in_dtype, int delta_softplus, void* stream) {
    if (!x || !dts || !Bs || !Cs || !A || !idx16 || !y) return TRAMBA_E_NULL;
    if (in_dtype != TRAMBA_F32 && in_dtype != TRAMBA_BF16) return TRAMBA_E_DTYPE;
    FusedArgs a;
    std::memset(&a, 0, sizeof(a));
    int rc = fused_geometry(a, order, K, B, D, H, W, N);
    if (rc != TRAMBA_OK) return rc;
    if (!aligned(x, 16) || !aligned(dts, 16) || !aligned(Bs, 16) || !aligned(Cs, 16) || !aligned(y, 16) ||
        !aligned(idx16, 4) || !aligned(A, 4) || !aligned(Dp, 4) || !aligned(delta_bias, 4))
        return TRAMBA_E_ALIGN;
    rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    const int es = elem_size(in_dtype);
    a.xb = ss2d_fused_fwd_smem(a.L, es, 2) <= 227 * 1024 ? 2 : 1;
    if (ss2d_fused_fwd_smem(a.L, es, a.xb) > 227 * 1024) return TRAMBA_E_LIMIT;
    a.x = x, a.dts = dts, a.Bs = Bs, a.Cs = Cs, a.A = A, a.Dp = Dp, a.bias = delta_bias, a.idx = idx16, a.y = y;
    a.softplus = delta_softplus ? 1 : 0;
    return cuda_rc(launch_ss2d_fused_fwd(a, in_dtype, static_cast<cudaStream_t>(stream)));
}

// Backward with dx: the merge of du (CrossScan*.backward, csms6s.py:25-31) happens inside the kernel.
extern "C" int tramba_ss2d_fused_bwd(const void* x, const void* dts, const void* Bs, const void* Cs,
                                        const float* A, const float* Dp, const float* delta_bias,
                                        const uint32_t* sched, int sched_len, const int32_t* kinds,
                                        const float* dy, float* dx, void* ddts, float* dBs, float* dCs,
                                        float* dA_part, float* dD_part, float* dbias_part, int order, int K, int B,
                                        int D, int H, int W, int N, int in_dtype, int delta_softplus,
                                        void* stream) {
    if (!x || !dts || !Bs || !Cs || !A || !sched || !kinds || !dy || !dx || !ddts || !dBs || !dCs || !dA_part)
        return TRAMBA_E_NULL;
    if ((Dp && !dD_part) || (delta_bias && !dbias_part)) return TRAMBA_E_NULL;
    if (in_dtype != TRAMBA_F32 && in_dtype != TRAMBA_BF16) return TRAMBA_E_DTYPE;
    FusedArgs g;
    std::memset(&g, 0, sizeof(g));
    int rc = fused_geometry(g, order, K, B, D, H, W, N);
    if (rc != TRAMBA_OK) return rc;
    if (!aligned(x, 16) || !aligned(dts, 16) || !aligned(Bs, 16) || !aligned(Cs, 16) || !aligned(dy, 16) ||
        !aligned(dx, 16) || !aligned(ddts, 16) || !aligned(dBs, 16) || !aligned(dCs, 16) || !aligned(sched, 8))
        return TRAMBA_E_ALIGN;
    if (sched_len <= 0 || sched_len % (2 * g.TP) != 0 || sched_len % 64 != 0) return TRAMBA_E_SHAPE;
    if (!ss2d_fused_bwd_fits(g.L, elem_size(in_dtype))) return TRAMBA_E_LIMIT;
    rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // every fp32 output is an accumulation target: zero them, with as few memset nodes as the caller's layout
    // allows (the Python host carves all six out of one workspace -> one memset)
    const size_t nbc = static_cast<size_t>(B) * K * g.L * sizeof(float);
    const size_t nrow = static_cast<size_t>(K) * D * sizeof(float);
    std::vector<std::pair<char*, size_t>> z = {{reinterpret_cast<char*>(dx), static_cast<size_t>(B) * D * g.L * sizeof(float)},
                                               {reinterpret_cast<char*>(dBs), nbc},
                                               {reinterpret_cast<char*>(dCs), nbc},
                                               {reinterpret_cast<char*>(dA_part), nrow}};
    if (dD_part) z.push_back({reinterpret_cast<char*>(dD_part), nrow});
    if (dbias_part) z.push_back({reinterpret_cast<char*>(dbias_part), nrow});
    std::sort(z.begin(), z.end());
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < z.size() && e == cudaSuccess;) {
        char* p0 = z[i].first;
        size_t n = z[i].second;
        size_t jn = i + 1;
        while (jn < z.size() && z[jn].first == p0 + n) n += z[jn].second, ++jn;
        e = cudaMemsetAsync(p0, 0, n, st);
        i = jn;
    }
    if (e != cudaSuccess) return cuda_rc(e);
    FusedBwdArgs a;
    std::memset(&a, 0, sizeof(a));
    a.x = x, a.dts = dts, a.Bs = Bs, a.Cs = Cs, a.A = A, a.Dp = Dp, a.bias = delta_bias, a.sched = sched;
    a.dy = dy, a.dx = dx, a.ddts = ddts, a.dBs = dBs, a.dCs = dCs, a.dA = dA_part, a.dD = dD_part,
    a.dbias = dbias_part;
    a.B = B, a.K = K, a.D = D, a.sched_len = sched_len;
    a.kinds = 0;
    if (K > 8) return TRAMBA_E_LIMIT;
    for (int k = 0; k < K; ++k) {
        if (kinds[k] < 0 || kinds[k] > 3) return TRAMBA_E_ORDER;
        a.kinds |= static_cast<uint32_t>(kinds[k]) << (4 * k);
        const int used = kinds[K + k];
        if (used <= 0 || used > sched_len || used % (2 * g.TP) != 0) return TRAMBA_E_SHAPE;
        a.sched_used[k] = used;
    }
    a.softplus = delta_softplus ? 1 : 0;
    e = launch_ss2d_fused_bwd(a, g.L, in_dtype, st);
    if (e == cudaErrorInvalidConfiguration) return TRAMBA_E_LIMIT;
    return cuda_rc(e);
}


// ------------------------------------------------------------------------------ LayerNorm2d
static bool ln_dtype_ok(int d) { return d == TRAMBA_F32 || d == TRAMBA_BF16; }

extern "C" int tramba_layernorm2d_tiles(int B, int C, int L) {
    if (B <= 0 || C <= 0 || L <= 0) return TRAMBA_E_SHAPE;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return layernorm2d_tiles(B, C, L);
}

extern "C" int tramba_layernorm2d_fwd(const void* x, const float* weight, const float* bias, void* y, float* mean,
                                      float* rstd, int B, int C, int L, float eps, int act, int in_dtype,
                                      int out_dtype, void* stream) {
    if (!x || !weight || !y) return TRAMBA_E_NULL;
    if (B <= 0 || C <= 0 || L <= 0 || B > 65535) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(in_dtype) || !ln_dtype_ok(out_dtype)) return TRAMBA_E_DTYPE;
    if (act != TRAMBA_ACT_NONE && act != TRAMBA_ACT_GELU) return TRAMBA_E_SHAPE;
    if (!aligned(x, 8) || !aligned(y, 8) || (mean && !aligned(mean, 8)) || (rstd && !aligned(rstd, 8)))
        return TRAMBA_E_ALIGN;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_layernorm2d_fwd(x, weight, bias, y, mean, rstd, B, C, L, eps, act, in_dtype, out_dtype,
                                          static_cast<cudaStream_t>(stream)));
}

extern "C" int tramba_layernorm2d_bwd(const void* gy, const void* x, const float* weight, const float* bias,
                                      const float* mean, const float* rstd, void* gx, float* gw_part, float* gb_part,
                                      int B, int C, int L, int act, int in_dtype, int g_dtype, void* stream) {
    if (!gy || !x || !weight || !mean || !rstd || !gx || !gw_part) return TRAMBA_E_NULL;
    if (B <= 0 || C <= 0 || L <= 0 || B > 65535) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(in_dtype) || !ln_dtype_ok(g_dtype)) return TRAMBA_E_DTYPE;
    if (act != TRAMBA_ACT_NONE && act != TRAMBA_ACT_GELU) return TRAMBA_E_SHAPE;
    if (!aligned(x, 8) || !aligned(gy, 8) || !aligned(gx, 8) || !aligned(mean, 8) || !aligned(rstd, 8))
        return TRAMBA_E_ALIGN;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_layernorm2d_bwd(gy, x, weight, bias, mean, rstd, gx, gw_part, gb_part, B, C, L, act,
                                          in_dtype, g_dtype, static_cast<cudaStream_t>(stream)));
}

// ------------------------------------------------------------------------------ depthwise conv
extern "C" int tramba_dwconv2d_supported(int B, int C, int H, int W, int K) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return dwconv2d_supported(B, C, H, W, K) ? 1 : 0;
}

extern "C" int tramba_dwconv2d_fwd(const void* x, const float* weight, const float* bias, void* y, int B, int C,
                                   int H, int W, int K, int act, int dtype, void* stream) {
    if (!x || !weight || !y) return TRAMBA_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(dtype)) return TRAMBA_E_DTYPE;
    if (act != TRAMBA_ACT_NONE && act != TRAMBA_ACT_SILU) return TRAMBA_E_SHAPE;
    if (!dwconv2d_supported(B, C, H, W, K)) return TRAMBA_E_LIMIT;
    if (!aligned(x, 16) || !aligned(y, 16)) return TRAMBA_E_ALIGN;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_dwconv2d_fwd(x, weight, bias, y, B, C, H, W, K, act != TRAMBA_ACT_NONE, dtype,
                                       static_cast<cudaStream_t>(stream)));
}

extern "C" int tramba_dwconv2d_bwd(const void* x, const float* weight, const float* bias, const void* dy, void* dx,
                                   float* dw_part, int B, int C, int H, int W, int K, int act, int dtype,
                                   void* stream) {
    if (!x || !weight || !dy || !dx || !dw_part) return TRAMBA_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(dtype)) return TRAMBA_E_DTYPE;
    if (act != TRAMBA_ACT_NONE && act != TRAMBA_ACT_SILU) return TRAMBA_E_SHAPE;
    if (!dwconv2d_supported(B, C, H, W, K)) return TRAMBA_E_LIMIT;
    if (!aligned(x, 16) || !aligned(dy, 16) || !aligned(dx, 16)) return TRAMBA_E_ALIGN;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_dwconv2d_bwd(x, weight, bias, dy, dx, dw_part, B, C, H, W, K, act != TRAMBA_ACT_NONE, dtype,
                                       static_cast<cudaStream_t>(stream)));
}

// ------------------------------------------------------------------------------ projected-channel gather
extern "C" int tramba_proj_gather(const void* src, void* dts, void* Bs, void* Cs, const int32_t* plan_dev, int K,
                                  int B, int R, int N, int L, int dtype, void* stream) {
    if (!src || !dts || !Bs || !Cs || !plan_dev) return TRAMBA_E_NULL;
    if (K <= 0 || B <= 0 || R <= 0 || N <= 0 || L <= 0) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(dtype)) return TRAMBA_E_DTYPE;
    if (static_cast<size_t>(L) * 4 > 200 * 1024) return TRAMBA_E_LIMIT;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_proj_gather(src, dts, Bs, Cs, plan_dev, K, B, R, N, L, dtype, static_cast<cudaStream_t>(stream)));
}

// Batch reduction of the projection weight gradients (Models/vmamba.py:233-236 backward) and of per-tile partials:
// out (G, n) = sum_b in (G, B, n).
extern "C" int tramba_batch_sum(const void* in, void* out, int G, int B, long long n, int dtype, void* stream) {
    if (!in || !out) return TRAMBA_E_NULL;
    if (G <= 0 || G > 65535 || B <= 0 || n <= 0) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(dtype)) return TRAMBA_E_DTYPE;
    if (!aligned(in, 16) || !aligned(out, 16)) return TRAMBA_E_ALIGN;  // (rows: n * elem_size % 16 != 0 -> scalar path)
    if (G > 1 && (n * (dtype == TRAMBA_F32 ? 4 : 2)) % 16 != 0) return TRAMBA_E_ALIGN;  // every group 16-byte aligned
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_batch_sum(in, out, G, B, n, dtype, static_cast<cudaStream_t>(stream)));
}

extern "C" int tramba_proj_scatter(const void* gdts, const void* gBs, const void* gCs, void* gsrc,
                                   const int32_t* plan_dev, int K, int B, int R, int N, int L, int dtype, int bc_dtype,
                                   void* stream) {
    if (!gdts || !gBs || !gCs || !gsrc || !plan_dev) return TRAMBA_E_NULL;
    if (K <= 0 || B <= 0 || R <= 0 || N <= 0 || L <= 0) return TRAMBA_E_SHAPE;
    if (!ln_dtype_ok(dtype) || !ln_dtype_ok(bc_dtype) || (dtype == TRAMBA_F32 && bc_dtype != TRAMBA_F32))
        return TRAMBA_E_DTYPE;
    if (static_cast<size_t>(L) * 4 > 200 * 1024) return TRAMBA_E_LIMIT;
    int rc = check_device();
    if (rc != TRAMBA_OK) return rc;
    return cuda_rc(launch_proj_scatter(gdts, gBs, gCs, gsrc, plan_dev, K, B, R, N, L, dtype, bc_dtype,
                                       static_cast<cudaStream_t>(stream)));
}
