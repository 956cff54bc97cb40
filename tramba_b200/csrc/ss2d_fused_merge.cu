// Cross-merge through the skewed scan-order table (sm_100a): y[b,d,p] = sum_k sum_{t: idx[k,t]=p} ys[b,k,d,t].
// Used by the fused SS2D backward to fold the scan-ordered du back to image order
// (= CrossScan*.backward, Models/SS2D/csms6s.py:25-31,75-80,124-129,175-185).  Same data model as
// the fused forward (ss2d_fused.cuh): a round = PP whole planes; per direction the PP token rows
// arrive with one TMA bulk copy (double-buffered), thread (q, j) adds its S tokens into the
// skewed shared image plane; after the K directions the plane is written out once.  fp32
// accumulation, directions in order k = 0..K-1; duplicate hits of a pixel (helix lines) use shared
// fp32 atomics.  The deterministic, reference-association-order merge is tramba_cross_merge.
#include "ss2d_fused.cuh"

namespace tramba {

namespace {

constexpr int S = kFusedS;
constexpr int NT = kFusedNT;
constexpr int TOK = kFusedTok;

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct MergeLayout {
    size_t d, y, bar, total;
    int lsk;
    __host__ __device__ MergeLayout(int L, int es) {
        lsk = fused_skew_len(L);
        const int pp = TOK / L;
        size_t o = 0;
        d = o, o += 2 * static_cast<size_t>(TOK) * es;
        y = o, o += align_up(static_cast<size_t>(pp) * lsk * 4, 16);
        bar = o, o += 4 * sizeof(uint64_t);
        total = o;
    }
};

template <typename T>
__global__ void __launch_bounds__(NT, 2) merge_skewed_kernel(const T* __restrict__ ys, T* __restrict__ y,
                                                             const uint16_t* __restrict__ idx, int K, int D, int L,
                                                             int TP, int PP, int rounds, uint32_t dup_mask) {
    extern __shared__ __align__(128) unsigned char smem[];
    const MergeLayout lay(L, sizeof(T));
    const int LSK = lay.lsk;
    T* sd = reinterpret_cast<T*>(smem + lay.d);          // [2][TOK]
    float* sy = reinterpret_cast<float*>(smem + lay.y);  // [PP][LSK]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.bar);
    const int tid = threadIdx.x;
    const int q = tid / TP, j = tid - q * TP;

    const int my_rounds = (rounds - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                          static_cast<int>(gridDim.x);
    const int nstages = my_rounds * K;
    if (nstages <= 0) return;
    if (tid == 0) {
        mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    __syncthreads();
    auto issue = [&](int s) {  // thread 0
        const int i = s / K, k = s - i * K;
        const long long plane0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP;
        const long long b = plane0 / D, d0 = plane0 - b * D;
        mbar_expect_tx(&bar[s & 1], TOK * sizeof(T));
        bulk_g2s(sd + static_cast<size_t>(s & 1) * TOK, ys + ((b * K + k) * D + d0) * L, TOK * sizeof(T), &bar[s & 1]);
    };
    if (tid == 0) issue(0);
    float* yq = sy + static_cast<size_t>(q) * LSK;
    int s = 0;
    for (int i = 0; i < my_rounds; ++i) {
        const long long plane0 = (static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x) * PP;
        for (int k = 0; k < K; ++k, ++s) {
            if (tid == 0 && s + 1 < nstages) issue(s + 1);  // buffer (s+1)&1 was released by the barrier of stage s-1
            uint32_t pk[S / 2];
            const uint32_t* tp = reinterpret_cast<const uint32_t*>(idx + static_cast<size_t>(k) * L + j * S);
#pragma unroll
            for (int m = 0; m < S / 2; ++m) pk[m] = __ldg(tp + m);
            mbar_wait(&bar[s & 1], static_cast<uint32_t>((s >> 1) & 1));
            float v[S];
            load_tokens<S>(sd + static_cast<size_t>(s & 1) * TOK + static_cast<size_t>(q) * L + j * S, v);
            const bool dup = (dup_mask >> k) & 1u;
            if (k == 0) {
#pragma unroll
                for (int m = 0; m < S; ++m) yq[(m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu)] = v[m];
            } else if (!dup) {
#pragma unroll
                for (int m = 0; m < S; ++m) yq[(m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu)] += v[m];
            } else {
#pragma unroll
                for (int m = 0; m < S; ++m) atomicAdd(yq + ((m & 1) ? (pk[m >> 1] >> 16) : (pk[m >> 1] & 0xffffu)), v[m]);
            }
            __syncthreads();  // orders the scatters of consecutive directions and releases the stage buffer
        }
        // write the merged planes out (un-skew, convert)
        T* gy = y + plane0 * L;
        for (int e = tid * 4; e < TOK; e += NT * 4) {
            const int qq = e / L, pp = e - qq * L;
            const float* src = sy + static_cast<size_t>(qq) * LSK + fused_skew(pp);
            const float o[4] = {src[0], src[1], src[2], src[3]};
            store_vec<4>(gy + e, o);
        }
        // the next round's first scatter follows that round's first barrier-free stage: make the copy-out visible
        __syncthreads();
    }
}

}  // namespace

template <typename T>
static cudaError_t launch_t(const T* ys, T* y, const uint16_t* idx, int K, int B, int D, int L, uint32_t dup_mask,
                            cudaStream_t st) {
    auto kern = merge_skewed_kernel<T>;
    if (cudaError_t e = ensure_dyn_smem(kern, 227 * 1024); e != cudaSuccess) return e;
    const int sms = device_sms();
    const int TP = L / S, PP = NT / TP;
    const int rounds = B * D / PP;
    const size_t smem = MergeLayout(L, sizeof(T)).total;
    const int cap = 2 * sms;  // two CTAs per SM (75 KB each for bf16)
    const int grid = rounds < cap ? rounds : cap;
    kern<<<grid, NT, smem, st>>>(ys, y, idx, K, D, L, TP, PP, rounds, dup_mask);
    note_launch();
    return cudaGetLastError();
}

#ifdef TRAMBA_NEXT
cudaError_t launch_merge_skewed_cg(const void* ys, void* y, const uint16_t* idx, int K, int B, int D, int L,
                                   uint32_t dup_mask, int dtype, cudaStream_t st);  // csrc/next/
#endif

cudaError_t launch_merge_skewed(const void* ys, void* y, const uint16_t* idx, int K, int B, int D, int L,
                                uint32_t dup_mask, int dtype, cudaStream_t st) {
#ifdef TRAMBA_NEXT  // candidate build (libtramba_b200_next.so)
    const cudaError_t e = launch_merge_skewed_cg(ys, y, idx, K, B, D, L, dup_mask, dtype, st);
    if (e != cudaErrorNotSupported) return e;
#endif
    if (dtype == TRAMBA_F32)
        return launch_t(static_cast<const float*>(ys), static_cast<float*>(y), idx, K, B, D, L, dup_mask, st);
    if (dtype == TRAMBA_BF16)
        return launch_t(static_cast<const __nv_bfloat16*>(ys), static_cast<__nv_bfloat16*>(y), idx, K, B, D, L,
                        dup_mask, st);
    return cudaErrorInvalidValue;
}

}  // namespace tramba
