// Micro-benchmarks that decide kernel design questions on the B200 (not product code):
//   1. TMA bulk reduce-add (cp.reduce.async.bulk .add.f32) shared -> global: chip throughput vs number of distinct
//      target arrays (contention), vs plain bulk store
//   2. red.global.add.v4.f32 from registers, same traffic pattern
//   3. FFMA vs FFMA2 (fma.rn.f32x2) issue throughput per SM
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ub/ubench.bin tools/ub/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0 = bulk store, 1 = bulk reduce add f32
__global__ void __launch_bounds__(256, 1) bulk_kernel(float* dst, int ntargets, int elems, int iters) {
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < elems; i += blockDim.x) sm[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    float* tgt = dst + (size_t)(blockIdx.x % ntargets) * elems;
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; ++it) {
            const uint32_t bytes = elems * 4;
            if (MODE == 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(tgt), "r"(s32(sm)), "r"(bytes) : "memory");
            else
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(tgt), "r"(s32(sm)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(512, 1) red_kernel(float* dst, int ntargets, int elems, int iters) {
    float* tgt = dst + (size_t)(blockIdx.x % ntargets) * elems;
    for (int it = 0; it < iters; ++it)
        for (int i = threadIdx.x * 4; i < elems; i += blockDim.x * 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(tgt + i), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
}

template <int MODE>  // 0 FFMA, 1 FFMA2, 2 FFMA + independent LOP interleave
__global__ void __launch_bounds__(512, 1) fma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    if (MODE == 0) {
        for (int i = 0; i < iters; ++i) {
            x0 = fmaf(x0, a, b), x1 = fmaf(x1, a, b), x2 = fmaf(x2, a, b), x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b), x5 = fmaf(x5, a, b), x6 = fmaf(x6, a, b), x7 = fmaf(x7, a, b);
        }
    } else {
        unsigned long long p0, p1, p2, p3, pa, pb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(x0), "f"(x1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(x2), "f"(x3));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(x4), "f"(x5));
        asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(x6), "f"(x7));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
        for (int i = 0; i < iters; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
        }
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(p0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(p1));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x4), "=f"(x5) : "l"(p2));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x6), "=f"(x7) : "l"(p3));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    printf("{\"device\": \"%s\", \"sms\": %d}\n", pr.name, sms);
    const int elems = 9216;  // one 96x96 plane of fp32 = 36 KB
    float* dst;
    CK(cudaMalloc(&dst, (size_t)sms * 2 * elems * 4));
    CK(cudaMemset(dst, 0, (size_t)sms * 2 * elems * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    CK(cudaFuncSetAttribute(bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int iters = 200;
    for (int mode = 0; mode < 3; ++mode) {
        for (int nt : {16, 32, 148, 296}) {
            for (int grid : {sms, 2 * sms}) {
                if (mode == 2 && grid != sms) continue;
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    if (mode == 0) bulk_kernel<0><<<grid, 256, elems * 4>>>(dst, nt, elems, iters);
                    if (mode == 1) bulk_kernel<1><<<grid, 256, elems * 4>>>(dst, nt, elems, iters);
                    if (mode == 2) red_kernel<<<grid, 512>>>(dst, nt, elems, iters);
                    cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1));
                    cudaEventElapsedTime(&ms, e0, e1);
                }
                const double bytes = (double)grid * iters * elems * 4;
                printf("{\"bench\": \"%s\", \"targets\": %d, \"grid\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n",
                       mode == 0 ? "bulk_store" : mode == 1 ? "bulk_reduce_add_f32" : "red.global.add.v4.f32", nt, grid, ms,
                       bytes / ms * 1e-6);
            }
        }
    }
    float* out;
    CK(cudaMalloc(&out, (size_t)sms * 512 * 4));
    for (int mode = 0; mode < 2; ++mode) {
        float ms = 0;
        const int it2 = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) fma_kernel<0><<<sms, 512>>>(out, it2, 1.0001f, 0.5f);
            else fma_kernel<1><<<sms, 512>>>(out, it2, 1.0001f, 0.5f);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double fmas = (double)sms * 512 * it2 * 8;
        printf("{\"bench\": \"%s\", \"ms\": %.4f, \"GFMA_per_s\": %.1f, \"fma_per_clk_per_sm_at_1965MHz\": %.1f}\n",
               mode == 0 ? "FFMA" : "FFMA2", ms, fmas / ms * 1e-6, fmas / (ms * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
