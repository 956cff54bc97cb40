// Self-checks of the simulator runtime (tools/cpu_sim/sim_common.h): shuffles, barriers, mbarriers, and the
// detection of the two failure modes it exists to catch -- a barrier mismatch and an mbarrier that never completes.
#include "sim_common.h"

namespace {

void warp_sum_kernel(const float* in, float* out) {     // butterfly sum + inclusive scan over 16-lane segments
    const int tid = threadIdx.x;
    float v = in[tid], s = in[tid];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    for (int off = 1; off < 16; off <<= 1) {
        const float p = __shfl_up_sync(0xffffffffu, s, off, 16);
        if ((tid & 15) >= off) s += p;
    }
    out[tid] = v;
    out[blockDim.x + tid] = s;
    __syncthreads();
    if (tid == 0) out[2 * blockDim.x] = out[blockDim.x - 1] + static_cast<float>(blockIdx.x);
}

void mismatch_kernel() {                                  // thread 5 skips the barrier
    if (threadIdx.x != 5) __syncthreads();
    __syncthreads();
    __syncthreads();
}

void mbar_kernel(int complete, float* out) {              // producer / consumer over an mbarrier with a transaction count
    unsigned char* smem = sim::smem();
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* buf = reinterpret_cast<float*>(smem + 16);
    static float src[4] = {1.f, 2.f, 3.f, 4.f};
    if (threadIdx.x == 0) tramba::mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        tramba::mbar_expect_tx(bar, 16);
        if (complete) tramba::bulk_g2s(buf, src, 16, bar);
    }
    tramba::mbar_wait(bar, 0);
    out[threadIdx.x] = buf[threadIdx.x & 3];
}

}  // namespace

extern "C" const char* sim_selftest_shuffle(const float* in, float* out, int grid) {
    return sim::launch(grid, 64, 64, [&] { warp_sum_kernel(in, out); });
}
extern "C" const char* sim_selftest_mismatch() {
    return sim::launch(1, 64, 64, [] { mismatch_kernel(); });
}
extern "C" const char* sim_selftest_mbar(int complete, float* out) {
    return sim::launch(1, 32, 64, [&] { mbar_kernel(complete, out); });
}
extern "C" void sim_set_lazy(int on) { sim::lazy = on != 0; }
extern "C" int sim_last_cluster_size() { return sim::last_cluster_size; }
extern "C" void sim_set_sm_count(int n) { sim::sm_count = n; }
extern "C" void sim_set_schedule(unsigned mode) { sim::sched_mode = mode; }

namespace {
void racy_kernel(float* out) {        // thread t reads its neighbour's slot without a barrier after the write
    float* buf = reinterpret_cast<float*>(sim::smem());
    const int tid = threadIdx.x;
    buf[tid] = static_cast<float>(tid + 1);
    // missing __syncthreads()
    out[tid] = buf[(tid + 1) % blockDim.x];
    __syncthreads();
}
}  // namespace
extern "C" const char* sim_selftest_race(float* out) {
    return sim::launch(1, 64, 64 * 4, [&] { racy_kernel(out); });
}
