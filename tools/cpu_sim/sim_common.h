// CPU functional simulator for the fused SS2D kernels -- TEST INFRASTRUCTURE, host only (g++, no nvcc, no GPU).
//
// tools/cpu_sim/build_sim.py copies the kernel sources (csrc/ss2d_fused_{fwd,bwd}.cu, csrc/next/*_cg.cu) next to
// this header with the inline PTX swapped for the functions below and compiles them as ordinary C++: this file
// stands in for csrc/common.cuh.  Execution model: one CTA at a time, every CUDA thread a cooperative fiber
// (ucontext) that yields only inside __syncthreads(), warp shuffles and mbarrier waits, so runs are deterministic
// and data-race free by construction.  Asynchronous copies (cp.async, TMA bulk copies) are modelled at the two
// extremes of what the hardware may do: EAGER (default) -- the copy lands the moment it is issued, which exposes a
// buffer that is refilled while slower threads still read it; LAZY (sim_set_lazy(1)) -- the destination is poisoned
// at issue and the data lands only when somebody waits on the copy's mbarrier (global->shared), respectively the
// shared source is read only at the issuing thread's cp.async.bulk.wait_group (shared->global), which exposes a
// read before the wait and a staging buffer overwritten before its bulk store has drained.
//
// What it checks: the kernels' index arithmetic, work decomposition, scan / carry logic, barrier and mbarrier
// protocols (a barrier mismatch shows up as a reported deadlock) and numerics against the oracle.
// What it cannot check: interleavings between the two extremes, memory-model ordering (fences, proxies), bank
// conflicts, performance.
#pragma once
#include <cuda_bf16.h>
#include <driver_types.h>
#include <vector_types.h>
#include <vector_functions.h>
#include <stdint.h>
#include <setjmp.h>
#include <ucontext.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <unordered_map>
#include <vector>

#include "tramba_b200.h"

#define __launch_bounds__(...)

namespace sim {

struct Idx3 {
    unsigned x, y, z;
};
struct Mbar {
    uint32_t count = 0, pending = 0, phase = 0;
    long long tx = 0;
};
struct Warp {
    uint32_t slot[32];
    uint32_t frag[32][6];   // mma.sync operands published by the lanes
    int count = 0, gen = 0;
};
struct Copy {
    void* dst;
    const void* src;
    size_t bytes;
    bool add_f32 = false;   // cp.reduce.async.bulk ... .add.f32: dst[i] += src[i] instead of a copy
};
inline void apply_copy(const Copy& c) {
    if (!c.add_f32) {
        std::memcpy(c.dst, c.src, c.bytes);
        return;
    }
    float* d = static_cast<float*>(c.dst);
    const float* s = static_cast<const float*>(c.src);
    for (size_t i = 0; i < c.bytes / 4; ++i) d[i] += s[i];
}
struct Cluster;
struct Cta {
    int nthreads = 0, bx = 0, gdim = 0;
    int by = 0, bz = 0, gdim_y = 1, gdim_z = 1;
    Cluster* cluster = nullptr;
    int rank = 0;                                                      // rank of this CTA in its cluster
    std::vector<unsigned char> sshared;                                // static __shared__ arena
    size_t sshared_used = 0;
    std::unordered_map<int, size_t> sshared_at;                        // declaration (source line) -> offset
    std::vector<std::vector<Copy>> thread_cp;                          // cp.async.commit_group style copies: open group
    std::vector<std::vector<std::vector<Copy>>> thread_cp_groups;      // ... committed groups, oldest first
    std::unordered_map<const void*, std::vector<Copy>> bar_copies;     // lazy g2s copies, per mbarrier
    std::unordered_map<const void*, std::vector<long long>> bar_tx;    // ... and their complete_tx bytes (-1: an arrive)
    std::vector<std::vector<Copy>> thread_async;                       // cp.async issued, not yet bound to a barrier
    std::vector<std::vector<Copy>> thread_stores;                      // lazy s2g bulk stores of each thread (open group)
    std::vector<std::vector<std::vector<Copy>>> thread_groups;         // ... and its committed bulk async-groups, oldest first
    unsigned char* smem = nullptr;
    ucontext_t* sched_ptr = nullptr;               // ucontext only BOOTSTRAPS a fiber (first entry); every later switch is
    jmp_buf* sched_jb = nullptr;                   // _setjmp / _longjmp: swapcontext costs a sigprocmask system call each time
    struct Jb {
        jmp_buf b;
    };
    std::vector<Jb> jb;
    std::vector<char> started;
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    int cur = 0;
    int bar_count = 0, bar_gen = 0;
    std::vector<Warp> warps;
    std::unordered_map<const void*, Mbar> mbars;
    long long progress = 0;
    const char* error = nullptr;
};
struct Cluster {
    std::vector<Cta*> ctas;
    int count = 0, gen = 0, threads = 0;
};
inline Cta* g = nullptr;
inline std::function<void()>* g_body = nullptr;
inline bool lazy = false;

inline Idx3 tidx() { return Idx3{static_cast<unsigned>(g->cur), 0, 0}; }
inline Idx3 bidx() { return Idx3{static_cast<unsigned>(g->bx), static_cast<unsigned>(g->by), static_cast<unsigned>(g->bz)}; }
inline Idx3 gdim() {
    return Idx3{static_cast<unsigned>(g->gdim), static_cast<unsigned>(g->gdim_y), static_cast<unsigned>(g->gdim_z)};
}
inline Idx3 bdim() { return Idx3{static_cast<unsigned>(g->nthreads), 1, 1}; }
inline unsigned char* smem() { return g->smem; }

inline void yield() {
    Cta* me = g;   // the scheduler repoints g at whichever CTA it resumes next; it is restored before we continue
    if (_setjmp(me->jb[me->cur].b) == 0) _longjmp(*me->sched_jb, 1);
}

inline void syncthreads() {
    const int gen = g->bar_gen;
    if (++g->bar_count == g->nthreads) {
        g->bar_count = 0, ++g->bar_gen, ++g->progress;
    } else {
        while (g->bar_gen == gen) yield();
    }
}
inline void warp_barrier() {
    Warp& w = g->warps[g->cur >> 5];
    const int gen = w.gen;
    if (++w.count == 32) {
        w.count = 0, ++w.gen, ++g->progress;
    } else {
        while (w.gen == gen) yield();
    }
}
// every lane of the warp publishes v, then reads lane `src` (src < 0: keeps its own value)
inline uint32_t shfl(uint32_t v, int src) {
    Warp& w = g->warps[g->cur >> 5];
    w.slot[g->cur & 31] = v;
    warp_barrier();
    const uint32_t r = src >= 0 ? w.slot[src] : v;
    warp_barrier();
    return r;
}
inline uint32_t f2u(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

inline void cluster_sync() {                 // barrier.cluster.arrive + wait over every thread of the cluster
    Cluster& c = *g->cluster;
    const int gen = c.gen;
    if (++c.count == c.threads) {
        c.count = 0, ++c.gen, ++g->progress;
    } else {
        while (c.gen == gen) yield();
    }
}
// static `__shared__ T name[...]` declarations: one arena per CTA, keyed by the declaration's source line
inline void* sshared(int key, size_t bytes) {
    auto it = g->sshared_at.find(key);
    if (it == g->sshared_at.end()) {
        const size_t off = (g->sshared_used + 15) & ~size_t(15);
        if (off + bytes > g->sshared.size()) {
            g->error = "static shared memory arena exhausted";
            return g->sshared.data();
        }
        g->sshared_used = off + bytes;
        it = g->sshared_at.emplace(key, off).first;
    }
    return g->sshared.data() + it->second;
}
template <typename T>
inline T* map_rank(T* p, unsigned rank) {     // cluster.map_shared_rank: same offset in the peer CTA's shared memory
    unsigned char* q = reinterpret_cast<unsigned char*>(const_cast<typename std::remove_const<T>::type*>(p));
    Cta* peer = g->cluster->ctas[rank];
    if (q >= g->sshared.data() && q < g->sshared.data() + g->sshared.size())
        return reinterpret_cast<T*>(peer->sshared.data() + (q - g->sshared.data()));
    return reinterpret_cast<T*>(peer->smem + (q - g->smem));
}
// cp.async + commit_group / wait_group (per-thread groups)
inline void cp_async_g(void* dst, const void* src, size_t n) {
    if (!lazy) {
        std::memcpy(dst, src, n);
        return;
    }
    std::memset(dst, 0xCD, n);
    g->thread_cp[g->cur].push_back(Copy{dst, src, n});
}
inline void cp_commit() {
    auto& open = g->thread_cp[g->cur];
    g->thread_cp_groups[g->cur].push_back(open);
    open.clear();
}
inline void cp_wait(size_t keep) {
    auto& groups = g->thread_cp_groups[g->cur];
    while (groups.size() > keep) {
        for (const Copy& c : groups.front()) apply_copy(c);
        groups.erase(groups.begin());
    }
}
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32: C (16x8) += A (16x8) . B (8x8), operands truncated to TF32
inline void mma_m16n8k8_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    Warp& w = g->warps[g->cur >> 5];
    const int lane = g->cur & 31;
    for (int i = 0; i < 4; ++i) w.frag[lane][i] = a[i] & 0xffffe000u;
    for (int i = 0; i < 2; ++i) w.frag[lane][4 + i] = b[i] & 0xffffe000u;
    warp_barrier();
    const int gid = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; ++i) {
        const int row = gid + ((i & 2) ? 8 : 0), col = 2 * t + (i & 1);
        float acc = c[i];
        for (int k = 0; k < 8; ++k) {
            const float av = u2f(w.frag[(row & 7) * 4 + (k & 3)][(row >= 8 ? 1 : 0) + (k >= 4 ? 2 : 0)]);
            const float bv = u2f(w.frag[col * 4 + (k & 3)][4 + (k >= 4 ? 1 : 0)]);
            acc += av * bv;
        }
        c[i] = acc;
    }
    warp_barrier();
}

// ---- mbarrier (phase completes when the pending arrivals and the pending transaction bytes are both zero)
inline void mbar_check(Mbar& b) {
    if (b.pending == 0 && b.tx == 0) b.phase ^= 1u, b.pending = b.count, ++g->progress;
}
inline void mbar_arrive(const void* bar) {
    Mbar& b = g->mbars[bar];
    if (b.pending == 0) g->error = "mbarrier: more arrivals than the barrier's count in one phase";
    else --b.pending;
    mbar_check(b);
}
inline void mbar_complete_tx(const void* bar, long long bytes) {
    Mbar& b = g->mbars[bar];
    b.tx -= bytes;
    mbar_check(b);
}
inline void red_add_v2(float* p, float a, float b) { p[0] += a, p[1] += b; }

// ---- asynchronous copies
inline void cp_async(void* dst, const void* src, size_t n) {            // cp.async.ca.shared.global (4 / 8 bytes)
    if (!lazy) {
        std::memcpy(dst, src, n);
        return;
    }
    std::memset(dst, 0xCD, n);
    g->thread_async[g->cur].push_back(Copy{dst, src, n});
}
inline void cp_async_arrive(const void* bar) {                          // cp.async.mbarrier.arrive.noinc
    if (!lazy) {
        mbar_arrive(bar);
        return;
    }
    auto& mine = g->thread_async[g->cur];
    auto& q = g->bar_copies[bar];
    q.insert(q.end(), mine.begin(), mine.end());
    mine.clear();
    g->bar_tx[bar].push_back(-1);
}
inline void g2s(void* dst, const void* src, size_t bytes, const void* bar) {
    if (!lazy) {
        std::memcpy(dst, src, bytes);
        mbar_complete_tx(bar, static_cast<long long>(bytes));
        return;
    }
    std::memset(dst, 0xCD, bytes);
    g->bar_copies[bar].push_back(Copy{dst, src, bytes});
    g->bar_tx[bar].push_back(static_cast<long long>(bytes));
}
inline void flush_barrier(const void* bar) {                            // somebody waits: everything in flight lands now
    auto it = g->bar_copies.find(bar);
    if (it != g->bar_copies.end()) {
        for (const Copy& c : it->second) std::memcpy(c.dst, c.src, c.bytes);
        it->second.clear();
    }
    auto jt = g->bar_tx.find(bar);
    if (jt != g->bar_tx.end()) {
        std::vector<long long> evs;
        evs.swap(jt->second);
        for (long long e : evs) {
            if (e < 0) mbar_arrive(bar);
            else mbar_complete_tx(bar, e);
        }
    }
}
inline void s2g(void* dst, const void* src, size_t bytes, bool add_f32 = false) {
    const Copy c{dst, src, bytes, add_f32};
    if (!lazy) apply_copy(c);
    else g->thread_stores[g->cur].push_back(c);
}
inline void commit_stores() {                                           // cp.async.bulk.commit_group
    auto& open = g->thread_stores[g->cur];
    g->thread_groups[g->cur].push_back(open);
    open.clear();
}
inline void drain_stores(size_t keep) {                                 // cp.async.bulk.wait_group[.read] N
    auto& groups = g->thread_groups[g->cur];
    while (groups.size() > keep) {
        for (const Copy& c : groups.front()) apply_copy(c);
        groups.erase(groups.begin());
    }
}
inline void drain_all_stores() {                                        // thread exit: everything in flight completes
    commit_stores();
    drain_stores(0);
}

// Runs `body` as grid x nthreads CUDA threads, CTA after CTA.  Returns nullptr or an error string.
struct Dim3 {
    int x = 1, y = 1, z = 1;
};
inline int last_cluster_size = 0;   // CTAs per cluster of the most recent launch (tests check which path ran)
// Order in which the runnable fibers of a sweep are resumed: 0 = ascending thread index, 1 = descending, >= 2 = a
// fresh pseudo-random permutation per sweep (seeded with the value).  Results must not depend on it: two threads that
// touch the same shared-memory word between two barriers (a missing __syncthreads) give order-dependent results.
inline unsigned sched_mode = 0;
inline unsigned sched_state = 1;
inline unsigned sched_rand() {
    sched_state = sched_state * 1664525u + 1013904223u;
    return sched_state >> 8;
}
// Runs `body` as grid x nthreads CUDA threads.  CTAs run one cluster at a time (cluster = {1,1,1}: one CTA at a
// time); the CTAs of a cluster run concurrently, which cluster.sync() and distributed shared memory need.
inline const char* launch3(Dim3 grid, Dim3 cl, int nthreads, size_t smem_bytes, std::function<void()> body) {
    static std::vector<char*> stacks;
    const size_t kStack = 192 * 1024;
    const int csize = cl.x * cl.y * cl.z;
    last_cluster_size = csize;
    if (grid.x % cl.x || grid.y % cl.y || grid.z % cl.z) return "grid is not a multiple of the cluster shape";
    while (static_cast<int>(stacks.size()) < nthreads * csize) stacks.push_back(static_cast<char*>(std::malloc(kStack)));
    std::vector<void*> raws(csize, nullptr);
    for (void*& r : raws)
#ifdef SIM_ASAN   // exact size: AddressSanitizer's red zones sit right behind the dynamic shared memory
        if (posix_memalign(&r, 1024, smem_bytes ? smem_bytes : 16) != 0) return "out of memory";
#else
        if (posix_memalign(&r, 1024, smem_bytes + 1024) != 0) return "out of memory";
#endif
    const char* err = nullptr;
    g_body = &body;
    for (int oz = 0; oz < grid.z && !err; oz += cl.z)
        for (int oy = 0; oy < grid.y && !err; oy += cl.y)
            for (int ox = 0; ox < grid.x && !err; ox += cl.x) {
                static std::vector<Cta> ctas;                 // pooled: their per-thread vectors keep their capacity
                if (static_cast<int>(ctas.size()) < csize) ctas.resize(csize);
                Cluster cluster;
                cluster.threads = nthreads * csize;
                ucontext_t sched;
                jmp_buf sched_jb;
                int r = 0;
                for (int cz = 0; cz < cl.z; ++cz)
                    for (int cy = 0; cy < cl.y; ++cy)
                        for (int cx = 0; cx < cl.x; ++cx, ++r) {
                            Cta& cta = ctas[r];
                            cta.bar_count = cta.bar_gen = 0, cta.progress = 0, cta.error = nullptr, cta.cur = 0;
                            cta.mbars.clear(), cta.bar_copies.clear(), cta.bar_tx.clear();
                            cta.sshared_used = 0, cta.sshared_at.clear();
                            cta.nthreads = nthreads, cta.rank = r, cta.cluster = &cluster;
                            cta.bx = ox + cx, cta.by = oy + cy, cta.bz = oz + cz;
                            cta.gdim = grid.x, cta.gdim_y = grid.y, cta.gdim_z = grid.z;
                            cta.smem = static_cast<unsigned char*>(raws[r]);
                            std::memset(cta.smem, 0xCD, smem_bytes);   // poison: uninitialised shared memory must show
                            if (cta.sshared.size() != 48 * 1024) cta.sshared.resize(48 * 1024);
                            std::memset(cta.sshared.data(), 0xCD, 16 * 1024);   // the kernels' static arrays fit 16 KB
                            cta.ctx.resize(nthreads), cta.done.assign(nthreads, 0);
                            cta.jb.resize(nthreads), cta.started.assign(nthreads, 0);
                            cta.sched_jb = &sched_jb;
                            cta.warps.assign((nthreads + 31) / 32, Warp{});
                            for (auto* v : {&cta.thread_async, &cta.thread_stores, &cta.thread_cp}) {
                                v->resize(nthreads);
                                for (auto& q : *v) q.clear();
                            }
                            for (auto* v : {&cta.thread_groups, &cta.thread_cp_groups}) {
                                v->resize(nthreads);
                                for (auto& q : *v) q.clear();
                            }
                            cluster.ctas.push_back(&cta);
                            for (int t = 0; t < nthreads; ++t) {
                                getcontext(&cta.ctx[t]);
                                cta.ctx[t].uc_stack.ss_sp = stacks[r * nthreads + t];
                                cta.ctx[t].uc_stack.ss_size = kStack;
                                cta.ctx[t].uc_link = &sched;
                                makecontext(&cta.ctx[t], reinterpret_cast<void (*)()>(+[]() {
                                                (*g_body)();
                                                drain_all_stores();   // a store still in flight at exit completes eventually
                                                cp_wait(0);
                                                g->done[g->cur] = 1;
                                                ++g->progress;
                                                _longjmp(*g->sched_jb, 1);     // never falls off the fiber's stack
                                            }),
                                            0);
                            }
                        }
                int alive = 1;
                std::vector<int> order(nthreads);
                for (int t = 0; t < nthreads; ++t) order[t] = sched_mode == 1 ? nthreads - 1 - t : t;
                if (sched_mode >= 2) sched_state = sched_mode;
                while (alive > 0 && !err) {
                    long long before = 0;
                    for (int ci = 0; ci < csize; ++ci) before += ctas[ci].progress;
                    alive = 0;
                    if (sched_mode >= 2)
                        for (int t = nthreads - 1; t > 0; --t) std::swap(order[t], order[sched_rand() % (t + 1)]);
                    for (int ci = 0; ci < csize; ++ci) {
                        Cta& cta = ctas[ci];
                        g = &cta;
                        for (int oi = 0; oi < nthreads; ++oi) {
                            const int t = order[oi];
                            if (cta.done[t]) continue;
                            cta.cur = t;
                            cta.sched_ptr = &sched;
                            if (_setjmp(sched_jb) == 0) {
                                if (!cta.started[t]) {
                                    cta.started[t] = 1;
                                    setcontext(&cta.ctx[t]);              // first entry: onto the fiber's own stack
                                } else {
                                    _longjmp(cta.jb[t].b, 1);
                                }
                            }
                            if (!cta.done[t]) ++alive;
                        }
                        if (cta.error) err = cta.error;
                    }
                    long long after = 0;
                    for (int ci = 0; ci < csize; ++ci) after += ctas[ci].progress;
                    if (alive > 0 && after == before && !err) err = "deadlock: every live thread is blocked";
                }
                g = nullptr;
            }
    for (void* r : raws) std::free(r);
    return err;
}
inline const char* launch(int grid, int nthreads, size_t smem_bytes, std::function<void()> body) {
    return launch3(Dim3{grid, 1, 1}, Dim3{1, 1, 1}, nthreads, smem_bytes, std::move(body));
}

// ---- what the host launchers of a translated .cu file see instead of the CUDA runtime
inline int sm_count = 3;                 // cudaDevAttrMultiProcessorCount: small grids keep the simulation quick
inline const char* last_error = nullptr; // first simulator error of the current driver call
template <typename F>
inline void launch_kernel(int grid, int block, size_t smem_bytes, F&& body) {
    const char* e = launch(grid, block, smem_bytes, std::function<void()>(body));
    if (e && !last_error) last_error = e;
}
template <typename F>
inline void launch_kernel(dim3 grid, int block, size_t smem_bytes, F&& body) {
    const char* e = launch3(Dim3{static_cast<int>(grid.x), static_cast<int>(grid.y), static_cast<int>(grid.z)},
                            Dim3{1, 1, 1}, block, smem_bytes, std::function<void()>(body));
    if (e && !last_error) last_error = e;
}
inline void red_add_v4(float* p, float a, float b, float c, float d) { p[0] += a, p[1] += b, p[2] += c, p[3] += d; }

}  // namespace sim

template <typename K>
inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) {
    return cudaSuccess;
}
inline cudaError_t cudaGetDevice(int* dev) {
    *dev = 0;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr attr, int) {
    *v = attr == cudaDevAttrMultiProcessorCount ? sim::sm_count : attr == cudaDevAttrComputeCapabilityMajor ? 10 : 0;
    return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return sim::last_error ? cudaErrorLaunchFailure : cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
    std::memset(p, v, n);
    return cudaSuccess;
}
inline const char* cudaGetErrorString(cudaError_t) { return "simulated CUDA error"; }
template <typename... KA, typename... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kern)(KA...), A... args) {
    sim::Dim3 cl;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) {
            cl.x = static_cast<int>(cfg->attrs[i].val.clusterDim.x);
            cl.y = static_cast<int>(cfg->attrs[i].val.clusterDim.y);
            cl.z = static_cast<int>(cfg->attrs[i].val.clusterDim.z);
        }
    const char* e = sim::launch3(sim::Dim3{static_cast<int>(cfg->gridDim.x), static_cast<int>(cfg->gridDim.y),
                                           static_cast<int>(cfg->gridDim.z)},
                                 cl, static_cast<int>(cfg->blockDim.x), cfg->dynamicSmemBytes,
                                 [=] { kern(args...); });
    if (e && !sim::last_error) sim::last_error = e;
    return cudaSuccess;
}

namespace cooperative_groups {   // the slice of <cooperative_groups.h> the kernels use
struct cluster_group {
    void sync() const { sim::cluster_sync(); }
    template <typename T>
    T* map_shared_rank(T* p, unsigned rank) const {
        return sim::map_rank(p, rank);
    }
    unsigned block_rank() const { return static_cast<unsigned>(sim::g->rank); }
    unsigned num_blocks() const { return static_cast<unsigned>(sim::g->cluster->ctas.size()); }
};
inline cluster_group this_cluster() { return cluster_group{}; }
}  // namespace cooperative_groups
template <typename K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) {
    *n = 1;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------- CUDA surface
// threadIdx / blockIdx / blockDim / gridDim as objects (not macros: `cfg.gridDim` in host code must keep working)
namespace sim {
template <int WHICH, int AXIS>
struct BuiltinField {
    operator unsigned() const {
        const Idx3 v = WHICH == 0 ? tidx() : WHICH == 1 ? bidx() : WHICH == 2 ? bdim() : gdim();
        return AXIS == 0 ? v.x : AXIS == 1 ? v.y : v.z;
    }
};
template <int WHICH>
struct Builtin3 {
    BuiltinField<WHICH, 0> x;
    BuiltinField<WHICH, 1> y;
    BuiltinField<WHICH, 2> z;
};
}  // namespace sim
static const sim::Builtin3<0> threadIdx{};
static const sim::Builtin3<1> blockIdx{};
static const sim::Builtin3<2> blockDim{};
static const sim::Builtin3<3> gridDim{};
inline void __syncthreads() { sim::syncthreads(); }
inline float __shfl_up_sync(unsigned, float v, int delta, int width = 32) {
    const int lane = sim::g->cur & 31, sl = lane % width;
    return sim::u2f(sim::shfl(sim::f2u(v), sl >= delta ? lane - delta : -1));
}
inline float __shfl_down_sync(unsigned, float v, int delta, int width = 32) {
    const int lane = sim::g->cur & 31, sl = lane % width;
    return sim::u2f(sim::shfl(sim::f2u(v), sl + delta < width ? lane + delta : -1));
}
inline float __shfl_xor_sync(unsigned, float v, int mask, int width = 32) {
    const int lane = sim::g->cur & 31, src = lane ^ mask;
    return sim::u2f(sim::shfl(sim::f2u(v), (src / width) == (lane / width) ? src : -1));
}
inline float __shfl_sync(unsigned, float v, int src_lane, int width = 32) {
    const int lane = sim::g->cur & 31;
    return sim::u2f(sim::shfl(sim::f2u(v), (lane / width) * width + (src_lane % width)));
}
inline void __syncwarp(unsigned = 0xffffffffu) { sim::warp_barrier(); }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
using std::max;
using std::min;
template <typename T>
inline T __ldg(const T* p) {
    return *p;
}
inline float __int_as_float(int i) { return sim::u2f(static_cast<uint32_t>(i)); }
inline int __float_as_int(float f) { return static_cast<int>(sim::f2u(f)); }
inline float __uint_as_float(uint32_t u) { return sim::u2f(u); }
inline uint32_t __float_as_uint(float f) { return sim::f2u(f); }
inline float atomicAdd(float* p, float v) {
    const float old = *p;
    *p = old + v;
    return old;
}

namespace tramba {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

void note_launch();   // capi.cu (translated) defines it

// division by a run-time constant (csrc/common.cuh FastDiv; host arithmetic here)
struct FastDiv {
    uint32_t d, mul, shift;
    void init(uint32_t d_) {
        d = d_;
        if (d <= 1) {
            mul = 0;
            shift = 0;
            return;
        }
        shift = 0;
        while ((1ull << shift) < d) ++shift;
        mul = static_cast<uint32_t>((((1ull << shift) - d) << 32) / d + 1);
    }
    uint32_t div(uint32_t n) const {
        return static_cast<uint32_t>(((static_cast<uint64_t>(n) * mul >> 32) + n) >> shift);
    }
};

// per-device launch configuration helpers of csrc/common.cuh (one simulated device; sm_count may change per test)
constexpr int kMaxDevices = 64;
template <typename Kern>
inline cudaError_t ensure_dyn_smem(Kern, int) {
    return cudaSuccess;
}
inline int device_sms() { return sim::sm_count; }

// ---- mbarrier + TMA bulk: the same names as csrc/common.cuh
inline void mbar_init(uint64_t* bar, uint32_t count) {
    sim::Mbar& b = sim::g->mbars[bar];
    b.count = b.pending = count, b.phase = 0, b.tx = 0;
}
inline void fence_mbar_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    sim::g->mbars[bar].tx += bytes;
    sim::mbar_arrive(bar);
}
inline void mbar_arrive(uint64_t* bar) { sim::mbar_arrive(bar); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    if (sim::lazy) sim::flush_barrier(bar);
    return sim::g->mbars[bar].phase != parity;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) sim::yield();
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    if ((reinterpret_cast<uintptr_t>(dst_smem) | reinterpret_cast<uintptr_t>(src_gmem) | bytes) & 15)
        sim::g->error = "bulk_g2s: address or size not a multiple of 16 bytes";
    sim::g2s(dst_smem, src_gmem, bytes, bar);
}
inline void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    if ((reinterpret_cast<uintptr_t>(dst_gmem) | reinterpret_cast<uintptr_t>(src_smem) | bytes) & 15)
        sim::g->error = "bulk_s2g: address or size not a multiple of 16 bytes";
    sim::s2g(dst_gmem, src_smem, bytes);
}
inline void bulk_reduce_add_f32_s2g(float* dst_gmem, const float* src_smem, uint32_t bytes) {
    if ((reinterpret_cast<uintptr_t>(dst_gmem) | reinterpret_cast<uintptr_t>(src_smem) | bytes) & 15)
        sim::g->error = "bulk_reduce_add_f32_s2g: address or size not a multiple of 16 bytes";
    sim::s2g(dst_gmem, src_smem, bytes, true);
}
inline void bulk_commit() { sim::commit_stores(); }
template <int N>
inline void bulk_wait_read() {
    sim::drain_stores(N);
}
template <int N>
inline void bulk_wait_all() {
    sim::drain_stores(N);
}

inline void red_add_v4(float* p, float a, float b, float c, float d) { sim::red_add_v4(p, a, b, c, d); }

// ---- math (libm instead of the .approx MUFU ops: differences are far below the parity tolerance)
// packed fp32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100): two IEEE operations
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float ex2f(float x) { return exp2f(x); }
inline float lg2f(float x) { return log2f(x); }
inline float rcpf(float x) { return 1.0f / x; }
}  // namespace tramba
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
namespace tramba {
template <typename T>
struct DTypeOf;
template <>
struct DTypeOf<float> {
    static constexpr int value = TRAMBA_F32;
};
template <>
struct DTypeOf<__nv_bfloat16> {
    static constexpr int value = TRAMBA_BF16;
};
inline float softplus20(float x, float& e_out, float& w_out) {
    const float e = ex2f(x * kLog2e);
    const float w = 1.0f + e;
    const float corr = (e < 1.0f) ? (e - (w - 1.0f)) : 0.0f;
    const float sp = fmaf(lg2f(w), kLn2, corr);
    e_out = e, w_out = w;
    return (x > 20.0f) ? x : sp;
}
inline float softplus20(float x, float& e_out) {
    float w;
    return softplus20(x, e_out, w);
}
inline float softplus20_fast(float x) {
    const float sp = lg2f(1.0f + ex2f(x * kLog2e)) * kLn2;
    return (x > 20.0f) ? x : sp;
}
inline float to_f32(float v) { return v; }
inline float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
inline T from_f32(float v);
template <>
inline float from_f32<float>(float v) {
    return v;
}
template <>
inline __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}
template <int CNT, typename T>
inline void load_vec(const T* p, float (&v)[CNT]) {
    for (int i = 0; i < CNT; ++i) v[i] = to_f32(p[i]);
}
template <int CNT, typename T>
inline void store_vec(T* p, const float (&v)[CNT]) {
    for (int i = 0; i < CNT; ++i) p[i] = from_f32<T>(v[i]);
}

}  // namespace tramba
