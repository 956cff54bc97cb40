#!/usr/bin/env python
"""Builds tools/cpu_sim/_gen/libtramba_sim.so: EVERY source file of libtramba_b200 (capi.cu, orders.cpp, all kernel .cu
files) compiled as HOST code against tools/cpu_sim/sim_common.h (see there), so
that the product's C ABI and kernels can be executed on a CPU.

The sources are copied, not edited in place; the only textual changes are (1) inline PTX -> the simulator's functions,
(2) `extern __shared__ ... name[]` / static `__shared__` arrays -> the simulator's per-CTA buffers, (3) `<<<...>>>`
launches -> sim::launch_kernel, (4) SM counts cached in function-local statics made per call, (5) small `extern "C"`
drivers appended for the tests that call kernels directly.  TEST INFRASTRUCTURE (g++ only, no nvcc, no GPU).

    python tools/cpu_sim/build_sim.py        # -> path of the library
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "tramba_b200", "csrc")
GEN = os.path.join(HERE, "_gen")
LIB = os.path.join(GEN, "libtramba_sim.so")


def sub1(s, old, new):
    assert s.count(old) == 1, (s.count(old), old[:80])
    return s.replace(old, new)


def gen_header():
    s = open(os.path.join(CSRC, "ss2d_fused.cuh")).read()
    s = sub1(s, '#include "common.cuh"', '#include "sim_common.h"')
    s = sub1(s, 'asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");',
             "sim::cp_async(dst_smem, src, 4);")
    s = sub1(s, 'asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");',
             "sim::cp_async(dst_smem, src, 8);")
    s = sub1(s, 'asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");',
             "sim::cp_async_arrive(bar);")
    # explicit shared-window accessors -> plain host accesses at (base + byte offset)
    s = sub1(s, 'asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(smem_u32(base) + off) : "memory");',
             "v = *reinterpret_cast<const uint16_t*>(static_cast<const unsigned char*>(base) + off);")
    s = sub1(s, 'asm volatile("st.shared.u16 [%0], %1;" ::"r"(smem_u32(base) + off), "h"(static_cast<uint16_t>(v)) : "memory");',
             "*reinterpret_cast<uint16_t*>(static_cast<unsigned char*>(base) + off) = static_cast<uint16_t>(v);")
    s = sub1(s, 'asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(base) + off) : "memory");',
             "v = *reinterpret_cast<const float*>(static_cast<const unsigned char*>(base) + off);")
    s = sub1(s, 'asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(base) + off), "f"(v) : "memory");',
             "*reinterpret_cast<float*>(static_cast<unsigned char*>(base) + off) = v;")
    assert "asm" not in s
    open(os.path.join(GEN, "ss2d_fused.cuh"), "w").write(s)


RED = re.compile(r'asm volatile\("red\.global\.add\.v2\.f32 \[%0\], \{%1, %2\};" ::"l"\((.*?)\), "f"\((.*?)\), '
                 r'"f"\((.*?)\)\s*: "memory"\);', re.S)

FUSED_DRIVER = r'''
// Drivers of the fused SS2D core launchers (csrc/ss2d_fused_fwd.cu / _bwd.cu) for tests that call the kernels without
// the C ABI.  `grid` plays the role of the SM count: the launchers size their persistent grids from it.
#include "ss2d_fused.cuh"
using namespace tramba;
static const char* done(cudaError_t e) {
    return sim::last_error ? sim::last_error : (e == cudaSuccess ? nullptr : "launcher rejected the arguments");
}
static FusedArgs with_depth(const FusedArgs& a, int in_dtype) {   // the staging depth tramba_ss2d_fused_fwd picks
    FusedArgs b = a;
    const int es = in_dtype == TRAMBA_F32 ? 4 : 2;
    b.xb = ss2d_fused_fwd_smem(a.L, es, 2) <= 227 * 1024 ? 2 : 1;
    return b;
}
extern "C" const char* sim_fwd(const FusedArgs* a, int in_dtype, int grid) {
    sim::sm_count = grid, sim::last_error = nullptr;
    return done(launch_ss2d_fused_fwd(with_depth(*a, in_dtype), in_dtype, nullptr));
}
'''

FUSED_FILES = ["ss2d_fused_fwd.cu", "ss2d_fused_bwd.cu"]


def gen_fused():
    drv = os.path.join(GEN, "fused_driver_sim.cpp")
    open(drv, "w").write(FUSED_DRIVER)
    return [gen_whole_file(f) for f in FUSED_FILES] + [drv]


LAUNCH = re.compile(r"([\w:]+(?:<[^<>;]*>)?)<<<(.+?), (.+?), (.+?), st>>>\((.*?)\);", re.S)
SHARED = re.compile(r"extern __shared__ (?:__align__\(\d+\) )?(\w+(?: \w+)?) (\w+)\[\];")
RED4 = re.compile(r'asm volatile\("red\.global\.add\.v4\.f32 \[%0\], \{%1, %2, %3, %4\};" ::"l"\((.*?)\), "f"\((.*?)\),\s*'
                  r'"f"\((.*?)\), "f"\((.*?)\), "f"\((.*?)\)\s*: "memory"\);', re.S)

SSCAN_DRIVER = r'''
// C-ABI-shaped drivers of the operator selective-scan kernels: the argument / plan set-up of tramba_sscan_fwd / _bwd
// (tramba_b200/csrc/capi.cu) followed by the translated launchers.
#include "sscan.cuh"
using namespace tramba;
static int es_of(int dt) { return dt == TRAMBA_F32 ? 4 : 2; }
static void fill_args(SScanArgs& a, const SScanPlan& p, int batch, int dim, int L, int N, int groups, int softplus) {
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
extern "C" size_t sim_sscan_ckpt_elems(int batch, int dim, int L, int N, int in_dtype) {
    const SScanPlan p = make_sscan_plan(static_cast<long long>(batch) * dim, L, es_of(in_dtype));
    return static_cast<size_t>(p.units) * p.cpu * N;
}
extern "C" const char* sim_sscan_fwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                                     const float* D, const float* bias, void* out, float* ckpt, int batch, int dim,
                                     int L, int N, int groups, int in_dtype, int out_dtype, int softplus, int sms) {
    sim::sm_count = sms, sim::last_error = nullptr;
    const SScanPlan p = make_sscan_plan(static_cast<long long>(batch) * dim, L, es_of(in_dtype));
    SScanArgs a;
    std::memset(&a, 0, sizeof(a));
    a.u = u, a.delta = delta, a.A = A, a.Bm = Bm, a.Cm = Cm, a.D = D, a.bias = bias, a.out = out, a.ckpt = ckpt;
    fill_args(a, p, batch, dim, L, N, groups, softplus);
    const cudaError_t e = launch_sscan_fwd(a, p, in_dtype, out_dtype, nullptr);
    return sim::last_error ? sim::last_error : (e == cudaSuccess ? nullptr : "launcher rejected the arguments");
}
extern "C" const char* sim_sscan_bwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                                     const float* D, const float* bias, const void* dout, const float* ckpt, void* du,
                                     void* ddelta, float* dA, float* dB, float* dC, float* dD, float* dbias, int batch,
                                     int dim, int L, int N, int groups, int in_dtype, int dout_dtype, int softplus,
                                     int sms) {
    sim::sm_count = sms, sim::last_error = nullptr;
    const SScanPlan p = make_sscan_plan(static_cast<long long>(batch) * dim, L, es_of(in_dtype));
    const size_t nbc = static_cast<size_t>(batch) * groups * N * L * sizeof(float);
    std::memset(dB, 0, nbc), std::memset(dC, 0, nbc), std::memset(dA, 0, static_cast<size_t>(dim) * N * sizeof(float));
    if (dD) std::memset(dD, 0, static_cast<size_t>(dim) * sizeof(float));
    if (dbias) std::memset(dbias, 0, static_cast<size_t>(dim) * sizeof(float));
    SScanArgs a;
    std::memset(&a, 0, sizeof(a));
    a.u = u, a.delta = delta, a.A = A, a.Bm = Bm, a.Cm = Cm, a.D = D, a.bias = bias;
    a.dout = dout, a.ckpt_in = ckpt, a.du = du, a.ddelta = ddelta;
    a.dA = dA, a.dB = dB, a.dC = dC, a.dD = dD, a.dbias = dbias;
    fill_args(a, p, batch, dim, L, N, groups, softplus);
    if (p.regular) {                      // dB / dC pre-reduction over RW rows, as capi.cu chooses it
        if (N == 1) {
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
    const cudaError_t e = launch_sscan_bwd(a, p, in_dtype, dout_dtype, nullptr);
    return sim::last_error ? sim::last_error : (e == cudaSuccess ? nullptr : "launcher rejected the arguments");
}
'''


MACRO_LAUNCH = re.compile(r"([\w:]+<[^;\n]*?>)<<<(\w[\w.]*), ([\w.]+), ([\w.]+), st>>>\(([^\n]*)\)$", re.M)
STATIC_SHARED = re.compile(r"^(\s*)__shared__ (\w+) (\w+)((?:\[[^\]]+\])+);", re.M)
MMA = re.compile(r'asm volatile\(\s*"mma\.sync\.aligned\.m16n8k8\.row\.col\.f32\.tf32\.tf32\.f32.*?\);', re.S)


def gen_whole_file(rel):
    """A .cu file translated as a whole: kernels AND their host launchers (plan selection, template dispatch);
    only the <<<...>>> launches, the shared-memory declaration and inline PTX are rewritten."""
    s = open(os.path.join(CSRC, rel)).read()
    s = SHARED.sub(r"\1* \2 = reinterpret_cast<\1*>(sim::smem());", s)
    s = STATIC_SHARED.sub(r"\1typedef \2 \3_sim_t\4; \3_sim_t& \3 = *reinterpret_cast<\3_sim_t*>("
                          r"sim::sshared(__LINE__, sizeof(\3_sim_t)));", s)
    s = s.replace("#include <cooperative_groups.h>", "")
    s = MMA.sub("sim::mma_m16n8k8_tf32(c, a, b);", s)
    s = s.replace('asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");',
                  "sim::cp_async_g(dst, src, 16);")
    s = s.replace('asm volatile("cp.async.commit_group;" ::: "memory");', "sim::cp_commit();")
    s = s.replace('asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");', "sim::cp_wait(N);")
    s = RED.sub(r"sim::red_add_v2(\1, \2, \3);", s)
    s = re.sub(r"static int sms = 0(, occ = 0)?;", r"int sms = 0\1;", s)     # the SM count is per call in the simulator
    s = MACRO_LAUNCH.sub(r"sim::launch_kernel(\2, \3, \4, [&] { \1(\5); })", s)
    s = LAUNCH.sub(r"sim::launch_kernel(\2, \3, \4, [&] { \1(\5); });", s)
    s = RED4.sub(r"sim::red_add_v4(\1, \2, \3, \4, \5);", s)
    assert "<<<" not in s and "__shared__" not in s and "asm" not in s.replace("__nv_bfloat16", ""), rel
    out = os.path.join(GEN, os.path.basename(rel).replace(".cu", "_sim.cpp"))
    open(out, "w").write(s)
    return out


SCAN_DRIVER = r'''
// C-ABI-shaped drivers of the table-driven cross-scan / cross-merge kernels (csrc/cross_scan.cu): what
// tramba_cross_scan / tramba_cross_merge do after argument validation.
#include "sim_common.h"
namespace tramba {
cudaError_t launch_cross_scan(const void* x, void* xs, const int32_t* plan, int K, int B, int C, int L, int dtype,
                              cudaStream_t st);
cudaError_t launch_cross_merge(const void* ys, void* y, const int32_t* plan, int order, int K, int B, int D, int L,
                               int dtype, cudaStream_t st);
}
extern "C" const char* sim_cross_scan(const void* x, void* xs, const int32_t* plan, int order, int K, int B, int C, int H,
                                      int W, int dtype) {
    sim::last_error = nullptr;
    const cudaError_t e = tramba::launch_cross_scan(x, xs, plan, K, B, C, H * W, dtype, nullptr);
    return sim::last_error ? sim::last_error : (e == cudaSuccess ? nullptr : "launcher rejected the arguments");
}
extern "C" const char* sim_cross_merge(const void* ys, void* y, const int32_t* plan, int order, int K, int B, int D, int H,
                                       int W, int dtype) {
    sim::last_error = nullptr;
    const cudaError_t e = tramba::launch_cross_merge(ys, y, plan, order, K, B, D, H * W, dtype, nullptr);
    return sim::last_error ? sim::last_error : (e == cudaSuccess ? nullptr : "launcher rejected the arguments");
}
'''


def gen_scan():
    src = gen_whole_file("cross_scan.cu")
    t = open(src).read()
    t = sub1(t, '#include "common.cuh"', '#include "sim_common.h"')
    open(src, "w").write(t)
    drv = os.path.join(GEN, "scan_driver_sim.cpp")
    open(drv, "w").write(SCAN_DRIVER)
    return [src, drv]


SIDE_DRIVER = r'''
// C-ABI-shaped drivers of the depthwise-convolution and projected-channel kernels (csrc/dwconv2d.cu, proj_gather.cu).
#include "sim_common.h"
namespace tramba {
cudaError_t launch_dwconv2d_fwd(const void* x, const float* w, const float* b, void* y, int B, int C, int H, int W,
                                int K, int act, int dtype, cudaStream_t st);
cudaError_t launch_dwconv2d_bwd(const void* x, const float* w, const float* b, const void* dy, void* dx, float* dwp,
                                int B, int C, int H, int W, int K, int act, int dtype, cudaStream_t st);
cudaError_t launch_proj_gather(const void* src, void* dts, void* Bs, void* Cs, const int32_t* plan, int K, int B, int R,
                               int N, int L, int dtype, cudaStream_t st);
cudaError_t launch_proj_scatter(const void* gdts, const void* gBs, const void* gCs, void* gsrc, const int32_t* plan,
                                int K, int B, int R, int N, int L, int dtype, int bc_dtype, cudaStream_t st);
}
namespace tramba {
int layernorm2d_tiles(int B, int C, int L);
cudaError_t launch_layernorm2d_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd,
                                   int B, int C, int L, float eps, int act, int in_dtype, int out_dtype, cudaStream_t st);
cudaError_t launch_layernorm2d_bwd(const void* gy, const void* x, const float* w, const float* b, const float* mean,
                                   const float* rstd, void* gx, float* gw_part, float* gb_part, int B, int C, int L,
                                   int act, int in_dtype, int g_dtype, cudaStream_t st);
bool dct2d_tc_supported(int H, int W);
cudaError_t launch_dct2d_fwd(const float* x, const float* wy, const float* wx, float* high, float* low, int planes, int H,
                             int W, cudaStream_t st);
cudaError_t launch_dct2d_bwd(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                             int planes, int H, int W, cudaStream_t st);
cudaError_t launch_dct2d_fwd_tc(const float* x, const float* wy, const float* wx, float* high, float* low, int planes,
                                int H, int W, cudaStream_t st);
cudaError_t launch_dct2d_bwd_tc(const float* ghigh, const float* glow, const float* wy, const float* wx, float* gx,
                                int planes, int H, int W, cudaStream_t st);
}
static const char* done(cudaError_t e) {
    return sim::last_error ? sim::last_error : (e == cudaSuccess ? nullptr : "launcher rejected the arguments");
}
extern "C" const char* sim_dwconv2d_fwd(const void* x, const float* w, const float* b, void* y, int B, int C, int H, int W,
                                        int K, int act, int dtype) {
    sim::last_error = nullptr;
    return done(tramba::launch_dwconv2d_fwd(x, w, b, y, B, C, H, W, K, act, dtype, nullptr));
}
extern "C" const char* sim_dwconv2d_bwd(const void* x, const float* w, const float* b, const void* dy, void* dx,
                                        float* dwp, int B, int C, int H, int W, int K, int act, int dtype) {
    sim::last_error = nullptr;
    return done(tramba::launch_dwconv2d_bwd(x, w, b, dy, dx, dwp, B, C, H, W, K, act, dtype, nullptr));
}
extern "C" int sim_layernorm2d_tiles(int B, int C, int L, int sms) {
    sim::sm_count = sms;
    return tramba::layernorm2d_tiles(B, C, L);
}
extern "C" const char* sim_layernorm2d_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd,
                                           int B, int C, int L, float eps, int act, int in_dtype, int out_dtype, int sms) {
    sim::sm_count = sms, sim::last_error = nullptr;
    return done(tramba::launch_layernorm2d_fwd(x, w, b, y, mean, rstd, B, C, L, eps, act, in_dtype, out_dtype, nullptr));
}
extern "C" const char* sim_layernorm2d_bwd(const void* gy, const void* x, const float* w, const float* b, const float* mean,
                                           const float* rstd, void* gx, float* gw_part, float* gb_part, int B, int C, int L,
                                           int act, int in_dtype, int g_dtype, int sms) {
    sim::sm_count = sms, sim::last_error = nullptr;
    return done(tramba::launch_layernorm2d_bwd(gy, x, w, b, mean, rstd, gx, gw_part, gb_part, B, C, L, act, in_dtype, g_dtype,
                                               nullptr));
}
extern "C" const char* sim_dct2d(int bwd, int tensor_pipe, const float* a, const float* b2, const float* wy, const float* wx,
                                 float* o1, float* o2, int planes, int H, int W) {
    sim::last_error = nullptr;
    if (tensor_pipe && !tramba::dct2d_tc_supported(H, W)) return "shape not covered by the tensor-pipe kernel";
    if (!bwd) return done(tensor_pipe ? tramba::launch_dct2d_fwd_tc(a, wy, wx, o1, o2, planes, H, W, nullptr)
                                      : tramba::launch_dct2d_fwd(a, wy, wx, o1, o2, planes, H, W, nullptr));
    return done(tensor_pipe ? tramba::launch_dct2d_bwd_tc(a, b2, wy, wx, o1, planes, H, W, nullptr)
                            : tramba::launch_dct2d_bwd(a, b2, wy, wx, o1, planes, H, W, nullptr));
}
extern "C" const char* sim_proj_gather(const void* src, void* dts, void* Bs, void* Cs, const int32_t* plan, int K, int B,
                                       int R, int N, int L, int dtype) {
    sim::last_error = nullptr;
    return done(tramba::launch_proj_gather(src, dts, Bs, Cs, plan, K, B, R, N, L, dtype, nullptr));
}
extern "C" const char* sim_proj_scatter(const void* gdts, const void* gBs, const void* gCs, void* gsrc,
                                        const int32_t* plan, int K, int B, int R, int N, int L, int dtype, int bc_dtype) {
    sim::last_error = nullptr;
    return done(tramba::launch_proj_scatter(gdts, gBs, gCs, gsrc, plan, K, B, R, N, L, dtype, bc_dtype, nullptr));
}
'''


def gen_side():
    out = []
    for rel in ("dwconv2d.cu", "proj_gather.cu", "layernorm2d.cu", "dct2d.cu", "dct2d_tc.cu"):
        src = gen_whole_file(rel)
        t = open(src).read()
        t = sub1(t, '#include "common.cuh"', '#include "sim_common.h"')
        open(src, "w").write(t)
        out.append(src)
    drv = os.path.join(GEN, "side_driver_sim.cpp")
    open(drv, "w").write(SIDE_DRIVER)
    return out + [drv]


def gen_abi():
    """capi.cu (argument validation, plan selection: the real extern "C" entry points) and orders.cpp (host tables):
    with them the simulator library exports the product's whole C ABI, computed by the product's kernels."""
    src = gen_whole_file("capi.cu")
    t = open(src).read()
    t = sub1(t, '#include "common.cuh"', '#include "sim_common.h"')
    open(src, "w").write(t)
    o = open(os.path.join(CSRC, "orders.cpp")).read()
    o = sub1(o, '#include "../../include/tramba_b200.h"', '#include "tramba_b200.h"')
    dst = os.path.join(GEN, "orders_sim.cpp")
    open(dst, "w").write(o)
    return [src, dst]


def gen_sscan():
    h = open(os.path.join(CSRC, "sscan.cuh")).read()
    h = sub1(h, '#include "common.cuh"', '#include "sim_common.h"')
    open(os.path.join(GEN, "sscan.cuh"), "w").write(h)
    drv = os.path.join(GEN, "sscan_driver_sim.cpp")
    open(drv, "w").write(SSCAN_DRIVER)
    return [gen_whole_file("sscan_fwd.cu"), gen_whole_file("sscan_bwd.cu"), drv]


def build(force=False, asan=False):
    global GEN, LIB
    if asan:     # AddressSanitizer build: out-of-bounds global / shared memory accesses of the kernels abort the run
        GEN = os.path.join(HERE, "_gen_asan")
        LIB = os.path.join(GEN, "libtramba_sim_asan.so")
    else:
        GEN = os.path.join(HERE, "_gen")
        LIB = os.path.join(GEN, "libtramba_sim.so")
    os.makedirs(GEN, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in FUSED_FILES] + [os.path.join(CSRC, f) for f in ("capi.cu", "orders.cpp","sscan.cuh", "sscan_fwd.cu",
                                                                                         "sscan_bwd.cu", "cross_scan.cu", "dwconv2d.cu",
                                                                                         "proj_gather.cu", "layernorm2d.cu",
                                                                                         "dct2d.cu", "dct2d_tc.cu")] + [
        os.path.join(CSRC, "ss2d_fused.cuh"),
                                                          os.path.join(HERE, "sim_common.h"), os.path.join(HERE, "selftest.cpp"),
                                                          os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    gen_header()
    srcs = gen_fused() + gen_sscan() + gen_scan() + gen_side() + gen_abi()
    srcs.append(os.path.join(HERE, "selftest.cpp"))
    flags = ["-std=c++17", "-O1", "-fPIC", "-w", "-U_FORTIFY_SOURCE", "-D_FORTIFY_SOURCE=0",   # _longjmp across fiber stacks
             "-I" + HERE, "-I" + GEN, "-I" + os.path.join(ROOT, "include"),
             "-I/usr/local/cuda/include"] + (
                 ["-fsanitize=address", "-fno-omit-frame-pointer", "-g", "-DSIM_ASAN"] if asan else [])
    objs = [os.path.join(GEN, os.path.splitext(os.path.basename(f))[0] + ".o") for f in srcs]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda so: subprocess.run(["g++"] + flags + ["-c", so[0], "-o", so[1]],
                                                        capture_output=True, text=True), zip(srcs, objs)))
    for res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("cpu_sim build failed")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + (["-fsanitize=address"] if asan else []))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
