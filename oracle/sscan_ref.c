/* CPU ORACLE (test infrastructure, NOT product code) -- selective-scan recurrence.
 *
 * PARITY UNPINNED against the reference's OWN implementation: the recurrence is not in /root/reference.  The
 * reference only calls an un-vendored third-party CUDA extension (MzeroMiko/VMamba kernels/selective_scan,
 * derived from state-spaces/mamba; no version pin anywhere in the reference, README.md:38) at
 *   Models/SS2D/csms6s.py:910      selective_scan_cuda_oflex.fwd(u,delta,A,B,C,D,delta_bias,softplus,1,oflex)
 *   Models/SS2D/csms6s.py:920-922  selective_scan_cuda_oflex.bwd(u,delta,A,B,C,D,delta_bias,dout,x,softplus,1)
 * This file restates the published algorithm of that dependency (the canonical
 * `selective_scan_ref`): fp32/fp64 internals, softplus with threshold 20, grouped B/C with
 * channel->group = d / (dim/groups), out = y + u*D.  What it IS pinned to (tests/test_oracle_sscan_pin.py,
 * fixture tests/golden/sscan_hf.npz from tests/golden/make_golden_sscan_hf.py): an independent, executed public
 * implementation of that algorithm -- Hugging Face transformers 5.5 `MambaMixer.slow_forward`, the pure-PyTorch
 * statement of state-spaces/mamba's selective scan -- forward to 1e-10 and all seven gradients against its
 * autograd in fp64; plus an independent torch-autograd statement of the same formulas (tests/test_oracle.py).
 * That is the strongest anchor available offline; the reference's own extension remains absent.
 *
 * Layouts (all contiguous, L innermost), as the reference call site passes them
 * (Models/vmamba.py:244-257):
 *   u, delta, out, dout, du, ddelta : (batch, dim, L)
 *   A : (dim, N)    D, delta_bias : (dim)    B, C, dB, dC : (batch, groups, N, L)
 *
 * Two instantiations: _f32 (float I/O, float arithmetic: the CPU baseline that is timed)
 * and _f64 (double I/O and arithmetic: the truth the CUDA kernels are compared with).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DEFINE_SSCAN(SUF, T, EXP, LOG1P)                                                         \
void sscan_fwd_##SUF(const T* u, const T* delta, const T* A, const T* Bm, const T* Cm,           \
                     const T* D, const T* dbias, T* out, int batch, int dim, int L, int N,       \
                     int groups, int softplus) {                                                 \
    const int dpg = dim / groups;                                                                \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
    for (int b = 0; b < batch; ++b)                                                              \
        for (int d = 0; d < dim; ++d) {                                                          \
            const int g = d / dpg;                                                               \
            const T* ur = u + ((size_t)b * dim + d) * L;                                         \
            const T* dr = delta + ((size_t)b * dim + d) * L;                                     \
            const T* Br = Bm + ((size_t)b * groups + g) * N * (size_t)L;                         \
            const T* Cr = Cm + ((size_t)b * groups + g) * N * (size_t)L;                         \
            T* yr = out + ((size_t)b * dim + d) * L;                                             \
            const T bias = dbias ? dbias[d] : (T)0;                                              \
            const T Dd = D ? D[d] : (T)0;                                                        \
            T h[64];                                                                             \
            for (int n = 0; n < N; ++n) h[n] = 0;                                                \
            for (int t = 0; t < L; ++t) {                                                        \
                T x = dr[t] + bias;                                                              \
                T dl = (softplus && x <= (T)20) ? LOG1P(EXP(x)) : x;                             \
                T y = Dd * ur[t];                                                                \
                for (int n = 0; n < N; ++n) {                                                    \
                    T a = EXP(dl * A[(size_t)d * N + n]);                                        \
                    h[n] = a * h[n] + dl * Br[(size_t)n * L + t] * ur[t];                        \
                    y += Cr[(size_t)n * L + t] * h[n];                                           \
                }                                                                                \
                yr[t] = y;                                                                       \
            }                                                                                    \
        }                                                                                        \
}                                                                                                \
                                                                                                 \
/* dA, dD, ddbias: (dim[,N]) summed over batch; dB, dC: (batch, groups, N, L) summed over the */ \
/* dim/groups channels of a group.  All outputs are overwritten.                              */ \
void sscan_bwd_##SUF(const T* u, const T* delta, const T* A, const T* Bm, const T* Cm,           \
                     const T* D, const T* dbias, const T* dout, T* du, T* ddelta, T* dA,         \
                     T* dB, T* dC, T* dD, T* ddbias, int batch, int dim, int L, int N,           \
                     int groups, int softplus) {                                                 \
    const int dpg = dim / groups;                                                                \
    T* pA = (T*)calloc((size_t)batch * dim * N, sizeof(T));                                      \
    T* pD = (T*)calloc((size_t)batch * dim, sizeof(T));                                          \
    T* pb = (T*)calloc((size_t)batch * dim, sizeof(T));                                          \
    memset(dB, 0, sizeof(T) * (size_t)batch * groups * N * L);                                   \
    memset(dC, 0, sizeof(T) * (size_t)batch * groups * N * L);                                   \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                                    \
    for (int b = 0; b < batch; ++b)                                                              \
        for (int g = 0; g < groups; ++g) {                                                       \
            T* hist = (T*)malloc(sizeof(T) * (size_t)(L + 1) * N);                               \
            T* dl = (T*)malloc(sizeof(T) * (size_t)L);                                           \
            const T* Br = Bm + ((size_t)b * groups + g) * N * (size_t)L;                         \
            const T* Cr = Cm + ((size_t)b * groups + g) * N * (size_t)L;                         \
            T* dBr = dB + ((size_t)b * groups + g) * N * (size_t)L;                              \
            T* dCr = dC + ((size_t)b * groups + g) * N * (size_t)L;                              \
            for (int d = g * dpg; d < (g + 1) * dpg; ++d) {                                      \
                const T* ur = u + ((size_t)b * dim + d) * L;                                     \
                const T* dr = delta + ((size_t)b * dim + d) * L;                                 \
                const T* gr = dout + ((size_t)b * dim + d) * L;                                  \
                T* dur = du + ((size_t)b * dim + d) * L;                                         \
                T* ddr = ddelta + ((size_t)b * dim + d) * L;                                     \
                const T bias = dbias ? dbias[d] : (T)0;                                          \
                const T Dd = D ? D[d] : (T)0;                                                    \
                /* forward replay: hist[(t+1)*N+n] = h_t[n], hist[0..N) = h_{-1} = 0 */          \
                for (int n = 0; n < N; ++n) hist[n] = 0;                                         \
                for (int t = 0; t < L; ++t) {                                                    \
                    T x = dr[t] + bias;                                                          \
                    dl[t] = (softplus && x <= (T)20) ? LOG1P(EXP(x)) : x;                        \
                    for (int n = 0; n < N; ++n) {                                                \
                        T a = EXP(dl[t] * A[(size_t)d * N + n]);                                 \
                        hist[(size_t)(t + 1) * N + n] =                                          \
                            a * hist[(size_t)t * N + n] + dl[t] * Br[(size_t)n * L + t] * ur[t]; \
                    }                                                                            \
                }                                                                                \
                T carry[64];                                                                     \
                for (int n = 0; n < N; ++n) carry[n] = 0;                                        \
                T sD = 0, sb = 0;                                                                \
                for (int t = L - 1; t >= 0; --t) {                                               \
                    T x = dr[t] + bias;                                                          \
                    T go = gr[t];                                                                \
                    T dd = 0, duv = Dd * go;                                                     \
                    for (int n = 0; n < N; ++n) {                                                \
                        T An = A[(size_t)d * N + n];                                             \
                        T a = EXP(dl[t] * An);                                                   \
                        T hp = hist[(size_t)t * N + n];                                          \
                        T hc = hist[(size_t)(t + 1) * N + n];                                    \
                        T Bt = Br[(size_t)n * L + t];                                            \
                        T dh = Cr[(size_t)n * L + t] * go + carry[n];                            \
                        duv += dh * dl[t] * Bt;                                                  \
                        dd += dh * (An * a * hp + Bt * ur[t]);                                   \
                        pA[((size_t)b * dim + d) * N + n] += dh * dl[t] * a * hp;                \
                        dBr[(size_t)n * L + t] += dh * dl[t] * ur[t];                            \
                        dCr[(size_t)n * L + t] += go * hc;                                       \
                        carry[n] = a * dh;                                                       \
                    }                                                                            \
                    T sg = (T)1;                                                                 \
                    if (softplus && x <= (T)20) sg = (T)1 / ((T)1 + EXP(-x));                    \
                    dur[t] = duv;                                                                \
                    ddr[t] = dd * sg;                                                            \
                    sD += go * ur[t];                                                            \
                    sb += dd * sg;                                                               \
                }                                                                                \
                pD[(size_t)b * dim + d] = sD;                                                    \
                pb[(size_t)b * dim + d] = sb;                                                    \
            }                                                                                    \
            free(hist);                                                                          \
            free(dl);                                                                            \
        }                                                                                        \
    for (int d = 0; d < dim; ++d) {                                                              \
        T sD = 0, sb = 0;                                                                        \
        for (int b = 0; b < batch; ++b) {                                                        \
            sD += pD[(size_t)b * dim + d];                                                       \
            sb += pb[(size_t)b * dim + d];                                                       \
        }                                                                                        \
        if (dD) dD[d] = sD;                                                                      \
        if (ddbias) ddbias[d] = sb;                                                              \
        for (int n = 0; n < N; ++n) {                                                            \
            T sA = 0;                                                                            \
            for (int b = 0; b < batch; ++b) sA += pA[((size_t)b * dim + d) * N + n];             \
            dA[(size_t)d * N + n] = sA;                                                          \
        }                                                                                        \
    }                                                                                            \
    free(pA);                                                                                    \
    free(pD);                                                                                    \
    free(pb);                                                                                    \
}

DEFINE_SSCAN(f32, float, expf, log1pf)
DEFINE_SSCAN(f64, double, exp, log1p)

int sscan_ref_max_dstate(void) { return 64; }
