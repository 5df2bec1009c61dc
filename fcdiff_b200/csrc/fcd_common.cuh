// Shared device/host helpers for the fcdiff_b200 kernels (sm_100a).
//
// Reference formulas are cited as fcdiff/<file>:<line> (reference tree).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "fcdiff_b200.h"
#include "fcd_math.cuh"

namespace fcd {

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
int  check_launch(const char* what);          // counts the launch, maps cudaGetLastError
int  sm_count();

#define FCD_REQUIRE(cond, ...)                                                   \
    do {                                                                         \
        if (!(cond)) {                                                           \
            ::fcd::set_error(__VA_ARGS__);                                       \
            return -1;                                                           \
        }                                                                        \
    } while (0)

constexpr double kHalfLog2Pi = 0.91893853320467274178032973640562;   // log(sqrt(2 pi))

// Workspace layout (doubles): [0, kWsPartials) per-CTA partial sums,
// then one 64-bit slot holding the arrival ticket, then 16 scratch doubles.
constexpr int     kMaxReduceBlocks = 2048;
constexpr int     kMaxReduceVals   = 8;
constexpr int64_t kWsPartials      = (int64_t)kMaxReduceBlocks * kMaxReduceVals;
constexpr int64_t kWsScratch       = kWsPartials + 2;      // 16 doubles of scratch for chained reductions
constexpr int64_t kWsDoubles       = kWsScratch + 16;

// Derived per-theta constants, passed by value to kernels (__grid_constant__).
struct ThetaDev {
    double mu[3];
    double isig[3];      // 1 / sigma_k
    double lc[3];        // -log(sigma_k)            (the -log sqrt(2 pi) is added where needed)
    double epsl[3];      // eps_l of fcdiff/fit.py:433-444
    double al[3];        // (1 - eps_l) / 2
    double bl[3];        // eps_l - a_l
    double log_gamma[3];
    double hq_a[3];      // healthy quadratic: sum_h logN_k(b) = hq_a*S2 + hq_b*S1 + hq_c
    double hq_b[3];
    double hq_c[3];
    double eta;
    double epsilon;
    double log_pi2[2];
};

ThetaDev make_theta_dev(const fcd_theta& th, int H);

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

// util.c_to_nm (fcdiff/util.py:82-84) with an exact integer correction of the
// floating-point square root.
__host__ __device__ inline void c_to_nm(int64_t c, int& n, int& m) {
    int64_t nn = (int64_t)((sqrt((double)(8 * c + 1)) - 1.0) * 0.5) + 1;
    while (nn * (nn - 1) / 2 > c) --nn;
    while ((nn + 1) * nn / 2 <= c) ++nn;
    n = (int)nn;
    m = (int)(c - nn * (nn - 1) / 2);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming load of two doubles (read-only path, no L1 allocation).
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream1(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// Per-element core shared by K2 / K2b / K3 / K4.
//
// For x = bt[c,u]:   t_k = -((x-mu_k)/sigma_k)^2/2 - log sigma_k,
//                    e_k = exp(t_k - max_j t_j),  o_k = sum_{j != k} e_j,
//                    Mp[k][l] = eps_l e_k + a_l o_k.
// Then log M_kl(x) of fcdiff/fit.py:117-122 equals
//                    log Mp[k][l] + tmax - log sqrt(2 pi)
// exactly (M_kl = eps_l p_k + (1-eps_l)/2 * sum_{j != k} p_j, fit.py:427-430),
// evaluated without the underflow of exp() in the far tails.
//
// e_k depends on (x, mu, sigma) only, and the reference never re-estimates mu,
// sigma (fit.py:232-237): the E/M/energy kernels therefore read a per-element
// *Gaussian cache* built once per fit by gauss_cache_kernel instead of x:
//   Ea = e of the first non-maximal component, Eb = e of the second one with the
//   index of the maximal component (whose e is exactly 1) in its two lowest
//   mantissa bits, Tm = max_k t_k.
struct ElemM {
    double e[3];
    double aS[3];     // a_l * (e_0 + e_1 + e_2)
    double mhS;       // -(e_0 + e_1 + e_2) / 2
};

struct GaussElem {
    double ea, ebc, tmax;
};

__device__ __forceinline__ GaussElem gauss_eval(double x, const ThetaDev& th) {
    double t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double z = (x - th.mu[k]) * th.isig[k];
        t[k] = fma(-0.5 * z, z, th.lc[k]);
    }
    // the maximal component has e = exp(0) = 1 exactly: only two exponentials
    const bool m0 = (t[0] >= t[1]) && (t[0] >= t[2]);
    const bool m1 = !m0 && (t[1] >= t[2]);
    const bool m01 = m0 || m1;
    GaussElem g;
    g.tmax = m0 ? t[0] : (m1 ? t[1] : t[2]);
    g.ea = exp_nonpos((m0 ? t[1] : t[0]) - g.tmax);
    const double eb = exp_nonpos((m01 ? t[2] : t[1]) - g.tmax);
    const int code = m0 ? 0 : (m1 ? 1 : 2);
    g.ebc = __hiloint2double(__double2hiint(eb), (__double2loint(eb) & ~3) | code);
    return g;
}

__device__ __forceinline__ ElemM elem_from_cache(double ea, double ebc, const ThetaDev& th) {
    const int code = __double2loint(ebc) & 3;
    const bool m0 = code == 0, m1 = code == 1;
    ElemM r;
    r.e[0] = m0 ? 1.0 : ea;
    r.e[1] = m1 ? 1.0 : (m0 ? ea : ebc);
    r.e[2] = (m0 || m1) ? ebc : 1.0;
    // Mp[k][l] = eps_l e_k + a_l (S - e_k) = b_l e_k + a_l S with b_l = eps_l - a_l: one FMA per
    // (k, l).  Where b_l < 0 the cancellation costs at most 1e-16 a_l / eps_l relative
    // (5e-12 at the optimiser's bound eps = 1e-5, 1e-15 at typical values).
    const double S = (1.0 + ea) + ebc;
#pragma unroll
    for (int l = 0; l < 3; ++l) r.aS[l] = th.al[l] * S;
    r.mhS = -0.5 * S;
    return r;
}

__device__ __forceinline__ double elem_Mp(const ElemM& r, const ThetaDev& th, int k, int l) {
    return fma(th.bl[l], r.e[k], r.aS[l]);
}

// e_k - o_k / 2 = 1.5 e_k - S / 2: numerator of d log M / d eps (fit.py:618-697)
__device__ __forceinline__ double elem_num(const ElemM& r, int k) { return fma(1.5, r.e[k], r.mhS); }

// q_R pair weights of fcdiff/fit.py:382-406.
__device__ __forceinline__ void pair_weights(double2 qn, double2 qm, double (&w)[3]) {
    w[0] = qn.x * qm.x;
    w[1] = qn.y * qm.y;
    w[2] = fma(qn.y, qm.x, qn.x * qm.y);
}

// Row walker shared by K2 / K3b / K4: every warp takes edge rows c = warp0,
// warp0 + nwarps, ...; within a row each lane owns the patients u = 2*lane,
// 2*lane + 1 (+64 per chunk), loaded with 128-bit requests (cache planes:
// ld.global.nc.L1::no_allocate; q_R rows of the edge's two regions: ld.global.nc).
// The loads of chunk j+1 (also across the row boundary) are issued before chunk
// j is computed, so the DRAM latency overlaps the ~400-cycle fp64 body instead
// of stalling its first use (ncu: 24 % of the samples before this change).
// Elements beyond U are presented with zero q_R, i.e. zero pair weights.
struct RowChunk {
    double2 xa, xb, tm, a0, a1, b0, b1;
};

template <bool WITH_TM>
__device__ __forceinline__ RowChunk load_chunk(const double* __restrict__ Ea, const double* __restrict__ Eb,
                                               const double* __restrict__ Tm, int64_t row_off,
                                               const double2* __restrict__ qn, const double2* __restrict__ qm,
                                               int u, int U) {
    RowChunk k;
    const double2 z = make_double2(0.0, 0.0);
    k.xa = k.xb = k.tm = k.a0 = k.a1 = k.b0 = k.b1 = z;
    if (u < U) {
        k.xa = ldg_stream2(Ea + row_off + u);
        k.xb = ldg_stream2(Eb + row_off + u);
        if (WITH_TM) k.tm = ldg_stream2(Tm + row_off + u);
        k.a0 = __ldg(qn + u);
        k.b0 = __ldg(qm + u);
        if (u + 1 < U) {
            k.a1 = __ldg(qn + u + 1);
            k.b1 = __ldg(qm + u + 1);
        }
    }
    return k;
}

// elem(ea, ebc, tm, qn, qm) is called for every (possibly zero-weight) element,
// row_end(c) once per row after its last element.  Requires pitchU even and
// 16-byte aligned planes.
template <bool WITH_TM, class ElemFn, class RowFn>
__device__ __forceinline__ void walk_rows(const double* __restrict__ Ea, const double* __restrict__ Eb,
                                          const double* __restrict__ Tm, int64_t C, int U, int64_t pitchU,
                                          const double* __restrict__ qR, int64_t c0,
                                          ElemFn&& elem, RowFn&& row_end) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double2* qR2 = reinterpret_cast<const double2*>(qR);
    int64_t c = warp0;
    if (c >= C) return;
    int n, m;
    c_to_nm(c0 + c, n, m);
    const double2* qn = qR2 + (int64_t)n * U;
    const double2* qm = qR2 + (int64_t)m * U;
    RowChunk cur = load_chunk<WITH_TM>(Ea, Eb, Tm, c * pitchU, qn, qm, 2 * lane, U);
    while (c < C) {
        const int64_t cn = c + nwarps;
        const double2 *qn2 = qn, *qm2 = qm;
        if (cn < C) {
            c_to_nm(c0 + cn, n, m);
            qn2 = qR2 + (int64_t)n * U;
            qm2 = qR2 + (int64_t)m * U;
        }
        for (int u0 = 0; u0 < U; u0 += 64) {
            RowChunk nxt;
            if (u0 + 64 < U) nxt = load_chunk<WITH_TM>(Ea, Eb, Tm, c * pitchU, qn, qm, u0 + 64 + 2 * lane, U);
            else if (cn < C) nxt = load_chunk<WITH_TM>(Ea, Eb, Tm, cn * pitchU, qn2, qm2, 2 * lane, U);
            else nxt = cur;
            elem(cur.xa.x, cur.xb.x, cur.tm.x, cur.a0, cur.b0);
            elem(cur.xa.y, cur.xb.y, cur.tm.y, cur.a1, cur.b1);
            cur = nxt;
        }
        row_end(c);
        c = cn;
        qn = qn2;
        qm = qm2;
    }
}

// Deterministic grid reduction: every CTA reduces NV values, writes its partial
// to ws[blockIdx.x*NV + i]; the CTA that arrives last sums the partials in a
// fixed order and writes out[0..NV).  The ticket (ws[kWsPartials]) is restored
// to zero.  Must be called by all threads of the block.
template <int NV, int THREADS>
__device__ __forceinline__ void grid_reduce_store(double (&v)[NV], double* ws, double* out) {
    static_assert(NV <= kMaxReduceVals, "too many values");
    __shared__ double s_part[NV][THREADS / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = warp_sum(v[i]);
        if (lane == 0) s_part[i][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < THREADS / 32) ? s_part[i][lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0) ws[(int64_t)blockIdx.x * NV + i] = x;
        }
    }
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(ws + kWsPartials);
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned long long t = atomicAdd(ticket, 1ull);
        s_last = (t == (unsigned long long)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += THREADS) {
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] += __ldcg(ws + (int64_t)b * NV + i);
    }
    __syncthreads();      // s_part reuse
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = warp_sum(acc[i]);
        if (lane == 0) s_part[i][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < THREADS / 32) ? s_part[i][lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0) out[i] = x;
        }
        if (lane == 0) *ticket = 0ull;
    }
}

#endif  // __CUDACC__

}  // namespace fcd
