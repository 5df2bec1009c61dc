// Lean fp64 transcendental building blocks for the per-element core.
//
// The hot kernels are bound by the fp64 pipe (64 DFMA/clk/SM on B200), not by
// HBM: per edge-patient element the reference needs 3 exp and 9 log
// (fcdiff/fit.py:115, 122).  CUDA's general-purpose exp()/log() cost ~450 fp64
// instructions per element (ncu, profiles/r01a_*); the routines below exploit
// what is known about the arguments:
//
//  * fast_log(y): y is a mixture weight in (0, 1].  MUFU.RCP64H gives r ~ 1/y;
//    r is rounded to 9 mantissa bits (r8, exactly representable), so that
//    u = y*r8 - 1 is one exact FMA with |u| <= 2^-10, and
//        log y = -log r8 + log1p(u),
//    -log r8 from a 10752-entry table in shared memory (indexed by the
//    exponent/mantissa bits of r8), log1p(u) = u - u^2/2 + u^3/3
//    (truncation u^4/4 <= 2.3e-13).  5 fp64 ops + 1 MUFU + 1 LDS instead of ~40.
//  * fast_log_rcp additionally returns 1/y = r8 * (1 - u + u^2 - u^3 + u^4), rel.
//    error <= u^5 = 2^-50, for the analytic gradient (4 more fp64 ops, no second
//    MUFU / Newton iteration).
//  * exp_nonpos(d): d <= 0, argument reduction with the 1.5*2^52 trick (no
//    F2I/I2F conversions), degree-11 polynomial, exponent patched in the ALU.
//
// Absolute error of fast_log < 3e-13, relative error of exp_nonpos < 1e-14:
// more than three orders of magnitude inside the 1e-9 per-term budget that 1e-6
// parity of the posteriors needs (DESIGN.md "Numerics").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace fcd {

constexpr int kLogTabBits    = 9;                       // mantissa bits kept in the rounded reciprocal
constexpr int kLogTabMinExp  = -1;                      // r8 in [2^-1, 2^20)  <=>  y in (2^-20, 2]
constexpr int kLogTabBinades = 21;
constexpr int kLogTabSize    = kLogTabBinades << kLogTabBits;      // 10752 doubles = 86,016 bytes
constexpr int kLogTabBase    = (1023 + kLogTabMinExp) << kLogTabBits;
constexpr size_t kLogTabBytes = (size_t)kLogTabSize * sizeof(double);      // dynamic shared memory of the FAST kernels

// Host: device address of the table for the current device (built on first use).
const double* log_table(cudaStream_t st);
// True when every mixture weight a_l + b_l p (0 <= p <= 1) lies inside the table's range.
bool log_table_covers(const double epsl[3], const double al[3]);

// The slots [lo, lo + n) of the table that arguments in [min_l min(a_l, eps_l),
// max_l max(a_l, eps_l)] can index: only this window is staged in shared memory
// (typically 8 of the 21 binades, 32 KB instead of 86 KB -> more CTAs per SM).
struct LogTabWindow {
    const double* g;      // device table (all slots)
    int lo;
    int n;
    size_t bytes() const { return (size_t)n * sizeof(double); }
};
// with_mantissa: the window also starts at slot 0, i.e. includes the binade of arguments in [1, 2]
// that log_pos (below) needs for the mantissa of an arbitrary positive double (+ 4 KB).
bool log_table_window(const double epsl[3], const double al[3], cudaStream_t st, LogTabWindow& w,
                      bool with_mantissa = false);

#ifdef __CUDACC__

// -log(r8) for table slot i.
__device__ __forceinline__ double log_table_r8(int i) {
    return __hiloint2double((i + kLogTabBase) << (20 - kLogTabBits), 0);
}

// Stages the window in shared memory; returns the pointer to index with the
// FULL-table slot number (s_tab - lo: only slots inside the window are touched).
template <bool FAST>
__device__ __forceinline__ const double* load_log_table(const LogTabWindow& w, double* s_tab) {
    if (FAST)
        for (int i = threadIdx.x; i < w.n; i += blockDim.x) s_tab[i] = w.g[w.lo + i];
    __syncthreads();
    return s_tab - w.lo;
}

// The same staging as ONE bulk copy (window slots and size are even, log_table_window): issued by a
// single thread ahead of a kernel's own ring fills, so that the table does not queue behind them;
// every thread waits on `bar` (phase 0) before its first logarithm.
template <bool FAST>
__device__ __forceinline__ void issue_log_table(const LogTabWindow& w, double* s_tab, uint64_t* bar) {
    if (!FAST) return;
    const uint32_t bytes = (uint32_t)w.n * 8u, b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(s_tab)), "l"(w.g + w.lo), "r"(bytes), "r"(b) : "memory");
}

struct LogParts {
    double r8;      // 8-bit reciprocal
    double u;       // y * r8 - 1
    int idx;        // table slot
};

// Requires 2^-20 < y <= 1 (+ rounding): guaranteed by the host for the mixture
// weights when min(eps_l, (1-eps_l)/2) > 2^-19 (log_table_covers); otherwise the
// host launches the SAFE kernel variants, which use log() and division.
__device__ __forceinline__ LogParts log_reduce(double y) {
    LogParts p;
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(y));            // MUFU.RCP64H
    const int hi = (__double2hiint(r0) + (1 << (19 - kLogTabBits))) & ~((1 << (20 - kLogTabBits)) - 1);
    p.r8 = __hiloint2double(hi, 0);
    p.idx = (hi >> (20 - kLogTabBits)) - kLogTabBase;
    p.u = fma(y, p.r8, -1.0);
    return p;
}

// |u| <= 2^-10: u - u^2/2 + u^3/3, truncation u^4/4 <= 2.3e-13
__device__ __forceinline__ double log1p_small(double u) {
    const double q = fma(u, 1.0 / 3.0, -0.5);
    return fma(u * u, q, u);
}

template <bool FAST>
__device__ __forceinline__ double fast_log(double y, const double* s_tab) {
    if (!FAST) return log(y);
    const LogParts p = log_reduce(y);
    return s_tab[p.idx] + log1p_small(p.u);
}

// 1 / y alone (same reduction, no table access)
template <bool FAST>
__device__ __forceinline__ double fast_rcp(double y) {
    if (!FAST) return 1.0 / y;
    const LogParts p = log_reduce(y);
    double g = fma(p.u, p.u, 1.0 - p.u);             // 1 - u + u^2
    g = fma(-p.u * p.u, p.u, g);                     // - u^3          (u^4 <= 9.1e-13 relative)
    return p.r8 * g;
}

template <bool FAST>
__device__ __forceinline__ double fast_log_rcp(double y, const double* s_tab, double& rcp) {
    if (!FAST) {
        rcp = 1.0 / y;
        return log(y);
    }
    const LogParts p = log_reduce(y);
    // 1 / (1 + u) = 1 - u + u^2 - u^3 + u^4 (- u^5 <= 2^-50): the gradient sums ~10^7 such terms with
    // cancellation, so a one-sided truncation error of 1e-9 per term would show at 1e-3 of the net gradient
    const double om = 1.0 - p.u, uu = p.u * p.u;
    const double g = fma(uu, fma(p.u, p.u, om), om);
    rcp = p.r8 * g;
    return s_tab[p.idx] + fma(uu, fma(p.u, 1.0 / 3.0, -0.5), p.u);
}

// Sum of logs as the log of a product.  Where a kernel adds log(y_i) with weight exactly 1 (tier T1),
// it multiplies the y_i into a running product instead and takes ONE logarithm per kProdMax factors:
// y_i > 2^-20 (log_table_covers) bounds a product of 2 x kProdMax factors below by 2^-960 (no underflow),
// and the relative rounding error of n multiplications, n 2^-53, is an ABSOLUTE error of the sum of logs
// -- smaller than the table logarithm's 3e-13 per term.  One multiplication instead of ~12 instructions.
constexpr int kProdMax = 24;      // factors per running product; two products are merged before the log

// log(x) for any positive normal double: x = m 2^e, m in [1, 2); needs the window's mantissa binade
// (log_table_window(..., with_mantissa = true)).
template <bool FAST>
__device__ __forceinline__ double log_pos(double x, const double* s_tab) {
    if (!FAST) return log(x);
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | (1023 << 20), __double2loint(x));
    const LogParts p = log_reduce(m);
    // (double)e without the conversion pipe: 2^52 + 2^31 + e as an integer bit pattern, minus the bias (exact)
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;
    return fma(ed, 6.93147180559945286227e-01, s_tab[p.idx] + log1p_small(p.u));
}

// 1 / y to rounding: MUFU.RCP64H (relative error e ~ 2^-18) and r0 (1 + e + e^2), e = 1 - y r0 exact
// to one rounding; the truncation e^3 ~ 2^-54 leaves no systematic bias in sums of 10^7 terms
// (a plain Newton step's -e^2 does).  No table, no integer work.
__device__ __forceinline__ double rcp_newton(double y) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(y));
    const double e = fma(-y, r0, 1.0);
    return fma(r0, fma(e, e, e), r0);
}

// e^d for d <= 0 (d is clamped at -700: e^-700 ~ 1e-304 is far below any
// weight that can matter next to the maximal component e = 1).
__device__ __forceinline__ double exp_nonpos(double d) {
    d = fmax(d, -700.0);
    const double kMagic = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(d, 1.4426950408889634074, kMagic);
    const int n = __double2loint(t);
    const double nd = t - kMagic;
    double r = fma(nd, -6.93147180369123816490e-01, d);       // ln2 hi (32 trailing zero bits)
    r = fma(nd, -1.90821492927058770002e-10, r);              // ln2 lo
    double p = 2.50521083854417187751e-08;                    // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);                // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);                // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);                // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);                // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);                // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);                // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);                // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);                // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// The same function with the polynomial in Estrin form: 4 dependent FMAs after the argument reduction
// instead of 11 (for the one place where the exponential sits on a dependent chain: the in-block steps of
// the blocked region sweep).  Same coefficients; the rounding differs from the Horner form by ~1e-16.
__device__ __forceinline__ double exp_nonpos_short(double d) {
    d = fmax(d, -700.0);
    const double kMagic = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = fma(d, 1.4426950408889634074, kMagic);
    const int n = __double2loint(t);
    const double nd = t - kMagic;
    double r = fma(nd, -6.93147180369123816490e-01, d);
    r = fma(nd, -1.90821492927058770002e-10, r);
    const double r2 = r * r;
    const double p01 = 1.0 + r;
    const double p23 = fma(r, 1.66666666666666666667e-01, 0.5);
    const double p45 = fma(r, 8.33333333333333333333e-03, 4.16666666666666666667e-02);
    const double p67 = fma(r, 1.98412698412698412698e-04, 1.38888888888888888889e-03);
    const double p89 = fma(r, 2.75573192239858906526e-06, 2.48015873015873015873e-05);
    const double pab = fma(r, 2.50521083854417187751e-08, 2.75573192239858906526e-07);
    const double r4 = r2 * r2;
    const double q0 = fma(r2, p23, p01);
    const double q1 = fma(r2, p67, p45);
    const double q2 = fma(r2, pab, p89);
    const double r8 = r4 * r4;
    const double p = fma(r8, q2, fma(r4, q1, q0));
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

#endif  // __CUDACC__

}  // namespace fcd
