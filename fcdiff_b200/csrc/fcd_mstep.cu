// M-step and free-energy kernels: K3a (pi / gamma sums), K3b (objective and
// analytic gradient of the (eta, epsilon) sub-problem), K4 (energy terms).
// All reductions accumulate in fp64: warp shuffle -> block -> deterministic
// last-CTA grid reduction.
#include "fcd_common.cuh"

namespace fcd {

constexpr int kRedThreads = 256;

// ------------------------------------------------------------------- K3a
// out[0..2] = sum_c exp(lqF[c,k]) (fcdiff/fit.py:219-220),
// out[3]    = sum_{n,u} exp(lqR[n,u,1]) (fit.py:212-213).
__global__ void __launch_bounds__(kRedThreads)
mstep_stats_kernel(const double* __restrict__ lqF, int64_t C,
                   const double* __restrict__ lqR, int64_t NU,
                   double* __restrict__ out, double* __restrict__ ws) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = tid; c < C; c += nth) {
        v[0] += exp(lqF[c * 3]);
        v[1] += exp(lqF[c * 3 + 1]);
        v[2] += exp(lqF[c * 3 + 2]);
    }
    for (int64_t i = tid; i < NU; i += nth) v[3] += exp(lqR[i * 2 + 1]);
    grid_reduce_store<4, kRedThreads>(v, ws, out);
}

// ------------------------------------------------------------------- K3b
// One pass over the Gaussian-cache planes of the local edge rows.  With
// A_k(c) = sum_u sum_l w_l log Mp_kl (fit.py:489-511 for the objective) and
// num_k = e_k - o_k / 2 (fit.py:600-697 for the gradient; the common factor
// exp(tmax)/sqrt(2 pi) of numerator and mixture cancels):
//   out[0] = sum_c sum_k qF_k A_k                         (theta-dependent part of E_lM)
//   out[1] = dE/d eta = -(2 eps - 1) sum_c sum_k qF_k sum_u w_2 num_k / Mp_k2
//   out[2] = dE/d eps = -sum_c sum_k qF_k sum_u sum_l s_l w_l num_k / Mp_kl,  s = (-1, 1, 2 eta - 1)
//   out[3] = sum_c (sum_k qF_k) sum_u (sum_l w_l)(tmax - log sqrt(2 pi))      (CONST; theta-free part)
// E_lM = out[0] + out[3].  When Aout != NULL the per-edge sums A_k(c) are also
// written: they are exactly what the next K2 and the energy need (no extra pass).
template <bool GRAD, bool FAST>
__device__ __forceinline__ void k3_elem(double ea, double ebc, double2 qn, double2 qm, double (&w)[3],
                                        const ThetaDev& th, double s2, const double* s_tab,
                                        double (&A)[3], double (&G)[3], double (&Hh)[3]) {
    pair_weights(qn, qm, w);
    const ElemM r = elem_from_cache(ea, ebc, th);
    const double sw2 = s2 * w[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double a = A[k], ds = 0.0, hs = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            const double M = elem_Mp(r, th, k, l);
            if (GRAD) {
                double rcp;
                a = fma(w[l], fast_log_rcp<FAST>(M, s_tab, rcp), a);
                if (l == 0) ds = -w[0] * rcp;
                else if (l == 1) ds = fma(w[1], rcp, ds);
                else {
                    ds = fma(sw2, rcp, ds);
                    hs = w[2] * rcp;
                }
            } else {
                a = fma(w[l], fast_log<FAST>(M, s_tab), a);
            }
        }
        A[k] = a;
        if (GRAD) {
            const double num = elem_num(r, k);
            G[k] = fma(num, ds, G[k]);
            Hh[k] = fma(num, hs, Hh[k]);
        }
    }
}

template <bool GRAD, bool CONST, bool VEC2, bool FAST>
__global__ void __launch_bounds__(kRedThreads, 2)
elm_kernel(const double* __restrict__ Ea, const double* __restrict__ Eb, const double* __restrict__ Tm,
           int64_t C, int U, int64_t pitchU,
           const double* __restrict__ qF, const double* __restrict__ qR, int N, int64_t c0,
           const __grid_constant__ ThetaDev th, const double* __restrict__ g_tab,
           double* __restrict__ Aout, double* __restrict__ out, double* __restrict__ ws) {
    extern __shared__ double s_tab[];               // kLogTabBytes when FAST, unused otherwise
    load_log_table<FAST>(g_tab, s_tab);
    const int lane = threadIdx.x & 31;
    const double s2 = 2.0 * th.eta - 1.0;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double A[3] = {0.0, 0.0, 0.0}, G[3] = {0.0, 0.0, 0.0}, Hh[3] = {0.0, 0.0, 0.0};
    double cs = 0.0;
    auto elem = [&](double ea, double ebc, double tm, double2 qn, double2 qm) {
        double w[3];
        k3_elem<GRAD, FAST>(ea, ebc, qn, qm, w, th, s2, s_tab, A, G, Hh);
        if (CONST) cs = fma(w[0] + w[1] + w[2], tm - kHalfLog2Pi, cs);
    };
    auto row_end = [&](int64_t c) {
        const double qf[3] = {__ldg(qF + c * 3), __ldg(qF + c * 3 + 1), __ldg(qF + c * 3 + 2)};
        if (Aout != nullptr) {                        // warp-uniform
#pragma unroll
            for (int k = 0; k < 3; ++k) A[k] = warp_sum(A[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) Aout[c * 3 + k] = A[k];
                acc[0] += fma(qf[0], A[0], fma(qf[1], A[1], qf[2] * A[2]));
            }
        } else {
            acc[0] += fma(qf[0], A[0], fma(qf[1], A[1], qf[2] * A[2]));
        }
        if (GRAD) {
            acc[1] -= fma(qf[0], Hh[0], fma(qf[1], Hh[1], qf[2] * Hh[2]));
            acc[2] -= fma(qf[0], G[0], fma(qf[1], G[1], qf[2] * G[2]));
        }
        if (CONST) acc[3] = fma(qf[0] + qf[1] + qf[2], cs, acc[3]);
#pragma unroll
        for (int k = 0; k < 3; ++k) A[k] = G[k] = Hh[k] = 0.0;
        cs = 0.0;
    };
    if (VEC2) {
        walk_rows<CONST>(Ea, Eb, Tm, C, U, pitchU, qR, c0, elem, row_end);
    } else {
        const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
        const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
        const double2* qR2 = reinterpret_cast<const double2*>(qR);
        for (int64_t c = warp0; c < C; c += nwarps) {
            int n, m;
            c_to_nm(c0 + c, n, m);
            const double2* qn = qR2 + (int64_t)n * U;
            const double2* qm = qR2 + (int64_t)m * U;
            for (int u = lane; u < U; u += 32)
                elem(ldg_stream1(Ea + c * pitchU + u), ldg_stream1(Eb + c * pitchU + u),
                     CONST ? ldg_stream1(Tm + c * pitchU + u) : 0.0, __ldg(qn + u), __ldg(qm + u));
            row_end(c);
        }
    }
    if (GRAD) acc[1] *= (2.0 * th.epsilon - 1.0);
    grid_reduce_store<4, kRedThreads>(acc, ws, out);
}

// ------------------------------------------------------------------- K4
// The six terms of fcdiff/fit.py:142-155; E_lM (v[3]) is produced by elm_kernel
// into the workspace scratch just before this kernel and folded in here.
//  v[0] E_lp_F = sum qF log gamma (fit.py:458)   v[1] E_lp_B_g_F (fit.py:472, via S1,S2)
//  v[2] E_lp_R = sum qR log[1-pi, pi] (fit.py:486)
//  v[4] E_lq_F = sum qF lqF (fit.py:525)         v[5] E_lq_R = sum qR lqR (fit.py:539)
__global__ void __launch_bounds__(kRedThreads)
energy_small_kernel(const double* __restrict__ S1, const double* __restrict__ S2,
                    const double* __restrict__ lqF, const double* __restrict__ qF, int64_t C,
                    const double* __restrict__ lqR, const double* __restrict__ qR, int64_t NU,
                    const __grid_constant__ ThetaDev th, int elm_known, double elm_value,
                    double* __restrict__ out6, double* __restrict__ ws) {
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0) v[3] = elm_known ? elm_value : ws[kWsScratch] + ws[kWsScratch + 3];
    for (int64_t c = tid; c < C; c += nth) {
        const double s1 = S1[c], s2 = S2[c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double l = lqF[c * 3 + k];
            const double q = qF[c * 3 + k];
            v[0] = fma(q, th.log_gamma[k], v[0]);
            v[1] = fma(q, fma(th.hq_a[k], s2, fma(th.hq_b[k], s1, th.hq_c[k])), v[1]);
            v[4] = fma(q, l, v[4]);
        }
    }
    for (int64_t i = tid; i < NU; i += nth) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double l = lqR[i * 2 + j];
            const double q = qR[i * 2 + j];
            v[2] = fma(q, th.log_pi2[j], v[2]);
            v[5] = fma(q, l, v[5]);
        }
    }
    grid_reduce_store<6, kRedThreads>(v, ws, out6);
}

static inline int red_grid(int64_t work_items, int items_per_block) {
    int64_t need = (work_items + items_per_block - 1) / items_per_block;
    int64_t cap = (int64_t)sm_count() * 8;
    if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static int launch_elm(const double* Ea, const double* Eb, const double* Tm,
                      int64_t C, int32_t U, int64_t pitchU,
                      const double* qF, const double* qR, int32_t N, int64_t c0,
                      const ThetaDev& th, bool grad, double* Aout, double* out4, double* ws, cudaStream_t st) {
    const double* tab = log_table(st);
    FCD_REQUIRE(tab != nullptr, "fcd_elm_obj_grad: log table initialisation failed");
    FCD_REQUIRE(grad || Tm != nullptr, "fcd_elm_obj_grad: nothing to compute (no gradient, no Tm)");
    int grid = red_grid(C, kRedThreads / 32);
    if (grid > sm_count() * 2) grid = sm_count() * 2;      // 2 CTAs / SM resident (86 KB table, <= 128 regs)
    uintptr_t al = reinterpret_cast<uintptr_t>(Ea) | reinterpret_cast<uintptr_t>(Eb) | reinterpret_cast<uintptr_t>(Tm);
    const bool vec2 = (pitchU % 2 == 0) && ((al & 15) == 0);
    const bool fast = log_table_covers(th.epsl, th.al);
#define FCD_ELM(G, K, V, F)                                                                        \
    do {                                                                                           \
        cudaFuncSetAttribute(elm_kernel<G, K, V, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)kLogTabBytes);                                                   \
        elm_kernel<G, K, V, F><<<grid, kRedThreads, (F) ? kLogTabBytes : 0, st>>>(                 \
            Ea, Eb, Tm, C, U, pitchU, qF, qR, N, c0, th, tab, Aout, out4, ws);                      \
    } while (0)
#define FCD_ELM_VF(G, K)                                                                           \
    do {                                                                                           \
        if (vec2) { if (fast) FCD_ELM(G, K, true, true); else FCD_ELM(G, K, true, false); }        \
        else      { if (fast) FCD_ELM(G, K, false, true); else FCD_ELM(G, K, false, false); }      \
    } while (0)
    if (grad) { if (Tm != nullptr) FCD_ELM_VF(true, true); else FCD_ELM_VF(true, false); }
    else      FCD_ELM_VF(false, true);
#undef FCD_ELM_VF
#undef FCD_ELM
    return check_launch("fcd_elm_obj_grad");
}

}  // namespace fcd

using namespace fcd;

extern "C" {

int fcd_mstep_stats(const double* lqF, int64_t C, const double* lqR, int64_t NU,
                    double* out4, double* ws, void* stream) {
    FCD_REQUIRE(C >= 0 && NU >= 0 && ws != nullptr, "fcd_mstep_stats: bad arguments");
    const int64_t work = C > NU ? C : NU;
    mstep_stats_kernel<<<red_grid(work, kRedThreads), kRedThreads, 0, (cudaStream_t)stream>>>(
        lqF, C, lqR, NU, out4, ws);
    return check_launch("fcd_mstep_stats");
}

int fcd_elm_obj_grad(const double* Ea, const double* Eb, const double* Tm,
                     int64_t C, int32_t U, int64_t pitchU,
                     const double* qF, const double* qR, int32_t N, int64_t c0,
                     const fcd_theta* theta_host, int32_t want_grad,
                     double* Aout, double* out4, double* ws, void* stream) {
    FCD_REQUIRE(theta_host != nullptr && ws != nullptr && Ea != nullptr && Eb != nullptr,
                "fcd_elm_obj_grad: NULL argument");
    FCD_REQUIRE(C >= 0 && U >= 1 && pitchU >= U && N >= 2, "fcd_elm_obj_grad: bad shape");
    FCD_REQUIRE(c0 >= 0 && c0 + C <= (int64_t)N * (N - 1) / 2, "fcd_elm_obj_grad: edge shard outside N=%d", N);
    const ThetaDev th = make_theta_dev(*theta_host, 0);
    return launch_elm(Ea, Eb, Tm, C, U, pitchU, qF, qR, N, c0, th, want_grad != 0, Aout, out4, ws,
                      (cudaStream_t)stream);
}

int fcd_energy_terms(const double* S1, const double* S2, int32_t H,
                     const double* Ea, const double* Eb, const double* Tm,
                     int64_t C, int32_t U, int64_t pitchU,
                     const double* lqF, const double* qF, const double* lqR, const double* qR,
                     int32_t N, int64_t c0, const fcd_theta* theta_host, const double* elm_host,
                     double* out6, double* ws, void* stream) {
    FCD_REQUIRE(theta_host != nullptr && ws != nullptr, "fcd_energy_terms: NULL argument");
    FCD_REQUIRE(C >= 0 && U >= 1 && pitchU >= U && N >= 2 && H >= 1, "fcd_energy_terms: bad shape");
    FCD_REQUIRE(c0 >= 0 && c0 + C <= (int64_t)N * (N - 1) / 2, "fcd_energy_terms: edge shard outside N=%d", N);
    const ThetaDev th = make_theta_dev(*theta_host, H);
    cudaStream_t st = (cudaStream_t)stream;
    if (elm_host == nullptr) {      // otherwise E_lM of this shard is known from the last K3b evaluation
        FCD_REQUIRE(Ea != nullptr && Eb != nullptr && Tm != nullptr, "fcd_energy_terms: cache planes are NULL");
        int rc = launch_elm(Ea, Eb, Tm, C, U, pitchU, qF, qR, N, c0, th, false, nullptr, ws + kWsScratch, ws, st);
        if (rc) return rc;
    }
    const int64_t NU = (int64_t)N * U;
    const int64_t work = C > NU ? C : NU;
    energy_small_kernel<<<red_grid(work, kRedThreads), kRedThreads, 0, st>>>(
        S1, S2, lqF, qF, C, lqR, qR, NU, th, elm_host != nullptr, elm_host ? *elm_host : 0.0, out6, ws);
    return check_launch("fcd_energy_terms");
}

}  // extern "C"
