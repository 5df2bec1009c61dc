/*
 * fcdiff_b200 -- C-ABI of the B200-native hot path of andy-sweet/fcdiff
 * (variational EM for the individual-anomalous-region model).
 *
 * The reference is pure Python and has no FFI of its own: the
 * drop-in boundary is its Python surface (fcdiff.fit / fcdiff.model /
 * fcdiff.util), re-implemented in the package `fcdiff_b200` (alias `fcdiff`)
 * which binds THIS library with ctypes.  Each entry point below cites the
 * reference function (path:line under the reference tree) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - arrays are float64, C-contiguous, row pitch given in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - every call returns 0 on success, <0 on error; fcd_last_error() gives the
 *     text of the last error raised on the calling thread;
 *   - no entry point allocates device memory: scratch comes from `ws`
 *     (a device buffer of at least fcd_workspace_bytes() bytes, zero-filled
 *     once by the caller before first use; calls restore it to zero);
 *   - edge order is the reference's lower-triangular row-major order
 *     c = n(n-1)/2 + m, m < n (fcdiff/util.py:40-84);
 *   - an edge shard is the contiguous range [c0, c0 + C) of the global edge
 *     list; `N` is always the global number of regions.
 */
#ifndef FCDIFF_B200_H
#define FCDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCD_VERSION 100

#if defined(__GNUC__)
#define FCD_API __attribute__((visibility("default")))
#else
#define FCD_API
#endif

/* Model parameters theta (fcdiff/model.py:31-38). */
typedef struct fcd_theta {
    double pi;        /* P(region anomalous)                                   */
    double eta;       /* P(edge anomalous | exactly one end-point anomalous)   */
    double epsilon;   /* P(typical edge differs from the template)             */
    double gamma[3];  /* template state prior (negative, none, positive)       */
    double mu[3];     /* Gaussian means per state                              */
    double sigma[3];  /* Gaussian standard deviations per state                */
} fcd_theta;

enum { FCD_LOOKUP_REFERENCE = 0, FCD_LOOKUP_SYMMETRIC = 1 };

/* ------------------------------------------------------------------ runtime */
FCD_API int fcd_version(void);
FCD_API const char* fcd_last_error(void);
/* SM count and compute capability of the current device. */
FCD_API int fcd_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);
FCD_API int64_t fcd_workspace_bytes(void);
/* Number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`), and reset. */
FCD_API int64_t fcd_launch_count(void);
FCD_API void fcd_launch_count_reset(void);

/* ------------------------------------------------------- index arithmetic   */
/* fcdiff/util.py:62-84 c_to_nm for c in [c0, c0+C): n_out[i], m_out[i] int32. */
FCD_API int fcd_c_to_nm(int64_t c0, int64_t C, int32_t* n_out, int32_t* m_out, void* stream);

/* ------------------------------------------------------- fused hot path     */
/* Healthy-subject sufficient statistics, computed once per fit:
 *   S1[c] = sum_h b[c,h],  S2[c] = sum_h b[c,h]^2.
 * They replace the (C,H,3) cache `_lp_B_g_F` of fcdiff/fit.py:111-114: the
 * healthy log-density sum of fit.py:171 is a quadratic in (S1, S2). */
FCD_API int fcd_healthy_stats(const double* b, int64_t C, int32_t H, int64_t pitchH,
                      double* S1, double* S2, void* stream);

/* Gaussian cache, built once per fit (the reference never re-estimates mu,
 * sigma, fcdiff/fit.py:232-237): for every patient correlation x = bt[c,u], with
 * t_k = log N(x; mu_k, sigma_k) + log sqrt(2 pi) and e_k = exp(t_k - max_j t_j),
 *   Ea[c,u] = e of the first non-maximal state, Eb[c,u] = e of the second one
 *   (index of the maximal state, whose e is exactly 1, in its 2 lowest mantissa
 *   bits), Tm[c,u] = max_j t_j (may be NULL).
 * All three have the shape and pitch of bt.  This is the patient half of
 * `_update_lps` (fit.py:115) in a form that cannot underflow; the kernels below
 * read these planes instead of bt. */
FCD_API int fcd_gauss_cache(const double* bt, int64_t C, int32_t U, int64_t pitchU,
                    const fcd_theta* theta_host, double* Ea, double* Eb, double* Tm, void* stream);

/* K2 -- E-step for the template posterior; replaces `_update_lq_F`
 * (fcdiff/fit.py:157-174) + `_eval_q_R_w` (fit.py:382-406) + the patient half
 * of `_update_lps` (fit.py:115-122) + `_eval_M` (fit.py:409-444).
 *   lqF[c,k] = log gamma_k + sum_h logN_k(b[c,h])
 *            + sum_u sum_l w_l(n,m,u) log M_kl(bt[c,u]),  minus logsumexp_k.
 * qR is [N][U][2] probabilities.  Outputs lqF [C][3] and qF=exp(lqF) [C][3]. */
FCD_API int fcd_estep_qF(const double* S1, const double* S2, int32_t H,
                 const double* Ea, const double* Eb, int64_t C, int32_t U, int64_t pitchU,
                 const double* qR, int32_t N, int64_t c0,
                 const fcd_theta* theta_host,
                 double* lqF, double* qF, void* stream);

/* K2 without a pass over the data: A[c][k] = sum_u sum_l w_l log M'_kl is the
 * per-edge output of an fcd_elm_obj_grad call made with the same q_R and
 * (eta, epsilon); only log gamma, the healthy term and the normalisation of
 * fit.py:165-174 remain. */
FCD_API int fcd_estep_qF_finish(const double* S1, const double* S2, int32_t H, const double* A, int64_t C,
                        const fcd_theta* theta_host, double* lqF, double* qF, void* stream);

/* Patient-major copy of the patient correlations: btT[u - u0][c] = bt[c][u]
 * for u in [u0, u0+Ul), c in [0, C).  Built once per fit. */
FCD_API int fcd_transpose_patients(const double* bt, int64_t C, int32_t U, int64_t pitchU,
                           int32_t u0, int32_t Ul, double* btT, int64_t pitchC, void* stream);

/* K2b part 1 -- q_R-independent half of `_update_lq_R` (fcdiff/fit.py:187-194):
 *   WT[u][c][l] = sum_k qF[c,k] * log M_kl(bt[c][u])   (patient-major; EaT/EbT are the
 *   patient-major copies [Ul][pitchC] of the Ea/Eb cache planes)
 * up to an additive per-(c,u) constant common to all l, which cancels in the
 * normalisation of fit.py:196. qF is [C][3] for ALL edges. */
FCD_API int fcd_region_weights(const double* EaT, const double* EbT, int32_t Ul, int64_t C, int64_t pitchC,
                       const double* qF, const fcd_theta* theta_host,
                       double* WT, void* stream);

/* K2b part 2 -- Gauss-Seidel sweep of `_update_lq_R` (fcdiff/fit.py:176-198)
 * for the patients [u0, u0+Ul).  qR/lqR are the full [N][U][2] arrays; only
 * the [u0, u0+Ul) columns are read and written.  log_pi2_host = {log(1-pi),
 * log(pi)} (the 2-vector convention of test_fcdiff/test_fit.py:477-487).
 * edge_lookup: FCD_LOOKUP_REFERENCE reproduces nm_to_c(n, m) for all m != n
 * (fit.py:185-186), FCD_LOOKUP_SYMMETRIC the unordered pair. */
FCD_API int fcd_estep_qR(const double* WT, int64_t C, int32_t N, int32_t U, int32_t u0, int32_t Ul,
                 const double* log_pi2_host, int32_t edge_lookup,
                 double* qR, double* lqR, void* stream);

/* K3a -- M-step sums; replaces `_update_pi` / `_update_gamma`
 * (fcdiff/fit.py:208-220).  out[0..2] = sum_c exp(lqF[c,k]),
 * out[3] = sum_{n,u} exp(lqR[n,u,1]).  The caller divides by counts (after an
 * all-reduce of out[0..2] when edges are sharded). */
FCD_API int fcd_mstep_stats(const double* lqF, int64_t C, const double* lqR, int64_t NU,
                    double* out4, double* ws, void* stream);

/* K3b -- objective and analytic gradient of the (eta, epsilon) sub-problem;
 * replaces `_opt_fun` (fcdiff/fit.py:270-286) = `_update_lps` + `_eval_E_lM`
 * (fit.py:489-511) and `_eval_dE_dh` / `_eval_dE_de` / `_eval_dlM_dh` /
 * `_eval_dlM_de` (fit.py:600-697) in ONE pass over the cache planes.
 *   out[0] + out[3] = E_lM = sum_c sum_k qF[c,k] sum_u sum_l w_l log M_kl(bt[c,u])
 *       out[3] is the part that does not depend on (eta, epsilon); it is only
 *       computed when Tm != NULL (else 0)
 *   out[1] = dE/d eta, out[2] = dE/d epsilon   (of E = -E_lM; only if want_grad)
 * Aout (may be NULL): per-edge sums A[c][k] for fcd_estep_qF_finish. */
FCD_API int fcd_elm_obj_grad(const double* Ea, const double* Eb, const double* Tm,
                     int64_t C, int32_t U, int64_t pitchU,
                     const double* qF, const double* qR, int32_t N, int64_t c0,
                     const fcd_theta* theta_host, int32_t want_grad,
                     double* Aout, double* out4, double* ws, void* stream);

/* K4 -- free-energy terms; replaces `_eval_energy` and `_eval_E_*`
 * (fcdiff/fit.py:142-155, 447-539).  out[0..5] = E_lp_F, E_lp_B_g_F, E_lp_R,
 * E_lM, E_lq_F, E_lq_R over the local edge shard (terms 2 and 5 involve q_R
 * only and are complete on every rank).  qF = exp(lqF), qR = exp(lqR) as left
 * by fcd_estep_qF / fcd_estep_qR (fit.py:146-147).  elm_host (may be NULL): the
 * shard's E_lM when the caller already has it from fcd_elm_obj_grad at the same
 * (q_F, q_R, theta); then no pass over the data is made. */
FCD_API int fcd_energy_terms(const double* S1, const double* S2, int32_t H,
                     const double* Ea, const double* Eb, const double* Tm,
                     int64_t C, int32_t U, int64_t pitchU,
                     const double* lqF, const double* qF, const double* lqR, const double* qR,
                     int32_t N, int64_t c0, const fcd_theta* theta_host, const double* elm_host,
                     double* out6, double* ws, void* stream);

/* ------------------------------------------------------- materialised API   */
/* The reference caches (C,H,3), (C,U,3) and (C,U,3,3) arrays and its tests
 * assign them directly (test_fcdiff/test_fit.py:440-446, 483-489).  These
 * entry points run the same steps from such arrays (API parity; the fused
 * path above never materialises them). */

/* `_update_lps` (fcdiff/fit.py:104-122): lpB [C][H][3], pBt [C][U][3],
 * lM [C][U][3][3] (k major, l minor), scipy.stats.norm op order. */
FCD_API int fcd_materialize_lps(const double* b, const double* bt, int64_t C, int32_t H, int32_t U,
                        const fcd_theta* theta_host,
                        double* lpB, double* pBt, double* lM, void* stream);

/* `_eval_M` (fcdiff/fit.py:409-430): M[i] from p[i][3], i < n. */
FCD_API int fcd_eval_M(const double* p, int64_t n, double eta, double epsilon, int32_t k, int32_t l,
               double* M, void* stream);

/* `_update_lq_F` from arrays (fcdiff/fit.py:157-174). */
FCD_API int fcd_lqF_from_arrays(const double* lpB, const double* lM, int64_t C, int32_t H, int32_t U,
                        const double* qR, int32_t N, const double* log_gamma_host,
                        double* lqF, void* stream);

/* W for `_update_lq_R` from a materialised lM: WT[u][c][l] = sum_k qF[c,k] lM[c,u,k,l]. */
FCD_API int fcd_region_weights_from_lM(const double* lM, int64_t C, int32_t U, const double* qF,
                               double* WT, void* stream);

/* `_eval_E_lM` (fcdiff/fit.py:489-511) from arrays. */
FCD_API int fcd_ElM_from_arrays(const double* qF, const double* qR, const double* lM,
                        int64_t C, int32_t N, int32_t U, double* out1, double* ws, void* stream);

/* `_eval_dE_dh` / `_eval_dE_de` (fcdiff/fit.py:600-664) from arrays:
 * out[0] = dE/d eta, out[1] = dE/d epsilon. norm [C][U][3], mix [C][U][3][3]. */
FCD_API int fcd_dE_from_arrays(const double* qR, const double* qF, const double* norm, const double* mix,
                       int64_t C, int32_t N, int32_t U, double eta, double epsilon,
                       double* out2, double* ws, void* stream);

/* `_eval_dlM_dh` / `_eval_dlM_de` (fcdiff/fit.py:618-641, 667-697), elementwise:
 * out[i] = (eps * norm[i][k] - 0.5 * eps * (norm[i][j] + norm[i][j'])) / mix[i]. */
FCD_API int fcd_dlM(const double* norm, const double* mix, int64_t n, double eps, int32_t k,
            double* out, void* stream);

/* `_eval_q_R_w` (fcdiff/fit.py:382-406): out[u][0..2] for the region pair (n, m). */
FCD_API int fcd_pair_weights(const double* qR, int32_t N, int32_t U, int32_t n, int32_t m,
                     double* out, void* stream);

/* sum_i a[i] * x[(i / a_div) % x_mod_or_n]: the broadcast dot products of
 * `_eval_E_lp_F`, `_eval_E_lp_B_g_F`, `_eval_E_lp_R`, `_eval_E_lq_F`,
 * `_eval_E_lq_R` (fcdiff/fit.py:447-486, 514-539).  See fit.py (host) for the
 * index maps used. out1[0] = sum_i a[ia(i)] * x[ix(i)] with
 *   ia(i) = (i / a_outer) * a_inner + i % a_inner   (a_outer >= a_inner)
 *   ix(i) = i % x_len. */
FCD_API int fcd_dot_broadcast(const double* a, int64_t a_outer, int64_t a_inner,
                      const double* x, int64_t x_len, int64_t n,
                      double* out1, double* ws, void* stream);

/* ------------------------------------------------------- sampler (K5)       */
/* Ancestral sampler of fcdiff/model.py:52-236 with counter-based Philox4x32-10
 * (key = seed, counter = (variable id, c, u)); util edge order everywhere.
 * Outputs (any may be NULL to skip): r [N][U] u8, t [C][U] u8, f [C][3] u8,
 * ft [C][U][3] u8, b [C][H] f64, bt [C][U] f64 (clipped to [-1,1]).
 * Each stage can also be driven from caller-provided parents (sample_T(r),
 * sample_F_tilde(f,t), sample_B(f,H), sample_B_tilde(f_tilde)). */
FCD_API int fcd_sample_R(uint64_t seed, uint64_t offset, int32_t N, int32_t U, double pi,
                 uint8_t* r, void* stream);
FCD_API int fcd_sample_T(uint64_t seed, uint64_t offset, const uint8_t* r, int32_t N, int32_t U,
                 double eta, int64_t c0, int64_t C, uint8_t* t, void* stream);
FCD_API int fcd_sample_F(uint64_t seed, uint64_t offset, int64_t c0, int64_t C,
                 const double* gamma3_host, uint8_t* f, void* stream);
FCD_API int fcd_sample_F_tilde(uint64_t seed, uint64_t offset, const uint8_t* f, const uint8_t* t,
                       int64_t c0, int64_t C, int32_t U, double epsilon, uint8_t* ft, void* stream);
FCD_API int fcd_sample_B(uint64_t seed, uint64_t offset, const uint8_t* f, int64_t c0, int64_t C, int32_t H,
                 const double* mu3_host, const double* sigma3_host, double* b, void* stream);
FCD_API int fcd_sample_B_tilde(uint64_t seed, uint64_t offset, const uint8_t* ft, int64_t c0, int64_t C, int32_t U,
                       const double* mu3_host, const double* sigma3_host, double* bt, void* stream);

/* ------------------------------------------------------- correlations (K1)  */
/* Region time series -> correlations (no reference counterpart).
 * ts [S][N][T] float32.  Per subject: rows are centred and scaled to unit
 * norm, R = Z Z^T on tcgen05 tensor cores with error-compensated split-TF32
 * (3 MMAs per product), and the lower triangle is written in util edge order to
 * out[c][s0 + s] (float64, row pitch `pitch`), Fisher-z transformed
 * (atanh) when `fisher` != 0, else clipped to [-1, 1].
 * zws: device scratch of fcd_corr_workspace_bytes(S, N, T) bytes. */
FCD_API int64_t fcd_corr_workspace_bytes(int32_t S, int32_t N, int32_t T);
FCD_API int fcd_corr_fisherz(const float* ts, int32_t S, int32_t N, int32_t T,
                     double* out, int64_t pitch, int32_t s0, int32_t fisher,
                     void* zws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FCDIFF_B200_H */
