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

#define FCD_VERSION 120

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
/* Copies `bytes` from device memory to (pinned) host memory on `stream` and
 * waits for the stream: how the few doubles a reduction kernel leaves behind
 * reach the host-side optimiser / convergence test. */
FCD_API int fcd_download(void* dst_host, const void* src, int64_t bytes, void* stream);
/* Number of kernels this library has launched in this process (bench.py's
 * `gpu_launches`), and reset. */
FCD_API int64_t fcd_launch_count(void);
FCD_API void fcd_launch_count_reset(void);
/* NVTX range around a kernel family (profiling aid, SURVEY 5 "tracing / profiling": the reference has
 * none).  `name_host` is copied by the tools; ranges nest per thread.  No-ops without a profiler. */
FCD_API void fcd_nvtx_push(const char* name_host);
FCD_API void fcd_nvtx_pop(void);

/* ------------------------------------------------------- small exchanges    */
/* One-shot all-reduce (sum) of n <= fcd_comm_max_vals() doubles over NVLink peer
 * memory, result delivered to the host without a copy or a stream synchronisation
 * (csrc/fcd_comm.cu).  Replaces, for the edge-sharded fit, the NCCL all-reduce +
 * download after every reduction kernel: the M-step sums (fcdiff/fit.py:208-220),
 * the objective / gradient partial sums of every (eta, epsilon) evaluation
 * (fit.py:270-286) and the energy terms (fit.py:142-155).
 *   window: fcd_comm_window_bytes() of device memory per rank, created here (the one
 *     explicit allocation of this library), exported / opened as a CUDA IPC handle of
 *     fcd_comm_handle_bytes() bytes; windows_host[world] lists every rank's window in
 *     THIS process' address space (own window at [rank]);
 *   result_host: mapped pinned memory from fcd_host_result_alloc;
 *   seq: 1, 2, 3, ... -- the same on every rank for the same exchange;
 *   fcd_allreduce_small launches one kernel on `stream`: vec (device, in place) and
 *     result_host receive the sums, formed in rank order (bit-identical on all ranks);
 *     world == 1 publishes vec to the host only;
 *   fcd_wait_result spins until exchange seq has been published, copies n doubles. */
FCD_API int64_t fcd_comm_window_bytes(void);
FCD_API int32_t fcd_comm_handle_bytes(void);
FCD_API int32_t fcd_comm_max_world(void);
FCD_API int32_t fcd_comm_max_vals(void);
FCD_API int fcd_comm_window_create(void** window_out_host);
FCD_API int fcd_comm_window_destroy(void* window);
FCD_API int fcd_comm_window_export(void* window, void* handle_host);
FCD_API int fcd_comm_window_open(const void* handle_host, void** peer_window_out_host);
FCD_API int fcd_comm_window_close(void* peer_window);
FCD_API int fcd_host_result_alloc(void** result_out_host);
FCD_API int fcd_host_result_free(void* result_host);
FCD_API int fcd_allreduce_small(double* vec, int32_t n, void* const* windows_host, int32_t rank, int32_t world,
                        uint64_t seq, double* result_host, void* stream);
/* The same exchange; behind the n sums, result_host also receives this rank's OWN values
 * vec[keep0 .. keep0 + nkeep) as they were before the sum (n + nkeep <= fcd_comm_max_vals()): one wait
 * serves the M-step sums of fit.py:208-220, the record counts of all ranks (total: which form the
 * (eta, epsilon) evaluations take, identical on every rank) and of this rank (the size of its lists). */
FCD_API int fcd_allreduce_small_keep(double* vec, int32_t n, int32_t keep0, int32_t nkeep, void* const* windows_host,
                             int32_t rank, int32_t world, uint64_t seq, double* result_host, void* stream);
FCD_API int fcd_wait_result(const double* result_host, int32_t n, uint64_t seq, double* out_host, int32_t timeout_ms);
/* Staging of the all-gather of the region posteriors `_lq_R` / exp(_lq_R) ((N, U, 2), fcdiff/fit.py:176-198)
 * when patients are sharded: pack this rank's columns [u0, u0+Ul) of both arrays into one
 * contiguous block [2][N][ch][2] (ch columns per rank, zero padded); unpack the gathered
 * [world][2][N][ch][2] into the two full arrays. */
FCD_API int fcd_pack_patients(const double* lqR, const double* qR, int32_t N, int32_t U, int32_t u0, int32_t Ul,
                      int32_t ch, double* out, void* stream);
FCD_API int fcd_unpack_patients(const double* gathered, int32_t world, int32_t N, int32_t U, int32_t ch, double* lqR,
                        double* qR, void* stream);

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

/* Edge table of the local shard: nm[i] = n | m << 16 for the edge c0 + i
 * (fcdiff/util.py:62-84 c_to_nm; at most 65535 regions).  Built once per fit. */
FCD_API int fcd_edge_table(int64_t c0, int64_t C, int32_t* nm, void* stream);

/* Responsibility planes, built once per fit (the reference never re-estimates
 * mu, sigma, fcdiff/fit.py:232-237): for every patient correlation x = bt[c,u]
 * with N_k = N(x; mu_k, sigma_k) (fit.py:115),
 *   P[k][c][u] = N_k / (N_0 + N_1 + N_2)   (three planes, `planeStride` elements apart)
 *   L[c][u]    = log(N_0 + N_1 + N_2)      (may be NULL)
 * each with the shape and pitch of bt.  Then (fit.py:117-122, 409-444)
 *   log M_kl = L + log(a_l + b_l p_k),  a_l = (1 - eps_l)/2,  b_l = eps_l - a_l,
 * exactly and without underflow; the kernels below read these planes, not bt. */
FCD_API int fcd_resp_cache(const double* bt, int64_t C, int32_t U, int64_t pitchU,
                   const fcd_theta* theta_host, double* P, int64_t planeStride, double* L, void* stream);

/* Peak states of the posteriors (tier selection, DESIGN.md "Tiers"):
 *   fstate[c]   = k if qF[c,k] == 1.0 and the other two <= 2^-60, else 3;
 *   rstate[n][u] = s if qR[n,u,s] == 1.0 and the other <= 2^-60, else 2;
 *                  4 in the padding columns u >= U (row pitch pitchS, a multiple of 256).
 * A term whose weight is <= 2^-60 is below the rounding error of the sums it
 * would enter and is skipped by the kernels below. */
FCD_API int fcd_peak_states_F(const double* qF, int64_t C, uint8_t* fstate, void* stream);
FCD_API int fcd_peak_states_R(const double* qR, int32_t N, int32_t U, int64_t pitchS, uint8_t* rstate, void* stream);

/* MAP labels of a log-posterior array lq [n][width]: labels[i] = argmax_k lq[i][k]
 * (first maximum on ties).  width 3: template state of an edge (q_F of
 * fit.py:157-174); width 2: anomaly flag of a (region, patient) pair (q_R of
 * fit.py:176-198). */
FCD_API int fcd_map_labels(const double* lq, int64_t n, int32_t width, uint8_t* labels, void* stream);

/* K2 -- E-step for the template posterior; replaces `_update_lq_F`
 * (fcdiff/fit.py:157-174) + `_eval_q_R_w` (fit.py:382-406) + the patient half
 * of `_update_lps` (fit.py:115-122) + `_eval_M` (fit.py:409-444).
 *   lqF[c,k] = log gamma_k + sum_h logN_k(b[c,h])
 *            + sum_u sum_l w_l(n,m,u) log M_kl(bt[c,u]),  minus logsumexp_k.
 * P: planes of the local edge rows; qR [N][U][2] probabilities, rstate their peak
 * states; nm the shard's edge table.  Outputs lqF [C][3] and qF=exp(lqF) [C][3]. */
FCD_API int fcd_estep_qF(const double* S1, const double* S2, int32_t H,
                 const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                 const double* qR, const uint8_t* rstate, int64_t pitchS, int32_t N, const int32_t* nm,
                 const fcd_theta* theta_host, double* lqF, double* qF, void* stream);

/* The same E-step driven by the code pass of the previous M-step (fcd_code_plane / fcd_code_records:
 * valid while q_R has not changed since): the coded elements need no peak-state decoding and no
 * queue -- one code byte selects the constants of a running product -- and the elements with a
 * mixed region are taken from the row's key lists (half records: three differences of logs with
 * one weight; full records: nine logs).  code / counts / keysF / keysH / rowoff / Hh as written by
 * the code pass for the same local rows; rows whose edge was unpeaked then (counts[c][0] == 3 U)
 * are evaluated in full. */
FCD_API int fcd_estep_qF_coded(const double* S1, const double* S2, int32_t H,
                       const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                       const double* qR, int32_t N, const int32_t* nm,
                       const uint8_t* code, int64_t pitchQ, const int32_t* counts, const uint64_t* keysF,
                       const uint64_t* keysH, const int64_t* rowoff, const double* Hh,
                       const fcd_theta* theta_host, double* lqF, double* qF, void* stream);

/* The same launch for the E-step of the NEXT iteration, enqueued behind a device-resident (eta, epsilon) solve
 * (fcd_elm_coded_solve / fcd_elm_tiered_solve, fcdiff/fit.py:228-241) whose result the host has not read yet:
 * eta and epsilon are taken from `solver_state` on the device when the kernel starts (theta_host supplies gamma,
 * mu, sigma; its eta / epsilon are ignored); [eps_lo, eps_hi] is the epsilon box of that solve and sizes the
 * logarithm table.  The reference runs these steps one after the other (fit.py:73-79); here the E-step fills the
 * GPU while the host waits for the solve and the free energy.  The caller discards lqF / qF when the solve had not
 * finished in the launches enqueued before (state.done == 0), ended pinned at the box, or the fit stops. */
FCD_API int fcd_estep_qF_coded_solved(const double* S1, const double* S2, int32_t H,
                       const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                       const double* qR, int32_t N, const int32_t* nm,
                       const uint8_t* code, int64_t pitchQ, const int32_t* counts, const uint64_t* keysF,
                       const uint64_t* keysH, const int64_t* rowoff, const double* Hh,
                       const fcd_theta* theta_host, const void* solver_state, double eps_lo, double eps_hi,
                       double* lqF, double* qF, void* stream);

/* The uniform start (fcdiff/fit.py:84-102: lq_R = -ln 2, lq_F = -ln 3 everywhere).  With a CONSTANT q_R every
 * element has the same pair weights w (fit.py:382-406), and the sums of fit.py:165-173 / 489-511 factor through
 *   S9[c][k*3 + l] = sum_u log(a_l + b_l p_k(c,u))       (a_l, b_l: fit.py:427-444 relative to the total density)
 * -- nine running products per row instead of nine logarithms per element (csrc/fcd_uniform.cu):
 *   fcd_row_logsums: S9 [C][9] from the planes P_0, P_1 ([C][pitchU], 16-byte aligned rows);
 *   fcd_estep_qF_rowsums: `_update_lq_F` (fit.py:157-174) from S9 and w3_host = {q0^2, q1^2, 2 q0 q1};
 *   fcd_elm_rowsums: out1[0] = sum_c sum_k qF[c,k] sum_l w_l S9[c][k][l], the theta-dependent part of E_lM
 *     (fit.py:489-511); the theta-free part is (sum_l w_l) sum_c (sum_k qF[c,k]) sum_u L[c,u]. */
FCD_API int fcd_row_logsums(const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                    const fcd_theta* theta_host, double* S9, void* stream);
FCD_API int fcd_estep_qF_rowsums(const double* S1, const double* S2, int32_t H, const double* S9, int64_t C,
                         const double* w3_host, const fcd_theta* theta_host, double* lqF, double* qF, void* stream);
FCD_API int fcd_elm_rowsums(const double* S9, const double* qF, int64_t C, const double* w3_host, double* out1,
                    double* ws, void* stream);

/* Replica sweeps (BASELINE.json configs[4]; no reference counterpart -- the reference stops at one
 * fit, fcdiff/fit.py:56-82): the responsibility planes of ALL S subjects are built once
 * (fcd_resp_cache on every column of the (C, S) correlation matrix, fcd_transpose_patients), and a
 * relabelling selects from them (csrc/fcd_replica.cu):
 *   fcd_healthy_stats_cols: S1 / S2 of fcd_healthy_stats over the columns cols[0..H) of X [C][pitchS];
 *   fcd_gather_columns: dst[p][c][j] = src[p][c][cols[j]], j < U (columns U..pitchD-1 zero), p < nplanes;
 *   fcd_gather_rows: dst[p][j][:] = src[p][rows[j]][:] (rows of `pitch` doubles, 16-byte aligned).
 * cols / rows: int32 device arrays. */
FCD_API int fcd_healthy_stats_cols(const double* X, int64_t C, int64_t pitchS, const int32_t* cols, int32_t H,
                           double* S1, double* S2, void* stream);
FCD_API int fcd_gather_columns(const double* src, int64_t planeStrideS, int64_t pitchS, int32_t nplanes, int64_t C,
                       const int32_t* cols, int32_t U, double* dst, int64_t planeStrideD, int64_t pitchD,
                       void* stream);
FCD_API int fcd_gather_rows(const double* src, int64_t planeStrideS, int32_t nplanes, const int32_t* rows, int32_t U,
                    int64_t pitch, double* dst, int64_t planeStrideD, void* stream);

/* Patient-major copy of an edge-major plane: dst[u - u0][c] = src[c][u]
 * for u in [u0, u0+Ul), c in [0, C).  Built once per fit. */
FCD_API int fcd_transpose_patients(const double* bt, int64_t C, int32_t U, int64_t pitchU,
                           int32_t u0, int32_t Ul, double* btT, int64_t pitchC, void* stream);

/* K2b part 1 -- q_R-independent half of `_update_lq_R` (fcdiff/fit.py:187-194):
 *   W_l[c,u] = sum_k qF[c,k] * log M_kl(bt[c][u])   (PT[k][u][c] are the patient-major
 *   copies [Ul][pitchC] of the responsibility planes)
 * up to an additive per-(c,u) constant common to all l, which cancels in the
 * normalisation of fit.py:196.  Only l_0 - l_1 survives that
 * normalisation, so the tensor holds two numbers per edge-patient (16 bytes):
 *   WT[u][c] = { W_0 - W_2, W_2 - W_1 }     ([Ul][C][2], patient-major)
 * qF / fstate cover ALL edges.
 * PsT / kcache (both or neither): caller-kept [Ul][pitchC] plane of each edge's
 * dominant-state responsibility and the [C] states it was gathered for (255 =
 * never; set by the caller once); the call refreshes the columns of edges whose
 * state changed and then reads one coalesced plane instead of three. */
FCD_API int fcd_region_weights(const double* PT, int64_t planeStride, int32_t Ul, int64_t C, int64_t pitchC,
                       const double* qF, const uint8_t* fstate, double* PsT, uint8_t* kcache,
                       const fcd_theta* theta_host, double* WT, void* stream);

/* K2b part 2 -- Gauss-Seidel sweep of `_update_lq_R` (fcdiff/fit.py:176-198)
 * for the patients [u0, u0+Ul) from WT [Ul][C][2] (fcd_region_weights).  qR/lqR are the full [N][U][2] arrays; only
 * the [u0, u0+Ul) columns are read and written.  log_pi2_host = {log(1-pi),
 * log(pi)} (the 2-vector convention of test_fcdiff/test_fit.py:477-487).
 * edge_lookup: FCD_LOOKUP_REFERENCE reproduces nm_to_c(n, m) for all m != n
 * (fit.py:185-186), FCD_LOOKUP_SYMMETRIC the unordered pair. */
FCD_API int fcd_estep_qR(const double* WT, int64_t C, int32_t N, int32_t U, int32_t u0, int32_t Ul,
                 const double* log_pi2_host, int32_t edge_lookup,
                 double* qR, double* lqR, void* stream);

/* K2b fused -- the whole of `_update_lq_R` (fcdiff/fit.py:176-198) for the
 * patients [u0, u0+Ul) with edge_lookup = FCD_LOOKUP_REFERENCE, without the WT
 * tensor: every edge's weights W[c][l] = sum_k qF[c,k] log M_kl are computed once
 * inside the sweep from the dominant-state plane PsT (kept by
 * fcd_region_weights' refresh; row pitch pitchC even), streamed through shared
 * memory by TMA, and held in a shared-memory ring while the sweep's window
 * passes.  PT (3 planes) is read only for edges whose q_F is not peaked.
 * fstate: peak states of ALL edges, padded to pitchF >= pitchC bytes (pitchF %
 * 16 == 0).  3 <= N <= 1024.  Other arguments as fcd_estep_qR. */
FCD_API int fcd_estep_qR_fused(const double* PsT, const double* PT, int64_t planeStride, int64_t pitchC,
                       const double* qF, const uint8_t* fstate, int64_t pitchF,
                       int64_t C, int32_t N, int32_t U, int32_t u0, int32_t Ul,
                       const double* log_pi2_host, const fcd_theta* theta_host,
                       double* qR, double* lqR, void* stream);

/* Refreshes the dominant-state plane PsT / kcache (see fcd_region_weights) for the
 * current peak states without computing region weights. */
FCD_API int fcd_pstar_refresh(const double* PT, int64_t planeStride, int32_t Ul, int64_t C, int64_t pitchC,
                      const uint8_t* fstate, double* PsT, uint8_t* kcache, void* stream);

/* The same two entries on the E-step's EDGE-major planes Pe [3][C][pitchU] (fcd_resp_cache), patients
 * [u0, u0 + Ul): no patient-major copy of the planes is needed -- the dominant-state plane PsT [Ul][pitchC] is
 * gathered by one transposing pass over the rows of the edges whose state changed (all of them the first time),
 * and the three-plane path of an unpeaked edge reads Pe with a stride.  Same results as fcd_region_weights /
 * fcd_pstar_refresh on the transposed planes (fcdiff/fit.py:176-198 needs the weights of every edge of a patient;
 * the layout they are read from is this implementation's business). */
FCD_API int fcd_pstar_refresh_em(const double* Pe, int64_t planeStride, int64_t pitchU, int32_t u0, int32_t Ul, int64_t C,
                         int64_t pitchC, const uint8_t* fstate, double* PsT, uint8_t* kcache, void* stream);
FCD_API int fcd_region_weights_em(const double* Pe, int64_t planeStride, int64_t pitchU, int32_t u0, int32_t Ul, int64_t C,
                          int64_t pitchC, const double* qF, const uint8_t* fstate, double* PsT, uint8_t* kcache,
                          const fcd_theta* theta_host, double* WT, void* stream);

/* K3a -- M-step sums; replaces `_update_pi` / `_update_gamma`
 * (fcdiff/fit.py:208-220).  out[0..2] = sum_c exp(lqF[c,k]),
 * out[3] = sum_{n,u} exp(lqR[n,u,1]).  The caller divides by counts (after an
 * all-reduce of out[0..2] when edges are sharded). */
FCD_API int fcd_mstep_stats(const double* lqF, int64_t C, const double* lqR, int64_t NU,
                    double* out4, double* ws, void* stream);

/* K3b -- objective and analytic gradient of the (eta, epsilon) sub-problem;
 * replaces `_opt_fun` (fcdiff/fit.py:270-286) = `_update_lps` + `_eval_E_lM`
 * (fit.py:489-511) and `_eval_dE_dh` / `_eval_dE_de` / `_eval_dlM_dh` /
 * `_eval_dlM_de` (fit.py:600-697) in ONE pass over the responsibility planes
 * (one plane, 8 bytes per edge-patient, wherever the posteriors are peaked).
 *   out[0] = sum_c sum_k qF[c,k] sum_u sum_l w_l log(a_l + b_l p_k)
 *            E_lM = out[0] + fcd_elm_const (the part independent of eta, epsilon)
 *   out[1] = dE/d eta, out[2] = dE/d epsilon   (of E = -E_lM; only if want_grad) */
FCD_API int fcd_elm_obj_grad(const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                     const double* qF, const uint8_t* fstate, const double* qR, const uint8_t* rstate,
                     int64_t pitchS, int32_t N, const int32_t* nm,
                     const fcd_theta* theta_host, int32_t want_grad,
                     double* out3, double* ws, void* stream);

/* out1[0] = sum_c (sum_k qF[c,k]) sum_u (sum_l w_l) L[c,u]: the part of E_lM
 * (fit.py:489-511) that does not depend on (eta, epsilon). */
FCD_API int fcd_elm_const(const double* L, int64_t C, int32_t U, int64_t pitchU,
                  const double* qF, const uint8_t* fstate, const double* qR, const uint8_t* rstate,
                  int64_t pitchS, int32_t N, const int32_t* nm, double* out1, double* ws, void* stream);

/* K3b over the coded dominant-state plane (DESIGN.md "Coded plane").  Between two
 * E-steps the optimiser evaluates E_lM(eta, epsilon) many times with q_F, q_R fixed
 * (fit.py:228-241): the code pass turns the planes into what an evaluation needs
 * and nothing else --
 *   PsE  [C][pitchQ] f64: responsibility of each edge's dominant state (rows re-gathered
 *        from P only when kcache[c] != fstate[c]; kcache starts at 255, PsE zero-filled);
 *        pitchQ = fcd_code_pitch(U): U rounded up to 16, so that code rows are 16-byte aligned;
 *   code [C][pitchQ] u8 (+ 16 bytes of slack): l* in 0..2 for elements whose edge and
 *        regions are peaked (weight exactly 1, one log); 4 + s for an element with ONE undecided,
 *        normalised region and the other in state s -- it counts as l = 2 with weight 1 and
 *        adds a half record; 3 for all others;
 *   half records {p, +-q_s} (16 bytes, sign bit = s), nh of them: the correction
 *        q_s (log M_s - log M_2) of the elements coded 4 + s (fit.py:382-406 with q_0 + q_1 = 1);
 *   records {p, w_0, w_1, w_2} (32 bytes) for all other elements, nd of them
 * -- and fcd_elm_coded reduces them: 9 bytes per edge-patient per evaluation, no row
 * structure, no partition of the elements.
 *   fcd_plane_sum: out1[0] = sum of a [C][U] plane -- the total of the L plane, once per cache;
 *   fcd_code_plane: refreshes PsE, writes code, counts[c] = {records, half records} of row c
 *     (int32 x 2), blockoff: scratch of 4 * fcd_bucket_blocks(C) int64 (totals and exclusive
 *     prefix sums of blocks of 16 rows), total2 = {nd, nh} (as doubles);
 *   fcd_code_records: D receives the nd records, Hh the nh half records; keysF [nd] / keysH [nh]
 *     uint64 = (tag << 48 | c << 16 | u) of every record / half record and rowoff [C][2] int64 =
 *     first key of each row in either list are kept for fcd_estep_qF_coded (not written when
 *     nd + nh == 0); out1[0] = the theta-free part of E_lM
 *     (same value as fcd_elm_const), formed as Lsum[0] (device, from fcd_plane_sum)
 *     corrected by the record elements only;
 *   fcd_elm_coded: nE = C * pitchQ elements of PsE / code; out3 as fcd_elm_obj_grad. */
FCD_API int64_t fcd_bucket_blocks(int64_t C);
FCD_API int64_t fcd_code_pitch(int32_t U);
FCD_API int fcd_plane_sum(const double* X, int64_t C, int32_t U, int64_t pitchU, double* out1, double* ws,
                  void* stream);
FCD_API int fcd_code_plane(const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                   const uint8_t* fstate, const uint8_t* rstate, int64_t pitchS, const int32_t* nm,
                   double* PsE, uint8_t* kcache, uint8_t* code, int64_t pitchQ, int32_t* counts, int64_t* blockoff,
                   double* total2, void* stream);
FCD_API int fcd_code_records(const double* P, int64_t planeStride, const double* PsE, const uint8_t* code,
                     int64_t pitchQ, const double* L, const double* Lsum, int64_t C, int32_t U, int64_t pitchU,
                     const double* qF, const uint8_t* fstate, const double* qR, const uint8_t* rstate, int64_t pitchS,
                     int32_t N, const int32_t* nm, const int32_t* counts, const int64_t* blockoff,
                     uint64_t* keysF, uint64_t* keysH, int64_t* rowoff, double* D, int64_t nd, double* Hh, int64_t nh,
                     double* out1, double* ws, void* stream);
FCD_API int fcd_elm_coded(const double* PsE, const uint8_t* code, int64_t nE, const double* D, int64_t nd,
                  const double* Hh, int64_t nh, const fcd_theta* theta_host, int32_t want_grad, double* out3,
                  double* ws, void* stream);

/* ------------------------------------------- device-resident (eta, epsilon) solver */
/* Replaces the host optimiser of fcdiff/fit.py:228-241 (scipy.optimize.minimize on
 * [1e-5, 1 - 1e-5]^2, objective fit.py:270-286 = -E_lM, gradient fit.py:600-697): a safeguarded
 * projected Newton iteration whose step is taken by the LAST CTA of every evaluation kernel
 * (csrc/fcd_solver.cuh) -- the evaluations of one solve are enqueued back to back and read the
 * iterate from device memory; nothing returns to the host between them.  With edge shards the
 * same CTA first exchanges the six partial sums over the peer windows (fcd_comm_window_*), in rank
 * order, so every rank takes bit-identical steps.
 *   state: fcd_solver_state_bytes() of device memory, set by fcd_solver_init (box lo/hi per
 *     parameter, start point, `tol`: a step with |dx_i| <= tol * min(x_i, 1 - x_i) for both parameters is
 *     taken without another evaluation and ends the solve -- it leaves a relative error ~ tol^2 behind;
 *     max_evals bounds the evaluations);
 *   fcd_elm_coded_solve / fcd_elm_tiered_solve: n_launches evaluations of the coded-plane /
 *     tiered form (arguments as fcd_elm_coded / fcd_elm_obj_grad) at the state's iterate; launches
 *     after the solve has finished exit at once.  [eps_lo, eps_hi] must contain the box of epsilon
 *     (it sizes the shared-memory logarithm table).  konst (device, may be NULL): this rank's
 *     theta-free part of E_lM (fcd_code_records / fcd_elm_const), added to the objective.
 *     published_host: fcd_host_mapped_alloc(fcd_solver_published_bytes()) -- every launch stores
 *     the state there followed by its sequence number seq0 + i;
 *   fcd_solver_wait spins (host) until launch `seq` has published and copies the state
 *     (fcd_solver_state_bytes(), layout `fcd_solver_state` below) to state_out_host. */
typedef struct fcd_solver_state {
    double x[2];         /* (eta, epsilon): next evaluation point; the solution when done */
    double xprev[2];
    double fprev;
    double f;            /* -E_lM at x when done (incl. the theta-free part) */
    double g[2];         /* gradient at the last evaluated point */
    double lo[2], hi[2];
    double tol;
    double step;
    int32_t have_prev, nfev, nback;
    int32_t done;        /* 0 running; 1 converged; 2 stopped (budget / no descent / non-finite); 3 peer time-out */
    int32_t max_evals, pad_;
} fcd_solver_state;
FCD_API int64_t fcd_solver_state_bytes(void);
FCD_API int64_t fcd_solver_published_bytes(void);
FCD_API int fcd_host_mapped_alloc(int64_t bytes, void** out_host);
FCD_API int fcd_host_mapped_free(void* p_host);
FCD_API int fcd_solver_init(void* state, double eta0, double eps0, const double* lo2_host, const double* hi2_host,
                    double tol, int32_t max_evals, void* stream);
FCD_API int fcd_elm_coded_solve(const double* PsE, const uint8_t* code, int64_t nE, const double* D, int64_t nd,
                        const double* Hh, int64_t nh, double eps_lo, double eps_hi, void* state, const double* konst,
                        void* const* windows_host, int32_t rank, int32_t world, void* published_host, uint64_t seq0,
                        int32_t n_launches, double* ws, void* stream);
FCD_API int fcd_elm_tiered_solve(const double* P, int64_t planeStride, int64_t C, int32_t U, int64_t pitchU,
                         const double* qF, const uint8_t* fstate, const double* qR, const uint8_t* rstate,
                         int64_t pitchS, int32_t N, const int32_t* nm, double eps_lo, double eps_hi, void* state,
                         const double* konst, void* const* windows_host, int32_t rank, int32_t world,
                         void* published_host, uint64_t seq0, int32_t n_launches, double* ws, void* stream);
/* Host-side twins of fcd_solver_init and of the optimiser transition the kernels' last CTA runs (same code):
 * sums6_host = {obj, ge, G_2, Q_0 + Q_1, Q_2, theta-free part} at state.x (csrc/fcd_solver.cuh). */
FCD_API int fcd_solver_init_host(void* state_host, double eta0, double eps0, const double* lo2_host,
                         const double* hi2_host, double tol, int32_t max_evals);
FCD_API int fcd_solver_step_host(void* state_host, const double* sums6_host);
FCD_API int fcd_solver_wait(const void* published_host, uint64_t seq, void* state_out_host, int32_t timeout_ms);

/* K4 -- free-energy terms; replaces `_eval_energy` and `_eval_E_*`
 * (fcdiff/fit.py:142-155, 447-539).  out[0..5] = E_lp_F, E_lp_B_g_F, E_lp_R,
 * E_lM, E_lq_F, E_lq_R over the local edge shard (terms 2 and 5 involve q_R
 * only and are complete on every rank).  qF = exp(lqF), qR = exp(lqR) as left
 * by fcd_estep_qF / fcd_estep_qR (fit.py:146-147); lqF / qF / S1 / S2 are the
 * shard's C rows.  `elm` = E_lM of the shard (fcd_elm_obj_grad + fcd_elm_const
 * at the same q_F, q_R, theta); or, with `solver_state` != NULL (device, fcd_solver_init): the E_lM
 * of the device-resident (eta, epsilon) solve enqueued before this call on the same stream (the
 * GLOBAL value: pass it on one rank only) -- NaN if that solve has not finished. */
FCD_API int fcd_energy_terms(const double* S1, const double* S2, int32_t H,
                     const double* lqF, const double* qF, int64_t C,
                     const double* lqR, const double* qR, int32_t N, int32_t U,
                     const fcd_theta* theta_host, double elm, const void* solver_state,
                     double* out6, double* ws, void* stream);

/* K3c -- per-state sufficient statistics of the correlations for the control
 * and the patient group (north_star subsystem 3; the reference ships only
 * disabled pieces of a mu / sigma update, fcdiff/fit.py:232-237, 542-597,
 * 709-733; doc/methods.rst:715-944).  out[0..8] controls: n_j = H sum_c qF[c,j],
 * sum qF[c,j] S1[c], sum qF[c,j] S2[c]; out[9..17] patients: sum R_j, sum R_j x,
 * sum R_j x^2 with R_j(c,u) the posterior weight that patient edge (c,u) is in
 * state j.  Pooled, they are the EM (lower-bound) update of mu_j, sigma_j. */
FCD_API int fcd_state_moments(const double* S1, const double* S2, int32_t H,
                      const double* bt, const double* P, int64_t planeStride,
                      int64_t C, int32_t U, int64_t pitchU,
                      const double* qF, const double* qR, int32_t N, const int32_t* nm,
                      const fcd_theta* theta_host, double* out18, double* ws, void* stream);

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

/* W for `_update_lq_R` from a materialised lM: W_l = sum_k qF[c,k] lM[c,u,k,l], stored as
 * WT[u][c] = { W_0 - W_2, W_2 - W_1 } ([U][C][2], the layout fcd_estep_qR reads). */
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
